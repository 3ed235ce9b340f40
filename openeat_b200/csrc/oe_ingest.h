// Native PCM ingest (host code, no CUDA): RIFF/WAVE header parsing and multi-threaded reads of a batch of utterances
// straight into one packed int16 buffer -- the part of _extract_feature in front of the kernels
// (openeat/dataset/dataset.py:55-75: sox_io_backend.info + torchaudio.load per utterance inside DataLoader workers,
// openeat/bin/train.py:110-116).  Included by oe_frontend.cu; the C ABI is declared in include/openeat_frontend.h.
#pragma once

#include <emmintrin.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>

#include "oe_flac.h"

namespace oe_ing {

// Persistent pool: the workers sleep between calls (creating 2 x 32 threads per batch costs more than reading it).
class Pool {
  public:
    explicit Pool(int threads) : n_threads_(std::max(1, threads)) {
        for (int i = 1; i < n_threads_; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            quit_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // runs fn(i) for i in [0, n) on the pool (the caller takes part); returns when all are done
    void run(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn;
            n_ = n;
            next_.store(0);
            pending_ = n;
            ++epoch_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    void work() {
        int did = 0;
        for (int i = next_.fetch_add(1); i < n_; i = next_.fetch_add(1)) {
            (*fn_)(i);
            ++did;
        }
        if (did) {
            std::lock_guard<std::mutex> lk(m_);
            pending_ -= did;
            if (pending_ == 0) done_.notify_all();
        }
    }
    void loop() {
        long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return quit_ || epoch_ != seen; });
                if (quit_) return;
                seen = epoch_;
            }
            work();
        }
    }
    int n_threads_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* fn_ = nullptr;
    std::atomic<int> next_{0};
    int n_ = 0, pending_ = 0;
    long epoch_ = 0;
    bool quit_ = false;
};

constexpr int kBounceBytes = 128 * 1024;      // per reader thread: stays in the core's L2

// dst may be unaligned only in its last partial 16 bytes' worth: the head is brought to 16-byte alignment with plain
// stores, the body goes out with MOVNTDQ (no read-for-ownership of the destination lines), the tail with plain stores
inline void stream_copy(char* dst, const char* src, size_t n) {
    size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    if (head > n) head = n;
    memcpy(dst, src, head);
    dst += head;
    src += head;
    n -= head;
    const size_t body = n & ~(size_t)63;
    for (size_t i = 0; i < body; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
    }
    memcpy(dst + body, src + body, n - body);
    _mm_sfence();
}

struct WavInfo {
    int sample_rate = 0, channels = 0, bits = 0;
    int64_t data_off = 0, frames = 0;        // byte offset of the PCM, frames in the file
    bool flac = false;                       // FLAC stream: decoded by the reader thread (oe_flac.h) instead of copied
};

}  // namespace oe_ing

// One asynchronous batch (oe_ingest_submit): owns copies of the request and the result arrays.
struct oe_ingest_job {
    oe_ingest* owner = nullptr;
    std::vector<std::string> paths;
    std::vector<const char*> cpaths;
    std::vector<double> starts, ends;
    int16_t* dst = nullptr;
    int64_t capacity = 0, total = 0;
    std::vector<int64_t> offsets;
    std::vector<int32_t> lens, rates, status;
    std::vector<std::string> errors;
    bool done = false;
    // FLAC batches for the GPU decoder (oe_flac_submit): the driver runs oe_flac_pack instead of probe + read
    bool flac = false;
    void* comp = nullptr;
    struct oe_flac_frame* frames = nullptr;
    int64_t comp_capacity = 0, frames_capacity = 0, comp_bytes = 0, n_frames = 0;
    std::vector<int64_t> comp_offsets;
    int rc = 0;
    std::string message;
};

struct oe_ingest {
    int threads;
    bool direct_read = false;               // OE_INGEST_DIRECT=1: pread straight into the destination (A/B timing only)
    bool flac_verify_md5 = false;           // OE_FLAC_VERIFY_MD5=1: also check the STREAMINFO signature (CRCs are always checked)
    oe_ing::Pool* pool;
    std::vector<std::string> errors;        // per entry of the most recent probe / read
    // the probe keeps every file open with its parsed header: the read that follows neither re-opens nor re-parses
    std::vector<int> fds;
    std::vector<oe_ing::WavInfo> infos;
    std::vector<int64_t> first;
    // asynchronous jobs: a driver thread takes them in order and runs probe -> layout -> read on the pool, so the
    // caller (a Python thread holding the GIL) only submits and, later, waits
    std::thread driver;
    std::mutex qm;
    std::condition_variable qcv, dcv;
    std::vector<oe_ingest_job*> queue;
    bool quit = false;
};

namespace oe_ing {

inline uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
inline uint32_t rd16(const unsigned char* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }

// Parses the chunk list of a RIFF/WAVE file.  Returns "" or the reason the file cannot be ingested.
inline std::string parse_wav(int fd, const char* path, WavInfo& w) {
    unsigned char h[12];
    if (pread(fd, h, 12, 0) != 12) return std::string(path) + ": too short for a RIFF header";
    struct stat st;
    if (fstat(fd, &st) != 0) return std::string(path) + ": " + strerror(errno);
    const int64_t id3 = oe_flac::id3v2_bytes(h, 12);                     // a tag in front of a FLAC stream is skipped
    unsigned char magic[4] = {0, 0, 0, 0};
    if (id3 > 0 && pread(fd, magic, 4, id3) != 4) memset(magic, 0, 4);
    if (memcmp(h, "fLaC", 4) == 0 || memcmp(magic, "fLaC", 4) == 0) {
        // STREAMINFO is always the first metadata block (RFC 9639 section 8.1): 4 + 4 + 34 bytes hold everything the
        // probe needs; a stream that does not announce its length is decoded once to count it
        unsigned char head[42];
        if (pread(fd, head, 42, memcmp(h, "fLaC", 4) == 0 ? 0 : id3) != 42) return std::string(path) + ": too short for a FLAC STREAMINFO block";
        oe_flac::Info fi;
        unsigned char one[42];
        memcpy(one, head, 42);
        one[4] |= 0x80;                                                  // parse the first block only
        std::string err = oe_flac::parse_streaminfo(one, 42, fi);
        if (!err.empty()) return std::string(path) + ": " + err;
        if (fi.bits != 16)
            return std::string(path) + ": " + std::to_string(fi.bits) + "-bit FLAC (the native ingest fills an int16 buffer: 16-bit streams only; read_wav decodes the others to fp32)";
        if (fi.total == 0) {
            std::vector<unsigned char> all((size_t)st.st_size);
            if (pread(fd, all.data(), all.size(), 0) != (ssize_t)all.size()) return std::string(path) + ": short read";
            int64_t n = 0;
            err = oe_flac::decode(all.data(), (int64_t)all.size(), 0, 0, 0, nullptr, false, fi, &n);
            if (!err.empty()) return std::string(path) + ": " + err;
            fi.total = n;
        }
        w.flac = true;
        w.sample_rate = fi.sample_rate;
        w.channels = fi.channels;
        w.bits = fi.bits;
        w.frames = fi.total;
        w.data_off = 0;
        return "";
    }
    if (memcmp(h, "RIFF", 4) != 0 || memcmp(h + 8, "WAVE", 4) != 0)
        return std::string(path) + ": not a RIFF/WAVE or FLAC file (the native ingest reads 16-bit PCM wav and 16-bit FLAC)";
    int64_t pos = 12;
    bool have_fmt = false;
    while (pos + 8 <= st.st_size) {
        unsigned char c[8];
        if (pread(fd, c, 8, pos) != 8) break;
        const int64_t size = rd32(c + 4);
        if (memcmp(c, "fmt ", 4) == 0) {
            unsigned char f[40];
            const int n = (int)std::min<int64_t>(size, 40);
            if (n < 16 || pread(fd, f, n, pos + 8) != n) return std::string(path) + ": truncated fmt chunk";
            int format = (int)rd16(f);
            w.channels = (int)rd16(f + 2);
            w.sample_rate = (int)rd32(f + 4);
            w.bits = (int)rd16(f + 14);
            if (format == 0xFFFE && n >= 26) format = (int)rd16(f + 24);      // WAVE_FORMAT_EXTENSIBLE: sub-format GUID
            if (format != 1)
                return std::string(path) + ": wav format tag " + std::to_string(format) + " is not integer PCM (16-bit PCM only)";
            if (w.bits != 16)
                return std::string(path) + ": " + std::to_string(w.bits) + "-bit samples (the native ingest reads 16-bit PCM only)";
            if (w.channels < 1 || w.sample_rate <= 0) return std::string(path) + ": bad channel count / sample rate";
            have_fmt = true;
        } else if (memcmp(c, "data", 4) == 0) {
            if (!have_fmt) return std::string(path) + ": data chunk before fmt chunk";
            w.data_off = pos + 8;
            const int64_t avail = std::min<int64_t>(size, st.st_size - w.data_off);   // streamed files write 0 / 0xFFFFFFFF sizes
            const int64_t bytes = (size == 0 || size == 0xFFFFFFFFll) ? st.st_size - w.data_off : avail;
            w.frames = bytes / (2 * w.channels);
            return "";
        }
        pos += 8 + size + (size & 1);
    }
    return std::string(path) + ": no data chunk";
}

// frames [first, first + count) of the segment the reference would load (dataset.py:64-72: frame_offset = int(start * sr),
// num_frames = int(end * sr) - frame_offset; a plain path loads everything), clipped to the file
inline void segment(const WavInfo& w, double start, double end, bool has_seg, int64_t& first, int64_t& count) {
    if (!has_seg) {
        first = 0;
        count = w.frames;
        return;
    }
    const int64_t s = (int64_t)(start * w.sample_rate), e = (int64_t)(end * w.sample_rate);
    first = std::min<int64_t>(std::max<int64_t>(s, 0), w.frames);
    count = std::max<int64_t>(0, std::min<int64_t>(e - s, w.frames - first));
}

}  // namespace oe_ing
