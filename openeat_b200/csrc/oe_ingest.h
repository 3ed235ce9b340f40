// Native PCM ingest (host code, no CUDA): RIFF/WAVE header parsing and multi-threaded reads of a batch of utterances
// straight into one packed int16 buffer -- the part of _extract_feature in front of the kernels
// (openeat/dataset/dataset.py:55-75: sox_io_backend.info + torchaudio.load per utterance inside DataLoader workers,
// openeat/bin/train.py:110-116).  Included by oe_frontend.cu; the C ABI is declared in include/openeat_frontend.h.
#pragma once

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <string>
#include <thread>

struct oe_ingest {
    int threads;
    std::vector<std::string> errors;        // per entry of the most recent probe / read
};

namespace oe_ing {

struct WavInfo {
    int sample_rate = 0, channels = 0, bits = 0;
    int64_t data_off = 0, frames = 0;        // byte offset of the PCM, frames in the file
};

inline uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
inline uint32_t rd16(const unsigned char* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }

// Parses the chunk list of a RIFF/WAVE file.  Returns "" or the reason the file cannot be ingested.
inline std::string parse_wav(int fd, const char* path, WavInfo& w) {
    unsigned char h[12];
    if (pread(fd, h, 12, 0) != 12) return std::string(path) + ": too short for a RIFF header";
    if (memcmp(h, "fLaC", 4) == 0)
        return std::string(path) + ": FLAC is not supported by the native ingest (16-bit PCM RIFF/WAVE only): decode it first (sox / ffmpeg) or pass decoded arrays";
    if (memcmp(h, "RIFF", 4) != 0 || memcmp(h + 8, "WAVE", 4) != 0)
        return std::string(path) + ": not a RIFF/WAVE file (the native ingest reads 16-bit PCM wav only)";
    struct stat st;
    if (fstat(fd, &st) != 0) return std::string(path) + ": " + strerror(errno);
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

template <class Fn>
inline void parallel_for(int threads, int n, Fn fn) {
    std::atomic<int> next(0);
    auto work = [&]() {
        for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
    };
    const int t = std::max(1, std::min(threads, n));
    std::vector<std::thread> pool;
    for (int i = 1; i < t; ++i) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
}

}  // namespace oe_ing
