// FLAC decoding on the host (no CUDA): what torchaudio.load does for the .flac lists of the LibriSpeech recipe
// (openeat/dataset/dataset.py:62-72 through libsox / libFLAC, third-party code that is neither in the reference tree nor
// in this image).  Written from the published format, RFC 9639: stream and metadata layout (section 8), frame header
// with its CRC-8 (9.1), constant / verbatim / fixed / linear-predictor subframes, wasted bits, partitioned Rice residuals
// with 4- and 5-bit parameters and escape partitions (9.2), stereo decorrelation (4.2), the frame's CRC-16 (9.3) and the
// MD5 signature of the decoded samples (8.2).  Pinned by the RFC's three worked examples (tests/test_flac.py).
// Included by oe_frontend.cu; the C ABI (oe_flac_info / oe_flac_decode) is declared in include/openeat_frontend.h.
#pragma once

#include <cstdint>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace oe_flac {

struct Info {
    int sample_rate = 0, channels = 0, bits = 0, min_block = 0, max_block = 0;
    int64_t total = 0;                       // samples per channel, 0 = not announced
    unsigned char md5[16] = {0};
    int64_t audio_off = 0;                   // byte offset of the first frame
};

// ---- MD5 (RFC 1321), for the STREAMINFO signature ----
class Md5 {
  public:
    Md5() { reset(); }
    void reset() {
        a_ = 0x67452301u, b_ = 0xefcdab89u, c_ = 0x98badcfeu, d_ = 0x10325476u;
        len_ = 0;
        fill_ = 0;
    }
    void update(const unsigned char* p, size_t n) {
        len_ += n;
        if (fill_) {
            const size_t take = std::min(n, (size_t)64 - fill_);
            memcpy(buf_ + fill_, p, take);
            fill_ += take, p += take, n -= take;
            if (fill_ < 64) return;
            block(buf_);
            fill_ = 0;
        }
        for (; n >= 64; p += 64, n -= 64) block(p);
        memcpy(buf_, p, n);
        fill_ = n;
    }
    void finish(unsigned char out[16]) {
        const uint64_t bits = len_ * 8;
        unsigned char pad[72] = {0x80};
        const size_t padn = (fill_ < 56 ? 56 : 120) - fill_;
        update(pad, padn);
        unsigned char l[8];
        for (int i = 0; i < 8; ++i) l[i] = (unsigned char)(bits >> (8 * i));
        update(l, 8);
        const uint32_t s[4] = {a_, b_, c_, d_};
        for (int i = 0; i < 16; ++i) out[i] = (unsigned char)(s[i >> 2] >> (8 * (i & 3)));
    }

  private:
    static uint32_t rol(uint32_t x, int s) { return x << s | x >> (32 - s); }
    void block(const unsigned char* p) {
        static const uint32_t K[64] = {
            0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af,
            0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa,
            0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8,
            0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
            0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97,
            0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1,
            0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
        static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,
                                  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
        uint32_t m[16];
        for (int i = 0; i < 16; ++i) m[i] = (uint32_t)p[4 * i] | (uint32_t)p[4 * i + 1] << 8 | (uint32_t)p[4 * i + 2] << 16 | (uint32_t)p[4 * i + 3] << 24;
        uint32_t a = a_, b = b_, c = c_, d = d_;
        for (int i = 0; i < 64; ++i) {
            uint32_t f;
            int g;
            if (i < 16) f = (b & c) | (~b & d), g = i;
            else if (i < 32) f = (d & b) | (~d & c), g = (5 * i + 1) & 15;
            else if (i < 48) f = b ^ c ^ d, g = (3 * i + 5) & 15;
            else f = c ^ (b | ~d), g = (7 * i) & 15;
            const uint32_t t = d;
            d = c;
            c = b;
            b = b + rol(a + f + K[i] + m[g], S[i]);
            a = t;
        }
        a_ += a, b_ += b, c_ += c, d_ += d;
    }
    uint32_t a_, b_, c_, d_;
    uint64_t len_;
    size_t fill_;
    unsigned char buf_[64];
};

// ---- CRCs (MSB first, no reflection, initial value 0): x^8+x^2+x+1 for the header, x^16+x^15+x^2+1 for the frame ----
struct CrcTables {
    uint8_t t8[256];
    uint16_t t16[256];
    CrcTables() {
        for (int i = 0; i < 256; ++i) {
            uint8_t c = (uint8_t)i;
            uint16_t w = (uint16_t)(i << 8);
            for (int k = 0; k < 8; ++k) {
                c = (uint8_t)((c & 0x80) ? (c << 1) ^ 0x07 : c << 1);
                w = (uint16_t)((w & 0x8000) ? (w << 1) ^ 0x8005 : w << 1);
            }
            t8[i] = c;
            t16[i] = w;
        }
    }
};
inline const CrcTables& crc_tables() {
    static const CrcTables t;
    return t;
}
inline uint8_t crc8(const unsigned char* p, size_t n) {
    const CrcTables& t = crc_tables();
    uint8_t c = 0;
    for (size_t i = 0; i < n; ++i) c = t.t8[c ^ p[i]];
    return c;
}
inline uint16_t crc16(const unsigned char* p, size_t n) {
    const CrcTables& t = crc_tables();
    uint16_t c = 0;
    for (size_t i = 0; i < n; ++i) c = (uint16_t)((c << 8) ^ t.t16[(c >> 8) ^ p[i]]);
    return c;
}

// ---- bit reader: a 64-bit window kept left-aligned, refilled a byte at a time; reads past the end deliver zeros and
// raise `over`, which the frame loop turns into an error (no reads outside [p, end)) ----
struct Bits {
    const unsigned char *p, *end;
    uint64_t win = 0;
    int have = 0;                            // valid bits at the top of win
    bool over = false;
    Bits(const unsigned char* b, const unsigned char* e) : p(b), end(e) {}
    inline void refill() {
        if (end - p >= 8) {                  // eight bytes at once: the bits below `have` are the stream's own next bits, so
            uint64_t w;                      // OR-ing them in again at the next refill changes nothing
            memcpy(&w, p, 8);
            win |= __builtin_bswap64(w) >> have;
            const int nb = (64 - have) >> 3;
            p += nb;
            have += nb * 8;
            return;
        }
        while (have <= 56) {
            uint64_t byte = 0;
            if (p < end) byte = *p;
            else if (p >= end + 8) { over = true; }
            ++p;
            win |= byte << (56 - have);
            have += 8;
        }
    }
    inline uint32_t u(int n) {               // 0 <= n <= 32
        if (n == 0) return 0;
        if (have < n) refill();
        const uint32_t v = (uint32_t)(win >> (64 - n));
        win <<= n;
        have -= n;
        return v;
    }
    inline int32_t s(int n) {                // 0 <= n <= 32 (33-bit side samples go through s64)
        if (n == 0) return 0;
        const uint32_t v = u(n);
        return (int32_t)(v << (32 - n)) >> (32 - n);
    }
    inline int64_t s64(int n) {              // n <= 33
        if (n <= 32) return s(n);
        const int64_t hi = s(n - 32 + 16);   // top (n - 16) bits, signed
        return hi * 65536 + (int64_t)u(16);
    }
    inline uint32_t unary() {                // zeros before the next one bit
        uint32_t q = 0;
        for (;;) {
            if (have == 0) refill();
            if (win == 0) {                  // all `have` bits are zero
                q += (uint32_t)have;
                have = 0;
                if (over || q > (1u << 24)) { over = true; return q; }
                continue;
            }
            const int z = __builtin_clzll(win);
            if (z >= have) {                 // the one bit found lies in the zero padding below the valid bits
                q += (uint32_t)have;
                win = 0;
                have = 0;
                continue;
            }
            q += (uint32_t)z;
            win = z == 63 ? 0 : win << (z + 1);
            have -= z + 1;
            return q;
        }
    }
    // bytes consumed so far, counted from `base` (call only when byte aligned)
    inline const unsigned char* byte_pos() const { return p - have / 8; }
    inline void align() {
        const int drop = have & 7;
        win <<= drop;
        have -= drop;
    }
};

// Length of an ID3v2 tag in front of the stream (some taggers prepend one; libFLAC skips it too), 0 if there is none.
// `d` needs to hold the tag's 10-byte header only.
inline int64_t id3v2_bytes(const unsigned char* d, int64_t size) {
    if (size < 10 || d[0] != 'I' || d[1] != 'D' || d[2] != '3' || d[3] == 0xFF || d[4] == 0xFF || ((d[6] | d[7] | d[8] | d[9]) & 0x80)) return 0;
    return 10 + ((int64_t)d[6] << 21 | (int64_t)d[7] << 14 | (int64_t)d[8] << 7 | d[9]) + ((d[5] & 0x10) ? 10 : 0);
}

inline std::string parse_streaminfo(const unsigned char* d, int64_t size, Info& info) {
    int64_t pos = id3v2_bytes(d, size);
    if (size - pos < 42 || memcmp(d + pos, "fLaC", 4) != 0) return "not a FLAC stream";
    pos += 4;
    bool have = false;
    for (;;) {
        if (pos + 4 > size) return "truncated metadata";
        const int last = d[pos] >> 7, kind = d[pos] & 0x7F;
        const int64_t len = (int64_t)d[pos + 1] << 16 | (int64_t)d[pos + 2] << 8 | d[pos + 3];
        if (kind == 127) return "invalid metadata block type";
        if (pos + 4 + len > size) return "truncated metadata";
        if (kind == 0) {
            if (len < 34) return "short STREAMINFO block";
            const unsigned char* b = d + pos + 4;
            info.min_block = b[0] << 8 | b[1];
            info.max_block = b[2] << 8 | b[3];
            info.sample_rate = b[10] << 12 | b[11] << 4 | b[12] >> 4;
            info.channels = ((b[12] >> 1) & 7) + 1;
            info.bits = ((b[12] & 1) << 4 | b[13] >> 4) + 1;
            info.total = (int64_t)(b[13] & 15) << 32 | (int64_t)b[14] << 24 | (int64_t)b[15] << 16 | (int64_t)b[16] << 8 | b[17];
            memcpy(info.md5, b + 18, 16);
            have = true;
        }
        pos += 4 + len;
        if (last) break;
    }
    if (!have) return "no STREAMINFO block";
    if (info.sample_rate <= 0) return "sample rate 0";
    if (info.bits < 4) return "fewer than 4 bits per sample";
    info.audio_off = pos;
    return "";
}

// One subframe of n samples, `bps` bits each, into out[0 .. n) (int64 range is needed only for 33-bit side channels
// of 32-bit streams; everything is kept in int64 and narrowed by the caller).
inline const char* decode_subframe(Bits& b, int n, int bps, int64_t* out) {
    if (b.u(1)) return "subframe padding bit set";
    const int kind = (int)b.u(6);
    int wasted = 0;
    if (b.u(1)) {
        wasted = (int)b.unary() + 1;
        if (wasted >= bps) return "wasted bits exceed the sample size";
        bps -= wasted;
    }
    if (kind == 0) {
        const int64_t v = b.s64(bps);
        for (int i = 0; i < n; ++i) out[i] = v;
    } else if (kind == 1) {
        for (int i = 0; i < n; ++i) out[i] = b.s64(bps);
    } else if ((kind >= 8 && kind <= 12) || kind >= 32) {
        const int order = kind >= 32 ? kind - 31 : kind - 8;
        if (order > n) return "predictor order exceeds the block size";
        for (int i = 0; i < order; ++i) out[i] = b.s64(bps);
        int shift = 0;
        int32_t coef[32];
        if (kind >= 32) {
            const int prec = (int)b.u(4) + 1;
            if (prec == 16) return "reserved predictor precision";
            shift = b.s(5);
            if (shift < 0) return "negative predictor shift";
            for (int i = 0; i < order; ++i) coef[i] = b.s(prec);
        }
        // residual
        const int method = (int)b.u(2);
        if (method > 1) return "reserved residual coding method";
        const int pbits = method == 0 ? 4 : 5;
        const int porder = (int)b.u(4);
        if (porder && ((n >> porder) << porder) != n) return "block size not divisible by the partition count";
        int at = order;
        for (int part = 0; part < (1 << porder); ++part) {
            const int cnt = (n >> porder) - (part == 0 ? order : 0);
            if (cnt < 0) return "partition shorter than the predictor order";
            const int k = (int)b.u(pbits);
            if (k == (1 << pbits) - 1) {
                const int raw = (int)b.u(5);
                for (int i = 0; i < cnt; ++i) out[at + i] = b.s(raw);
            } else {
                for (int i = 0; i < cnt; ++i) {
                    const uint32_t q = b.unary();
                    const uint64_t v = ((uint64_t)q << k) | b.u(k);
                    out[at + i] = (int64_t)(v >> 1) ^ -(int64_t)(v & 1);
                }
            }
            at += cnt;
            if (b.over) return "frame runs past the end of the stream";
        }
        // prediction
        if (kind < 32) {
            switch (order) {
                case 0: break;
                case 1: for (int i = 1; i < n; ++i) out[i] += out[i - 1]; break;
                case 2: for (int i = 2; i < n; ++i) out[i] += 2 * out[i - 1] - out[i - 2]; break;
                case 3: for (int i = 3; i < n; ++i) out[i] += 3 * out[i - 1] - 3 * out[i - 2] + out[i - 3]; break;
                default: for (int i = 4; i < n; ++i) out[i] += 4 * out[i - 1] - 6 * out[i - 2] + 4 * out[i - 3] - out[i - 4]; break;
            }
        } else {
            for (int i = order; i < n; ++i) {
                int64_t acc = 0;
                for (int j = 0; j < order; ++j) acc += (int64_t)coef[j] * out[i - 1 - j];
                out[i] += acc >> shift;
            }
        }
    } else {
        return "reserved subframe type";
    }
    if (wasted)
        for (int i = 0; i < n; ++i) out[i] *= (int64_t)1 << wasted;
    return b.over ? "frame runs past the end of the stream" : nullptr;
}

// Decodes the whole stream.  Samples [first, first + count) of `channel` go to out (int32; may be null: count only);
// returns "" or the reason, *decoded = samples per channel found in the stream.  verify_md5: also checks the
// STREAMINFO signature when the stream carries one (needs every channel, costs ~1 ns per byte).
inline std::string decode(const unsigned char* d, int64_t size, int channel, int64_t first, int64_t count, int32_t* out,
                          bool verify_md5, Info& info, int64_t* decoded) {
    std::string err = parse_streaminfo(d, size, info);
    if (!err.empty()) return err;
    if (channel < 0 || channel >= info.channels) return "no such channel";
    static const int kBlock[16] = {0, 192, 576, 1152, 2304, 4608, 0, 0, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768};
    static const int kBps[8] = {0, 8, 12, -1, 16, 20, 24, 32};
    bool md5_on = verify_md5;
    if (md5_on) {
        bool any = false;
        for (int i = 0; i < 16; ++i) any |= info.md5[i] != 0;
        md5_on = any;
    }
    Md5 md5;
    std::vector<int64_t> buf;
    std::vector<unsigned char> inter;
    const unsigned char* const end = d + size;
    const unsigned char* p = d + info.audio_off;
    int64_t at = 0;                          // samples per channel decoded so far
    const int nbytes = (info.bits + 7) / 8;
    while (p < end) {
        if (end - p < 6) return "truncated frame header at byte " + std::to_string(p - d);
        if (p[0] != 0xFF || (p[1] & 0xFE) != 0xF8) return "lost frame sync at byte " + std::to_string(p - d);
        const int bs_code = p[2] >> 4, sr_code = p[2] & 15, ch_code = p[3] >> 4, bps_code = (p[3] >> 1) & 7;
        if (p[3] & 1) return "reserved header bit set";
        const unsigned char* q = p + 4;
        int extra = 0;                       // UTF-8-like coded frame / sample number: skip it
        while (extra < 8 && (q[0] & (0x80 >> extra))) ++extra;
        if (extra == 1 || extra == 8) return "bad coded frame number";
        const int numlen = extra ? extra : 1;
        if (end - q < numlen + 5) return "truncated frame header";
        for (int i = 1; i < numlen; ++i)
            if ((q[i] & 0xC0) != 0x80) return "bad coded frame number";
        q += numlen;
        int n;
        if (bs_code == 0) return "reserved block size code";
        if (bs_code == 6) n = q[0] + 1, q += 1;
        else if (bs_code == 7) n = (q[0] << 8 | q[1]) + 1, q += 2;
        else n = kBlock[bs_code];
        if (sr_code == 12) q += 1;
        else if (sr_code == 13 || sr_code == 14) q += 2;
        else if (sr_code == 15) return "invalid sample rate code";
        if (q >= end) return "truncated frame header";
        if (crc8(p, (size_t)(q - p)) != q[0]) return "frame header CRC-8 mismatch at byte " + std::to_string(p - d);
        ++q;
        const int bps = bps_code == 0 ? info.bits : kBps[bps_code];
        if (bps < 0) return "reserved sample size code";
        if (bps != info.bits) return "sample size changes inside the stream";
        int nch;
        if (ch_code < 8) nch = ch_code + 1;
        else if (ch_code <= 10) nch = 2;
        else return "reserved channel assignment";
        if (nch != info.channels) return "channel count changes inside the stream";
        buf.resize((size_t)n * nch);
        Bits b(q, end);
        for (int c = 0; c < nch; ++c) {
            const bool side = (ch_code == 8 && c == 1) || (ch_code == 9 && c == 0) || (ch_code == 10 && c == 1);
            const char* e = decode_subframe(b, n, bps + (side ? 1 : 0), buf.data() + (size_t)c * n);
            if (e) return std::string(e) + " (frame at byte " + std::to_string(p - d) + ")";
        }
        int64_t* c0 = buf.data();
        int64_t* c1 = buf.data() + n;
        if (ch_code == 8) {
            for (int i = 0; i < n; ++i) c1[i] = c0[i] - c1[i];
        } else if (ch_code == 9) {
            for (int i = 0; i < n; ++i) c0[i] += c1[i];
        } else if (ch_code == 10) {
            for (int i = 0; i < n; ++i) {
                const int64_t s = c1[i], m = c0[i] * 2 + (s & 1);
                c0[i] = (m + s) >> 1;
                c1[i] = (m - s) >> 1;
            }
        }
        b.align();
        const unsigned char* fe = b.byte_pos();
        if (b.over || fe + 2 > end) return "frame runs past the end of the stream";
        if (crc16(p, (size_t)(fe - p)) != (uint16_t)(fe[0] << 8 | fe[1])) return "frame CRC-16 mismatch at byte " + std::to_string(p - d);
        if (md5_on) {
            inter.resize((size_t)n * nch * nbytes);
            unsigned char* o = inter.data();
            for (int i = 0; i < n; ++i)
                for (int c = 0; c < nch; ++c) {
                    const int64_t v = buf[(size_t)c * n + i];
                    for (int k = 0; k < nbytes; ++k) *o++ = (unsigned char)(v >> (8 * k));
                }
            md5.update(inter.data(), inter.size());
        }
        if (out) {
            const int64_t lo = std::max<int64_t>(first, at), hi = std::min<int64_t>(first + count, at + n);
            const int64_t* src = buf.data() + (size_t)channel * n;
            for (int64_t i = lo; i < hi; ++i) out[i - first] = (int32_t)src[i - at];
        }
        at += n;
        p = fe + 2;
    }
    if (info.total && at != info.total)
        return "STREAMINFO announces " + std::to_string(info.total) + " samples, the frames hold " + std::to_string(at);
    if (md5_on) {
        unsigned char sig[16];
        md5.finish(sig);
        if (memcmp(sig, info.md5, 16) != 0) return "MD5 signature mismatch";
    }
    if (decoded) *decoded = at;
    return "";
}

// ---- frame index for the GPU decoder (oe_flac_gpu.cuh): where every audio frame starts and ends -------------------
// A frame does not announce its length (RFC 9639 section 9): a decoder learns it by decoding.  The host only needs the
// boundaries, so it walks the stream from header to header: a header is accepted when its sync code, reserved bits,
// blocking strategy, CRC-8 AND its coded frame / sample number (the predecessor's + 1 / + block size) all fit.  A false
// positive inside compressed data needs ~2^-40 luck per byte; the GPU decoder still catches it, because each frame must
// end exactly where the next one starts and carry a matching CRC-16.
struct FrameHeader {
    int block = 0, hdr_bytes = 0, ch_code = 0, bps = 0;
    bool variable = false;
    int64_t number = 0;                      // frame number (fixed block size) or first sample number (variable)
};

// Parses the header at p; returns false when p does not hold a valid header (no message: the scan probes candidates).
inline bool parse_frame_header(const unsigned char* p, const unsigned char* end, const Info& info, FrameHeader& h) {
    static const int kBlock[16] = {0, 192, 576, 1152, 2304, 4608, 0, 0, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768};
    static const int kBps[8] = {0, 8, 12, -1, 16, 20, 24, 32};
    if (end - p < 6 || p[0] != 0xFF || (p[1] & 0xFE) != 0xF8 || (p[3] & 1)) return false;
    h.variable = p[1] & 1;
    const int bs_code = p[2] >> 4, sr_code = p[2] & 15, bps_code = (p[3] >> 1) & 7;
    h.ch_code = p[3] >> 4;
    if (bs_code == 0 || sr_code == 15 || h.ch_code > 10) return false;
    const unsigned char* q = p + 4;
    int extra = 0;
    while (extra < 8 && (q[0] & (0x80 >> extra))) ++extra;
    if (extra == 1 || extra == 8) return false;
    const int numlen = extra ? extra : 1;
    if (end - q < numlen + 5) return false;
    int64_t num = extra ? (q[0] & (0x7F >> extra)) : q[0];
    for (int i = 1; i < numlen; ++i) {
        if ((q[i] & 0xC0) != 0x80) return false;
        num = num << 6 | (q[i] & 0x3F);
    }
    h.number = num;
    q += numlen;
    if (bs_code == 6) h.block = q[0] + 1, q += 1;
    else if (bs_code == 7) h.block = (q[0] << 8 | q[1]) + 1, q += 2;
    else h.block = kBlock[bs_code];
    if (sr_code == 12) q += 1;
    else if (sr_code == 13 || sr_code == 14) q += 2;
    if (q >= end || crc8(p, (size_t)(q - p)) != q[0]) return false;
    h.hdr_bytes = (int)(q + 1 - p);
    h.bps = bps_code == 0 ? info.bits : kBps[bps_code];
    return h.bps > 0;
}

struct FrameSpan {
    int64_t off = 0, bytes = 0, first_sample = 0;   // byte offset in the stream, frame length incl. CRC-16 (<= 0: unknown,
    FrameHeader h;                                  // -bytes = what is left of the stream), first sample it holds
};

// Walks the whole stream; returns "" or the reason.  The last frame's length is only bounded by the end of the stream
// (trailing tags are legal), so it is reported as -(bytes left).
inline std::string scan_frames(const unsigned char* d, int64_t size, const Info& info, std::vector<FrameSpan>& spans) {
    spans.clear();
    const unsigned char* const end = d + size;
    const unsigned char* p = d + info.audio_off;
    int64_t at = 0, index = 0;
    FrameHeader h;
    if (p >= end) return "";
    if (!parse_frame_header(p, end, info, h)) return "lost frame sync at byte " + std::to_string(p - d);
    for (;;) {
        FrameSpan s;
        s.off = p - d;
        s.first_sample = at;
        s.h = h;
        // next header: first candidate behind this frame's smallest possible body that parses and continues the numbering
        const unsigned char* q = p + h.hdr_bytes + 2;
        FrameHeader nh;
        bool found = false;
        while (q < end) {
            q = static_cast<const unsigned char*>(memchr(q, 0xFF, (size_t)(end - q)));
            if (!q) break;
            if (parse_frame_header(q, end, info, nh) && nh.variable == h.variable &&
                nh.number == h.number + (h.variable ? h.block : 1)) {
                found = true;
                break;
            }
            ++q;
        }
        at += h.block;
        ++index;
        if (!found) {
            s.bytes = -(int64_t)(end - p);
            spans.push_back(s);
            break;
        }
        s.bytes = q - p;
        spans.push_back(s);
        p = q;
        h = nh;
    }
    return "";
}

// ---- encoder (fixed predictors, partitioned Rice coding; mono / 16 bit) ---------------------------------------------
// Writes the streams the bench's FLAC legs and the tests read -- there is no FLAC encoder in the image.  Per frame: the
// fixed predictor (order 0-4) with the smallest sum of |residual|, partition order `porder` (lowered until it divides the
// block), per partition the Rice parameter floor(log2(mean folded residual)); STREAMINFO carries the MD5 signature.
class BitWriter {
  public:
    std::vector<unsigned char>& out;
    uint64_t acc = 0;
    int fill = 0;
    explicit BitWriter(std::vector<unsigned char>& o) : out(o) {}
    inline void u(uint64_t v, int n) {       // n <= 32
        if (n == 0) return;
        acc = acc << n | (v & ((n == 64) ? ~0ull : ((1ull << n) - 1)));
        fill += n;
        while (fill >= 8) {
            out.push_back((unsigned char)(acc >> (fill - 8)));
            fill -= 8;
        }
    }
    inline void unary(uint32_t q) {
        while (q >= 32) u(0, 32), q -= 32;
        u(1, (int)q + 1);
    }
    inline void align() {
        if (fill) u(0, 8 - fill);
    }
};

inline void encode(const int16_t* pcm, int64_t n, int sample_rate, int block, int porder, std::vector<unsigned char>& out) {
    out.clear();
    const unsigned char magic[4] = {'f', 'L', 'a', 'C'};
    out.insert(out.end(), magic, magic + 4);
    out.push_back(0x80);
    out.push_back(0), out.push_back(0), out.push_back(34);
    {
        BitWriter w(out);
        const uint64_t bs = (uint64_t)(n > block ? block : std::max<int64_t>(16, n));   // both bounds exclude a shorter last block
        w.u(bs, 16);
        w.u(bs, 16);
        w.u(0, 24), w.u(0, 24);
        w.u((uint64_t)sample_rate, 20);
        w.u(0, 3);
        w.u(15, 5);
        w.u((uint64_t)(n >> 32), 4), w.u((uint64_t)(n & 0xFFFFFFFFu), 32);
        Md5 md5;
        md5.update(reinterpret_cast<const unsigned char*>(pcm), (size_t)n * 2);     // little-endian host
        unsigned char sig[16];
        md5.finish(sig);
        for (int i = 0; i < 16; ++i) w.u(sig[i], 8);
    }
    std::vector<int32_t> res((size_t)block);
    std::vector<unsigned char> frame;
    int64_t index = 0;
    for (int64_t at = 0; at < n; at += block, ++index) {
        const int m = (int)std::min<int64_t>(block, n - at);
        const int16_t* s = pcm + at;
        frame.clear();
        BitWriter w(frame);
        w.u(0xFFF8, 16);
        w.u(m <= 256 ? 6 : 7, 4);
        w.u(0, 4);
        w.u(0, 4);
        w.u(4, 3);
        w.u(0, 1);
        {                                                                            // coded frame number
            const uint64_t v = (uint64_t)index;
            if (v < 0x80) w.u(v, 8);
            else {
                int len = 2;
                while (len < 7 && v >= (1ull << (5 * len + 1))) ++len;
                w.u((0xFF00u >> len & 0xFF) | (unsigned)(v >> (6 * (len - 1))), 8);
                for (int i = len - 2; i >= 0; --i) w.u(0x80 | ((v >> (6 * i)) & 0x3F), 8);
            }
        }
        if (m <= 256) w.u((uint64_t)m - 1, 8);
        else w.u((uint64_t)m - 1, 16);
        w.u(crc8(frame.data(), frame.size()), 8);
        // predictor choice
        int64_t cost[5] = {0, 0, 0, 0, 0};
        for (int i = 4; i < m; ++i) {
            const int64_t a = s[i], b = s[i - 1], c = s[i - 2], d = s[i - 3], e = s[i - 4];
            cost[0] += std::llabs(a);
            cost[1] += std::llabs(a - b);
            cost[2] += std::llabs(a - 2 * b + c);
            cost[3] += std::llabs(a - 3 * b + 3 * c - d);
            cost[4] += std::llabs(a - 4 * b + 6 * c - 4 * d + e);
        }
        int order = 0;
        for (int o = 1; o <= 4; ++o)
            if (cost[o] < cost[order]) order = o;
        if (order >= m) order = 0;
        bool constant = true;
        for (int i = 1; i < m && constant; ++i) constant = s[i] == s[0];
        if (constant) {
            w.u(0, 8);
            w.u((uint64_t)(uint16_t)s[0], 16);
        } else {
            w.u((uint64_t)(8 + order) << 1, 8);
            for (int i = 0; i < order; ++i) w.u((uint64_t)(uint16_t)s[i], 16);
            for (int i = order; i < m; ++i) {
                const int64_t a = s[i], b = i >= 1 ? s[i - 1] : 0, c = i >= 2 ? s[i - 2] : 0, d = i >= 3 ? s[i - 3] : 0, e = i >= 4 ? s[i - 4] : 0;
                int64_t r = a;
                if (order == 1) r = a - b;
                else if (order == 2) r = a - 2 * b + c;
                else if (order == 3) r = a - 3 * b + 3 * c - d;
                else if (order == 4) r = a - 4 * b + 6 * c - 4 * d + e;
                res[(size_t)i] = (int32_t)r;
            }
            int po = porder;
            while (po && (((m >> po) << po) != m || (m >> po) <= order)) --po;
            w.u(1, 2);                                                               // 5-bit Rice parameters
            w.u((uint64_t)po, 4);
            int at_r = order;
            for (int part = 0; part < (1 << po); ++part) {
                const int cnt = (m >> po) - (part == 0 ? order : 0);
                uint64_t sum = 0;
                for (int i = 0; i < cnt; ++i) {
                    const int32_t r = res[(size_t)(at_r + i)];
                    sum += ((uint32_t)r << 1) ^ (uint32_t)(r >> 31);
                }
                int k = 0;
                if (cnt > 0) {
                    const uint64_t mean = sum / (uint64_t)cnt;
                    while (k < 30 && (2ull << k) <= mean) ++k;
                }
                w.u((uint64_t)k, 5);
                for (int i = 0; i < cnt; ++i) {
                    const int32_t r = res[(size_t)(at_r + i)];
                    const uint32_t z = ((uint32_t)r << 1) ^ (uint32_t)(r >> 31);
                    w.unary(z >> k);
                    w.u(z & ((1u << k) - 1), k);
                }
                at_r += cnt;
            }
        }
        w.align();
        const uint16_t c16 = crc16(frame.data(), frame.size());
        frame.push_back((unsigned char)(c16 >> 8));
        frame.push_back((unsigned char)(c16 & 0xFF));
        out.insert(out.end(), frame.begin(), frame.end());
    }
}

}  // namespace oe_flac
