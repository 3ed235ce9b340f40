// FLAC frames decoded on the GPU: the compressed corpus crosses PCIe (about half the bytes of its PCM for speech), one
// thread decodes one frame straight into the packed int16 buffer the fbank kernels read.  Replaces the per-utterance
// libsox / libFLAC decode behind torchaudio.load (openeat/dataset/dataset.py:62-72) for the LibriSpeech-style lists;
// the format is RFC 9639 (section 9: frames, subframes, partitioned Rice residuals), the arithmetic is integer and the
// result is bit-exact with the host decoder in oe_flac.h.
//
// Why one thread per frame: frames are the format's only independent units (each starts with its own header and warm-up
// samples); inside a frame both the entropy code (variable-length, no resynchronisation points) and the predictor (an
// IIR recursion) are serial.  A batch of 256 utterances holds ~6 000 frames = ~190 warps, one or two per SM scheduler,
// each a latency-bound serial chain of ~60 instructions per sample; HBM is not the limit (2 bytes out per sample), the
// dependent chain clz -> shift -> extract -> predict is.  Every thread runs the SAME code whatever its frame's predictor
// is: fixed predictors are rewritten as linear predictors, every order runs the 12-tap body with zero coefficients
// (orders above 12 exist only outside the format's streamable subset and are sent back to the host decoder), so warps
// only diverge at partition boundaries and in the rare escape / verbatim / constant subframes.
//
// The host (oe_flac_pack) finds the frame boundaries by walking header to header; the kernel checks what the host
// could not: each frame must end exactly where the next one starts and its CRC-16 must match.
#pragma once

#include <cstdint>
#include <cstring>

#include "../../include/openeat_frontend.h"

namespace oe_flacgpu {

constexpr int kMaxOrder = 12;

constexpr int kRingWords = 64;               // per-lane staging ring in shared memory: two 128-byte lines of the lane's stream

#ifdef __CUDACC__
__device__ __forceinline__ void line_copy_async(void* smem_dst, const void* gmem_src, int chunks) {   // chunks x 16 bytes (<= 8), 16-byte aligned
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j < chunks)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16 * j), "l"(static_cast<const char*>(gmem_src) + 16 * j) : "memory");
}
__device__ __forceinline__ void line_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void line_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void line_wait_all_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ int64_t mad_wide(int32_t a, int32_t b, int64_t c) {
    int64_t r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint32_t big_endian(uint32_t w) { return __byte_perm(w, 0, 0x0123); }
__device__ __forceinline__ int clz32(uint32_t v) { return __clz((int)v); }
#else   // host emulation (oe_emul.cpp): same arithmetic, copies are immediate
inline void line_copy_async(void* d, const void* s, int chunks) { memcpy(d, s, 16 * (size_t)(chunks < 0 ? 0 : chunks > 8 ? 8 : chunks)); }
inline void line_commit() {}
inline void line_wait_all() {}
inline void line_wait_all_but_one() {}
inline int64_t mad_wide(int32_t a, int32_t b, int64_t c) { return (int64_t)a * b + c; }
inline uint32_t big_endian(uint32_t w) { return __builtin_bswap32(w); }
inline int clz32(uint32_t v) { return v ? __builtin_clz(v) : 32; }
#endif

// Bit reader of one lane.  The lane's stream is staged through a private 256-byte ring in shared memory, refilled a
// 128-byte line at a time with cp.async one line AHEAD of use (issued ~100 samples before its first word is read, so no
// lane ever waits for L2 / HBM: the 32 lanes of a warp walk 32 different streams and every line is a first touch for
// somebody).  The bit window is a left-aligned 64-bit register with at least 32 valid bits; it is topped up from the ring
// without a branch (the next word is loaded speculatively, merged under a predicate), because a warp whose lanes branch
// independently executes every side of every branch in every iteration.  A first version that loaded words on demand with
// branches measured 1.8 ms for 6 000 frames (~850 cycles per sample: ~200 dependent instructions at one warp per scheduler).
struct Reader {
    const unsigned char* src16;              // 16-byte aligned address that holds the stream's first byte
    uint32_t* ring;                          // this lane's kRingWords words of shared memory
    int64_t src_bytes;                       // bytes readable from src16 (multiple of 16)
    uint64_t win;                            // next bits, left aligned; bits below `have` are zero
    int have;
    int wpos;                                // next stream word (from src16) to merge into the window
    int next_issue;                          // when wpos reaches it, the ring half behind is refilled
    int start_bits;
    bool over;

    __device__ __forceinline__ void issue_line(int line) {    // stream bytes [128 line, 128 line + 128) -> ring half (line & 1)
        const int64_t at = (int64_t)line * 128;
        const int64_t left = (src_bytes - at) >> 4;           // whole 16-byte chunks that exist (the buffer ends in a partial line)
        line_copy_async(ring + (line & 1) * 32, src16 + at, (int)(left < 0 ? 0 : left > 8 ? 8 : left));
        line_commit();
    }
    __device__ __forceinline__ void service() {
        if (wpos >= next_issue) {            // the half behind wpos is consumed: bring in the line after the one in use
            issue_line((next_issue >> 5) + 1);
            line_wait_all_but_one();         // the line requested one round ago (in use from now on) has landed
            next_issue += 32;
        }
    }
    __device__ __forceinline__ uint32_t word() const { return big_endian(ring[wpos & (kRingWords - 1)]); }
    __device__ __forceinline__ void top_up(uint32_t w) {      // branch-free: merge `w` (= word()) when 32 bits or fewer are left
        const bool need = have <= 32;
        const uint64_t add = (uint64_t)w << ((32 - have) & 63);
        win |= need ? add : 0ull;
        have += need ? 32 : 0;
        if (need && (int64_t)wpos * 4 >= src_bytes) over = true;
        wpos += need ? 1 : 0;
    }
    __device__ __forceinline__ void init(const unsigned char* base, int64_t byte_off, int64_t limit, uint32_t* lane_ring) {
        const unsigned char* p = base + byte_off;
        const int mis = (int)(reinterpret_cast<uintptr_t>(p) & 15);
        src16 = p - mis;
        src_bytes = (base + limit) - src16;
        ring = lane_ring;
        issue_line(0);
        issue_line(1);
        line_wait_all();
        next_issue = 32;
        wpos = mis >> 2;
        start_bits = (mis & ~3) * 8 + (mis & 3) * 8;
        over = false;
        win = (uint64_t)word() << 32;
        ++wpos;
        win |= word();
        ++wpos;
        have = 64 - (mis & 3) * 8;
        win <<= (mis & 3) * 8;
        top_up(word());
    }
    // bits consumed since init
    __device__ __forceinline__ int64_t consumed() const { return (int64_t)wpos * 32 - have - start_bits; }
    __device__ __forceinline__ uint32_t take(int n) {          // 0 <= n <= 32
        service();
        const uint32_t w = word();
        const uint32_t v = (uint32_t)((win >> 1) >> (63 - n));
        win <<= n;
        have -= n;
        top_up(w);
        return v;
    }
    __device__ __forceinline__ int32_t take_signed(int n) {    // 1 <= n <= 32
        const uint32_t v = take(n);
        return (int32_t)(v << (32 - n)) >> (32 - n);
    }
    __device__ __forceinline__ uint32_t unary() {              // zeros before the next one bit (any length)
        uint32_t q = 0;
        for (;;) {
            const int z = clz32((uint32_t)(win >> 32));         // at least 32 valid bits: a one among them ends the run
            if (z < 32) {
                q += (uint32_t)z;
                take(z + 1);
                return q;
            }
            q += 32;
            take(32);
            if (over || q > (1u << 20)) {
                over = true;
                return q;
            }
        }
    }
    // One Rice symbol with parameter k (<= 30): quotient in unary, then k bits.  When quotient + 1 + k <= 32 (almost always)
    // the symbol is cut out of the window's top half in one step.
    __device__ __forceinline__ uint32_t rice(int k) {
        service();
        const uint32_t w = word();
        const uint32_t hi = (uint32_t)(win >> 32);
        const int z = clz32(hi);
        const int n = z + 1 + k;
        if (n <= 32) {
            const uint32_t t = (hi << z) << 1;
            const uint32_t rem = (t >> 1) >> (31 - k);
            win <<= n;
            have -= n;
            top_up(w);
            return ((uint32_t)z << k) | rem;
        }
        const uint32_t q = unary();
        return (q << k) | take(k);
    }
};

// error bits per utterance (oe_flac_decode_batch's d_errors)
constexpr int kErrEnd = 1;                   // the frame did not end where the next one starts
constexpr int kErrCrc = 2;                   // CRC-16 mismatch
constexpr int kErrHost = 4;                  // legal FLAC this kernel does not decode (predictor order > 12): use the host decoder
constexpr int kErrFormat = 8;                // reserved codes / values the format forbids
constexpr int kErrOverrun = 16;              // the bit stream ran past the end of the buffer

__global__ void __launch_bounds__(32) oe_flac_decode_kernel(const unsigned char* __restrict__ comp, int64_t comp_limit,
                                                           const oe_flac_frame* __restrict__ frames, int64_t n_frames,
                                                           int16_t* __restrict__ pcm, int32_t* __restrict__ errors, int verify_crc) {
    __shared__ uint16_t t16[256];
    __shared__ __align__(16) uint32_t rings[32][kRingWords];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint16_t w = (uint16_t)(i << 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) w = (uint16_t)((w & 0x8000) ? (w << 1) ^ 0x8005 : w << 1);
        t16[i] = w;
    }
    __syncthreads();
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const oe_flac_frame fr = frames[f];
    Reader r;
    r.init(comp, fr.comp_off + fr.hdr_bytes, comp_limit, rings[threadIdx.x]);
    int err = 0;
    const int n = fr.block;
    int bps = fr.bps;
    const int lo = fr.skip, hi = fr.skip + fr.take;
    int16_t* const out = pcm + (fr.out_off - fr.skip);

    const uint32_t head = r.take(8);
    if (head & 0x80) err |= kErrFormat;
    const int kind = (head >> 1) & 63;
    int wasted = 0;
    if (head & 1) {
        wasted = (int)r.unary() + 1;
        bps -= wasted;
        if (bps < 1) err |= kErrFormat, bps = 1;
    }
#define OE_EMIT(i, v)                                                  \
    do {                                                               \
        if ((i) >= lo && (i) < hi) out[i] = (int16_t)((v) << wasted);  \
    } while (0)

    if (kind == 0) {
        const int32_t v = r.take_signed(bps);
        for (int i = lo; i < hi; ++i) out[i] = (int16_t)(v << wasted);
    } else if (kind == 1) {
        for (int i = 0; i < n; ++i) {
            const int32_t v = r.take_signed(bps);
            OE_EMIT(i, v);
        }
    } else if ((kind >= 8 && kind <= 12) || kind >= 32) {
        const int order = kind >= 32 ? kind - 31 : kind - 8;
        if (order > kMaxOrder) err |= kErrHost;
        else if (order > n) err |= kErrFormat;
        else {
            int32_t c[kMaxOrder], h[kMaxOrder];                 // h[0] = most recent sample
#pragma unroll
            for (int j = 0; j < kMaxOrder; ++j) c[j] = 0, h[j] = 0;
            for (int i = 0; i < order; ++i) {
                const int32_t v = r.take_signed(bps);
                OE_EMIT(i, v);
#pragma unroll
                for (int j = kMaxOrder - 1; j > 0; --j) h[j] = h[j - 1];
                h[0] = v;
            }
            int shift = 0;
            if (kind >= 32) {
                const int prec = (int)r.take(4) + 1;
                if (prec == 16) err |= kErrFormat;
                shift = r.take_signed(5);
                if (shift < 0) err |= kErrFormat, shift = 0;
#pragma unroll
                for (int j = 0; j < kMaxOrder; ++j)
                    if (j < order) c[j] = r.take_signed(prec);
            } else {
                c[0] = order == 1 ? 1 : order == 2 ? 2 : order == 3 ? 3 : order == 4 ? 4 : 0;
                c[1] = order == 2 ? -1 : order == 3 ? -3 : order == 4 ? -6 : 0;
                c[2] = order == 3 ? 1 : order == 4 ? 4 : 0;
                c[3] = order == 4 ? -1 : 0;
            }
            const int method = (int)r.take(2);
            if (method > 1) err |= kErrFormat;
            const int pbits = 4 + (method & 1);
            const int porder = (int)r.take(4);
            const int psize = n >> porder;
            if ((porder && (psize << porder) != n) || psize < order) err |= kErrFormat;
            if (!err) {
                int k = (int)r.take(pbits);
                int raw = -1;                                   // >= 0: escape partition with `raw` bits per residual
                if (k == (1 << pbits) - 1) raw = (int)r.take(5);
                int part_end = psize;
                // the taps that do not wait for the newest sample are summed one iteration ahead: the recursion's chain per
                // sample is one multiply-add, one shift and one add, next to (not behind) the entropy decoder's chain
                int64_t ahead = 0;
#pragma unroll
                for (int j = 1; j < kMaxOrder; ++j) ahead = mad_wide(c[j], h[j], ahead);
                for (int i = order; i < n; ++i) {
                    if (i == part_end) {
                        k = (int)r.take(pbits);
                        raw = -1;
                        if (k == (1 << pbits) - 1) raw = (int)r.take(5);
                        part_end += psize;
                    }
                    int32_t res;
                    if (raw < 0) {
                        const uint32_t v = r.rice(k);
                        res = (int32_t)(v >> 1) ^ -(int32_t)(v & 1);
                    } else {
                        res = raw ? r.take_signed(raw) : 0;
                    }
                    const int32_t s = res + (int32_t)(mad_wide(c[0], h[0], ahead) >> shift);
                    OE_EMIT(i, s);
#pragma unroll
                    for (int j = kMaxOrder - 1; j > 0; --j) h[j] = h[j - 1];
                    h[0] = s;
                    int64_t a0 = 0, a1 = 0;
#pragma unroll
                    for (int j = 1; j < kMaxOrder; j += 2) a0 = mad_wide(c[j], h[j], a0);
#pragma unroll
                    for (int j = 2; j < kMaxOrder; j += 2) a1 = mad_wide(c[j], h[j], a1);
                    ahead = a0 + a1;
                    if (r.over) break;
                }
            }
        }
    } else {
        err |= kErrFormat;
    }
#undef OE_EMIT
    if (r.over) err |= kErrOverrun;
    if (!err) {
        // the frame ends on the next byte boundary, followed by its CRC-16; the next frame's sync code comes right behind
        const int64_t body = (r.consumed() + 7) >> 3;
        const int64_t total = fr.hdr_bytes + body + 2;
        if (fr.frame_bytes > 0 ? total != fr.frame_bytes : total > -fr.frame_bytes) err |= kErrEnd;
        else if (verify_crc) {
            const unsigned char* p = comp + fr.comp_off;
            const unsigned char* const e = p + total - 2;
            uint32_t crc = 0;
            while (p < e && (reinterpret_cast<uintptr_t>(p) & 3)) crc = ((crc << 8) ^ t16[((crc >> 8) ^ *p++) & 0xFF]) & 0xFFFF;
            for (; p + 4 <= e; p += 4) {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
                crc = ((crc << 8) ^ t16[((crc >> 8) ^ w) & 0xFF]) & 0xFFFF;
                crc = ((crc << 8) ^ t16[((crc >> 8) ^ (w >> 8)) & 0xFF]) & 0xFFFF;
                crc = ((crc << 8) ^ t16[((crc >> 8) ^ (w >> 16)) & 0xFF]) & 0xFFFF;
                crc = ((crc << 8) ^ t16[((crc >> 8) ^ (w >> 24)) & 0xFF]) & 0xFFFF;
            }
            while (p < e) crc = ((crc << 8) ^ t16[((crc >> 8) ^ *p++) & 0xFF]) & 0xFFFF;
            if (crc != (uint32_t)(e[0] << 8 | e[1])) err |= kErrCrc;
        }
    }
    if (err) atomicOr(errors + fr.utt, err);
}

}  // namespace oe_flacgpu
