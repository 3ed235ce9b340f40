// FLAC frames decoded on the GPU: the compressed corpus crosses PCIe (about half the bytes of its PCM for speech) and is
// decoded straight into the packed int16 buffer the fbank kernels read.  Replaces the per-utterance libsox / libFLAC decode
// behind torchaudio.load (openeat/dataset/dataset.py:62-72) for the LibriSpeech-style lists; the format is RFC 9639
// (section 9: frames, subframes, partitioned Rice residuals), the arithmetic is integer and the result is bit-exact with
// the host decoder in oe_flac.h.
//
// Parallelism: frames are the format's only independent units (each starts with its own header and warm-up samples);
// inside a frame both the entropy code (variable-length, no resynchronisation points) and the predictor (an IIR
// recursion) are serial.  A batch of 256 utterances holds ~6 000 frames, so this is a latency problem -- a few hundred
// warps on 592 schedulers, HBM is not the limit (2 bytes out per sample) -- and the design cuts the dependent chain per
// sample: a block takes 32 frames; lane l of warp 0 turns frame l's bit stream into residuals, lane l of warp 1 runs the
// predictor and stores PCM, lane l of warp 2 checks the CRC-16 meanwhile (see "three warps per block" below).  Every lane
// runs the SAME code whatever its frame's predictor is: fixed predictors are rewritten as linear predictors, every order
// runs the 12-tap body with zero coefficients (orders above 12 exist only outside the format's streamable subset and are
// sent back to the host decoder), constant / verbatim subframes travel as residuals of a zero predictor.
//
// The host (oe_flac_pack) finds the frame boundaries by walking header to header; the kernel checks what the host
// could not: each frame must end exactly where the next one starts and its CRC-16 must match.
//
// The reader / predictor / CRC code below also compiles for the host (no __CUDACC__: oe_emul.cpp), where the CPU test
// suite runs it frame by frame over every stream of the test matrix.
#pragma once

#include <cstdint>
#include <cstring>

#include "../../include/openeat_frontend.h"

namespace oe_flacgpu {

constexpr int kMaxOrder = 12;

constexpr int kRingWords = 128;              // per-lane staging ring in shared memory: four 128-byte lines of the lane's stream

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
// high / low word of a 64-bit pair shifted by 0..32 (the shift clamps at 32, unlike C's << and >>)
__device__ __forceinline__ uint32_t shl_pair_hi(uint32_t hi, uint32_t lo, int n) { return __funnelshift_lc(lo, hi, n); }
__device__ __forceinline__ uint32_t shr_pair_lo(uint32_t hi, uint32_t lo, int n) { return __funnelshift_rc(lo, hi, n); }
#else   // host emulation (oe_emul.cpp): same arithmetic, copies are immediate
inline void line_copy_async(void* d, const void* s, int chunks) { memcpy(d, s, 16 * (size_t)(chunks < 0 ? 0 : chunks > 8 ? 8 : chunks)); }
inline void line_commit() {}
inline void line_wait_all() {}
inline void line_wait_all_but_one() {}
inline int64_t mad_wide(int32_t a, int32_t b, int64_t c) { return (int64_t)a * b + c; }
inline uint32_t big_endian(uint32_t w) { return __builtin_bswap32(w); }
inline int clz32(uint32_t v) { return v ? __builtin_clz(v) : 32; }
inline uint32_t shl_pair_hi(uint32_t hi, uint32_t lo, int n) { return (uint32_t)(((((uint64_t)hi << 32) | lo) << (n > 32 ? 32 : n)) >> 32); }
inline uint32_t shr_pair_lo(uint32_t hi, uint32_t lo, int n) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (n > 32 ? 32 : n)); }
#endif

// Bit reader of one lane.  The lane's stream is staged through a private 256-byte ring in shared memory, refilled a
// 128-byte line at a time with cp.async one line AHEAD of use (issued ~100 samples before its first word is read, so no
// lane ever waits for L2 / HBM: the 32 lanes of a warp walk 32 different streams and every line is a first touch for
// somebody).  The bit window is a left-aligned register pair (hi:lo) with at least 32 valid bits; it is topped up from the ring
// without a branch (the next word is loaded speculatively, merged under a predicate), because a warp whose lanes branch
// independently executes every side of every branch in every iteration.  A first version that loaded words on demand with
// branches measured 1.8 ms for 6 000 frames (~850 cycles per sample: ~200 dependent instructions at one warp per scheduler).
struct Reader {
    const unsigned char* src16;              // 16-byte aligned address that holds the stream's first byte
    uint32_t* ring;                          // this lane's kRingWords words of shared memory
    int64_t src_bytes;                       // bytes readable from src16 (multiple of 16)
    uint32_t hi, lo;                         // next bits, left aligned in the pair hi:lo; bits below `have` are zero
    int have;
    int wpos;                                // next stream word (from src16) to merge into the window
    int issued;                              // lines requested so far: line L lives in ring slot L & 3 until wpos passes its last word
    int start_bits;
    bool over;

    __device__ __forceinline__ void issue_line(int line) {    // stream bytes [128 line, 128 line + 128) -> ring slot (line & 3)
        const int64_t at = (int64_t)line * 128;
        const int64_t left = (src_bytes - at) >> 4;           // whole 16-byte chunks that exist (the buffer ends in a partial line)
        line_copy_async(ring + (line & 3) * 32, src16 + at, (int)(left < 0 ? 0 : left > 8 ? 8 : left));
        line_commit();
    }
    // Keeps the ring as full as it can be: afterwards at least 96 words beyond wpos are staged or in flight, and everything
    // but the newest request has landed, i.e. the next 64 words are readable.  Called once per round of the sample loop (a
    // round's fast symbols consume at most 32 words) and by every slow-path read.
    __device__ __forceinline__ void service() {
        if ((issued - 3) * 32 <= wpos) {
            do issue_line(issued++);
            while ((issued - 3) * 32 <= wpos);
            line_wait_all_but_one();
            if ((int64_t)wpos * 4 >= src_bytes) over = true;
        }
    }
    __device__ __forceinline__ uint32_t word() const { return big_endian(ring[wpos & (kRingWords - 1)]); }
    // Branch-free: merges `w` (= word()) behind the valid bits when 32 or fewer are left (then lo is empty: all of them sit in
    // hi).  Two funnel shifts and two predicated moves.
    __device__ __forceinline__ void top_up(uint32_t w) {
        const bool need = have <= 32;
        const uint32_t a = shr_pair_lo(0u, w, have);           // w >> have          (0 when have == 32)
        const uint32_t b = shr_pair_lo(w, 0u, have);           // w << (32 - have)   (0 when have == 0)
        hi |= need ? a : 0u;
        lo = need ? b : lo;
        have += need ? 32 : 0;
        wpos += need ? 1 : 0;
    }
    __device__ __forceinline__ void drop(int n) {             // 0 <= n <= 32 bits consumed
        hi = shl_pair_hi(hi, lo, n);
        lo = shl_pair_hi(lo, 0u, n);
        have -= n;
    }
    __device__ __forceinline__ void init(const unsigned char* base, int64_t byte_off, int64_t limit, uint32_t* lane_ring) {
        const unsigned char* p = base + byte_off;
        const int mis = (int)(reinterpret_cast<uintptr_t>(p) & 15);
        src16 = p - mis;
        src_bytes = (base + limit) - src16;
        ring = lane_ring;
        for (issued = 0; issued < 4; ++issued) issue_line(issued);
        line_wait_all();
        wpos = mis >> 2;
        start_bits = (mis & ~3) * 8 + (mis & 3) * 8;
        over = false;
        hi = word();
        ++wpos;
        lo = word();
        ++wpos;
        have = 64;
        drop((mis & 3) * 8);
        top_up(word());
    }
    // bits consumed since init
    __device__ __forceinline__ int64_t consumed() const { return (int64_t)wpos * 32 - have - start_bits; }
    __device__ __forceinline__ uint32_t take(int n) {          // 0 <= n <= 32
        service();
        const uint32_t w = word();
        const uint32_t v = shr_pair_lo(0u, hi, 32 - n);          // the top n bits (n == 0: nothing)
        drop(n);
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
            const int z = clz32(hi);                            // at least 32 valid bits: a one among them ends the run
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
    // the symbol is cut out of the window's top half in one step (rice_fast; the caller has checked n <= 32).
    __device__ __forceinline__ uint32_t rice_fast(int k, int z, int n, uint32_t w) {
        const uint32_t rem = (hi >> (32 - n)) & ((1u << k) - 1u);    // the k bits behind the quotient's terminating one
        drop(n);
        top_up(w);
        return ((uint32_t)z << k) | rem;
    }
    __device__ __forceinline__ uint32_t rice(int k) {
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

// ---- three warps per block of 32 frames ---------------------------------------------------------------------------
// A frame's decode is two serial chains -- the entropy decoder (bit window -> residual) and the predictor (residual ->
// sample, an IIR recursion) -- plus a CRC over its bytes.  One thread running all three issued ~126 dependent instructions
// per sample at ~3.8 cycles each (one warp per scheduler, nothing to hide a dependent issue behind): 0.98 ms for 6 000
// frames.  Here lane l of warp 0 (the reader) turns frame l's bit stream into residuals, lane l of warp 1 (the predictor)
// turns them into samples, lane l of warp 2 checks the frame's CRC-16 meanwhile; the three warps sit on three schedulers of
// the SM.  Reader and predictor walk the sample index in lockstep, 32 samples per round, through a double-buffered
// [2][48 samples][32 lanes] residual tile in shared memory with ONE named barrier per round.  Constant and verbatim
// subframes, warm-up samples and escape partitions all travel as "residuals" of a predictor with zero coefficients, so the
// predictor warp runs one uniform loop.
constexpr int kChunk = 48;                   // samples per round (a multiple of kMaxOrder, see predictor_chunk)

struct Hand {                                // reader -> predictor, per lane, written before the first barrier
    int32_t c[kMaxOrder];
    int32_t order, shift, wasted, err;
};

struct ReaderState {
    Reader r;
    int n, order, k, raw, cval, pbits, psize, part_end, part_base, fast_until, err;
    bool rice;                               // false: every sample is `raw` bits (raw == 0: the constant cval)
};

// Parses the subframe header, the warm-up samples (into tile chunk 0) and the predictor; leaves the reader at the first
// residual.  tile: this lane's column of chunk 0, stride 32 words.
__device__ __forceinline__ void reader_prologue(ReaderState& st, Hand& hand, int32_t* tile0, int bps) {
    Reader& r = st.r;
    int err = 0;
    const int n = st.n;
    const uint32_t head = r.take(8);
    if (head & 0x80) err |= kErrFormat;
    const int kind = (head >> 1) & 63;
    int wasted = 0;
    if (head & 1) {
        wasted = (int)r.unary() + 1;
        bps -= wasted;
        if (bps < 1) err |= kErrFormat, bps = 1;
    }
    int order = 0, shift = 0;
#pragma unroll
    for (int j = 0; j < kMaxOrder; ++j) hand.c[j] = 0;
    st.rice = false;
    st.raw = 0;
    st.cval = 0;
    st.k = 0;
    st.pbits = 4;
    st.psize = n;
    st.part_end = n;
    st.part_base = 0;
    st.fast_until = 0;
    if (kind == 0) {
        st.cval = r.take_signed(bps);
    } else if (kind == 1) {
        st.raw = bps;
    } else if ((kind >= 8 && kind <= 12) || kind >= 32) {
        order = kind >= 32 ? kind - 31 : kind - 8;
        if (order > kMaxOrder) err |= kErrHost, order = 0;
        else if (order > n) err |= kErrFormat, order = 0;
        else {
            for (int i = 0; i < order; ++i) {                                        // order <= 12 < kChunk; folded like every tile entry
                const int32_t ws = r.take_signed(bps);
                tile0[i * 32] = (int32_t)(((uint32_t)ws << 1) ^ (uint32_t)(ws >> 31));
            }
            if (kind >= 32) {
                const int prec = (int)r.take(4) + 1;
                if (prec == 16) err |= kErrFormat;
                shift = r.take_signed(5);
                if (shift < 0) err |= kErrFormat, shift = 0;
                for (int j = 0; j < order; ++j) hand.c[j] = r.take_signed(prec);
            } else {
                hand.c[0] = order == 1 ? 1 : order == 2 ? 2 : order == 3 ? 3 : order == 4 ? 4 : 0;
                hand.c[1] = order == 2 ? -1 : order == 3 ? -3 : order == 4 ? -6 : 0;
                hand.c[2] = order == 3 ? 1 : order == 4 ? 4 : 0;
                hand.c[3] = order == 4 ? -1 : 0;
            }
            const int method = (int)r.take(2);
            if (method > 1) err |= kErrFormat;
            st.pbits = 4 + (method & 1);
            const int porder = (int)r.take(4);
            st.psize = n >> porder;
            if ((porder && (st.psize << porder) != n) || st.psize < order) err |= kErrFormat;
            st.part_end = order;             // the first partition's parameter is read when the loop reaches i == order
            st.psize = st.psize > 0 ? st.psize : n;
        }
    } else {
        err |= kErrFormat;
    }
    if (err) {                               // nothing more is read: the frame decodes to zeros and is reported
        st.fast_until = 0;
        st.rice = false;
        st.raw = 0;
        st.cval = 0;
        st.part_end = n;
        order = 0;
#pragma unroll
        for (int j = 0; j < kMaxOrder; ++j) hand.c[j] = 0;
    }
    st.order = order;
    st.fast_until = order;                   // the first residual goes the slow way and reads its partition's parameter
    st.err = err;
    hand.order = order;
    hand.shift = shift;
    hand.wasted = wasted;
    hand.err = err;
}

// One sample the slow way: partition parameters, escape / verbatim / constant values, long Rice symbols, samples outside
// the residual range (nothing to do).  Returns the residual.
__device__ __forceinline__ int32_t reader_slow_sample(ReaderState& st, int i) {
    Reader& r = st.r;
    if ((unsigned)(i - st.order) >= (unsigned)(st.n - st.order)) return 0;            // warm-up (already in the tile) or past the end
    while (i == st.part_end) {               // a partition may be empty (predictor order == partition size): then the next one starts here too
        const int k = (int)r.take(st.pbits);
        st.rice = true;
        st.k = k;
        if (k == (1 << st.pbits) - 1) {
            st.rice = false;
            st.raw = (int)r.take(5);
            st.cval = 0;
        }
        st.part_base += st.psize;
        st.part_end = st.part_base;
    }
    st.fast_until = st.rice ? (st.part_end < st.n ? st.part_end : st.n) : st.order;     // == order: no fast samples
    if (st.rice) {
        r.service();
        const int z = clz32(r.hi);
        const int n = z + 1 + st.k;
        const uint32_t v = n <= 32 ? r.rice_fast(st.k, z, n, r.word()) : r.rice(st.k);   // a partition's first symbol is as short as any
        return (int32_t)(v >> 1) ^ -(int32_t)(v & 1);
    }
    return st.raw ? r.take_signed(st.raw) : st.cval;
}

// Residuals i0 .. i0 + kChunk of this lane's frame into its tile column (stride 32 words).  The hot path (a Rice symbol of
// at most 32 bits inside a partition) is straight-line code behind ONE test; everything else goes through reader_slow_sample.
__device__ __forceinline__ void reader_chunk(ReaderState& st, int32_t* tile, int i0) {
    Reader& r = st.r;
    r.service();
    unsigned fast_span = (unsigned)(st.fast_until - st.order);       // refreshed whenever the slow path moved fast_until
#pragma unroll 1
    for (int i = i0; i < i0 + kChunk; ++i, tile += 32) {
        const uint32_t w = r.word();
        const int z = clz32(r.hi);
        const int n = z + 1 + st.k;
        // the tile carries FOLDED residuals (the Rice code's own unsigned form, 2r for r >= 0 and -2r - 1 for r < 0): unfolding
        // costs the predictor warp three instructions it has time for, and takes four off this warp, which is the critical one
        uint32_t v;
        if ((unsigned)(i - st.order) < fast_span && n <= 32) {
            v = r.rice_fast(st.k, z, n, w);
        } else {
            const int32_t res = reader_slow_sample(st, i);
            v = ((uint32_t)res << 1) ^ (uint32_t)(res >> 31);
            fast_span = (unsigned)(st.fast_until - st.order);
            if (i < st.order) continue;      // warm-up samples are in the tile already
        }
        *tile = (int32_t)v;
    }
}

struct PredictorState {
    int32_t c[kMaxOrder], h[kMaxOrder];
    int64_t ahead;
    int n, order, shift, wasted, lo, hi;
    int16_t* out;
};

__device__ __forceinline__ void predictor_init(PredictorState& p, const Hand& hand) {
#pragma unroll
    for (int j = 0; j < kMaxOrder; ++j) p.c[j] = hand.c[j], p.h[j] = 0;
    p.order = hand.order;
    p.shift = hand.shift;
    p.wasted = hand.wasted;
    p.ahead = 0;
    if (hand.err) p.lo = p.hi = 0;           // a rejected frame writes nothing
}

// kChunk is a multiple of kMaxOrder: inside a group of 12 samples the history slots rotate at compile time (sample u of
// the group overwrites slot 11 - u, tap t of the next sample reads slot (12 - u + t) % 12 ... ), so no register is moved.
__device__ __forceinline__ void predictor_chunk(PredictorState& p, const int32_t* tile, int i0) {
#pragma unroll 1
    for (int g = 0; g < kChunk; g += kMaxOrder) {
        if (i0 + g >= p.n) break;
#pragma unroll
        for (int u = 0; u < kMaxOrder; ++u) {
            // slot of the sample t steps back, before sample u of the group is stored: (kMaxOrder - u + t) % kMaxOrder
            const int i = i0 + g + u;
            const uint32_t fv = (uint32_t)tile[(g + u) * 32];
            const int32_t res = (int32_t)(fv >> 1) ^ -(int32_t)(fv & 1);
            const int32_t pred = (int32_t)(mad_wide(p.c[0], p.h[(kMaxOrder - u) % kMaxOrder], p.ahead) >> p.shift);
            const int32_t s = res + (i >= p.order ? pred : 0);
            if ((unsigned)(i - p.lo) < (unsigned)(p.hi - p.lo)) p.out[i] = (int16_t)(s << p.wasted);
            p.h[kMaxOrder - 1 - u] = s;      // overwrites the oldest sample; it is now "0 steps back" for sample u + 1
            int64_t a0 = 0, a1 = 0;
#pragma unroll
            for (int t = 1; t < kMaxOrder; t += 2) a0 = mad_wide(p.c[t], p.h[(kMaxOrder - (u + 1) + t) % kMaxOrder], a0);
#pragma unroll
            for (int t = 2; t < kMaxOrder; t += 2) a1 = mad_wide(p.c[t], p.h[(kMaxOrder - (u + 1) + t) % kMaxOrder], a1);
            p.ahead = a0 + a1;
        }
    }
}

__device__ __forceinline__ uint32_t crc16_bytes(const uint16_t* t16, const unsigned char* p, const unsigned char* e) {
    uint32_t crc = 0;
    while (p < e && (reinterpret_cast<uintptr_t>(p) & 3)) crc = ((crc << 8) ^ t16[((crc >> 8) ^ *p++) & 0xFF]) & 0xFFFF;
    for (; p + 4 <= e; p += 4) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
        crc = ((crc << 8) ^ t16[((crc >> 8) ^ w) & 0xFF]) & 0xFFFF;
        crc = ((crc << 8) ^ t16[((crc >> 8) ^ (w >> 8)) & 0xFF]) & 0xFFFF;
        crc = ((crc << 8) ^ t16[((crc >> 8) ^ (w >> 16)) & 0xFF]) & 0xFFFF;
        crc = ((crc << 8) ^ t16[((crc >> 8) ^ (w >> 24)) & 0xFF]) & 0xFFFF;
    }
    while (p < e) crc = ((crc << 8) ^ t16[((crc >> 8) ^ *p++) & 0xFF]) & 0xFFFF;
    return crc;
}

// Reader's epilogue: the frame must end on the next byte boundary + CRC-16 exactly where the host found the next header.
// Returns the frame length it decoded (header + subframe + CRC).
__device__ __forceinline__ int64_t reader_finish(ReaderState& st, const oe_flac_frame& fr) {
    if (st.r.over) st.err |= kErrOverrun;
    const int64_t total = fr.hdr_bytes + ((st.r.consumed() + 7) >> 3) + 2;
    if (!st.err && (fr.frame_bytes > 0 ? total != fr.frame_bytes : total > -fr.frame_bytes)) st.err |= kErrEnd;
    return total;
}

#ifdef __CUDACC__
__device__ __forceinline__ void pair_barrier() { asm volatile("bar.sync 1, 64;" ::: "memory"); }   // reader + predictor warps

__global__ void __launch_bounds__(96) oe_flac_decode_kernel(const unsigned char* __restrict__ comp, int64_t comp_limit,
                                                           const oe_flac_frame* __restrict__ frames, int64_t n_frames,
                                                           int16_t* __restrict__ pcm, int32_t* __restrict__ errors, int verify_crc) {
    __shared__ uint16_t t16[256];
    __shared__ __align__(16) uint32_t rings[32][kRingWords];
    __shared__ int32_t tile[2][kChunk][32];
    __shared__ Hand hands[32];
    __shared__ int32_t decoded_bytes[32];
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint16_t w = (uint16_t)(i << 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) w = (uint16_t)((w & 0x8000) ? (w << 1) ^ 0x8005 : w << 1);
        t16[i] = w;
    }
    const int64_t f = (int64_t)blockIdx.x * 32 + lane;
    const bool live = f < n_frames;
    oe_flac_frame fr;
    fr.block = 0;
    if (live) fr = frames[f];
    int nmax = live ? fr.block : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, d));
    const int rounds = (nmax + kChunk - 1) / kChunk;
    __syncthreads();
    if (role == 0) {                         // ---- reader ----
        ReaderState st;
        st.n = live ? fr.block : 0;
        st.order = 0;
        st.err = 0;
        st.part_end = st.n;
        st.rice = false, st.raw = 0, st.cval = 0, st.k = 0, st.pbits = 4, st.psize = 1, st.part_base = 0, st.fast_until = 0;
        if (live) {
            st.r.init(comp, fr.comp_off + fr.hdr_bytes, comp_limit, rings[lane]);
            reader_prologue(st, hands[lane], &tile[0][0][lane], fr.bps);
        } else {
            hands[lane].err = kErrFormat;
        }
        pair_barrier();
        for (int rd = 0; rd < rounds; ++rd) {
            if (live) reader_chunk(st, &tile[rd & 1][0][lane], rd * kChunk);
            pair_barrier();
        }
        int32_t total = 0;
        if (live) {
            total = (int32_t)reader_finish(st, fr);
            if (st.err) atomicOr(errors + fr.utt, st.err);
        }
        decoded_bytes[lane] = st.err ? 0 : total;
    } else if (role == 1) {                  // ---- predictor ----
        PredictorState p;
        p.n = live ? fr.block : 0;
        p.lo = live ? fr.skip : 0;
        p.hi = live ? fr.skip + fr.take : 0;
        p.out = live ? pcm + (fr.out_off - fr.skip) : pcm;
        pair_barrier();
        predictor_init(p, hands[lane]);
        for (int rd = 0; rd < rounds; ++rd) {
            pair_barrier();
            predictor_chunk(p, &tile[rd & 1][0][lane], rd * kChunk);
        }
    }
    // ---- CRC-16: warp 2 works while the other two decode; a last frame's length is only known afterwards, so it is
    // checked over "everything that is left" (no trailing bytes: the normal case) and redone by its reader lane otherwise ----
    uint32_t crc = 0;
    int32_t assumed = 0;
    if (role == 2 && live && verify_crc) {
        assumed = fr.frame_bytes > 0 ? fr.frame_bytes : -fr.frame_bytes;
        if (assumed >= 2) crc = crc16_bytes(t16, comp + fr.comp_off, comp + fr.comp_off + assumed - 2);
    }
    __syncthreads();
    if (role == 2 && live && verify_crc && decoded_bytes[lane] > 0) {
        const int32_t total = decoded_bytes[lane];
        const unsigned char* const b = comp + fr.comp_off;
        if (total != assumed) crc = crc16_bytes(t16, b, b + total - 2);
        if (crc != (uint32_t)(b[total - 2] << 8 | b[total - 1])) atomicOr(errors + fr.utt, kErrCrc);
    }
}
#else
// Host emulation of one block of ONE frame (oe_emul.cpp): the same reader / predictor / CRC code, run round by round.
inline void emulate_frame(const unsigned char* comp, int64_t comp_limit, const oe_flac_frame& fr, int16_t* pcm, int32_t* errors,
                          int verify_crc) {
    static uint16_t t16[256];
    for (int i = 0; i < 256; ++i) {
        uint16_t w = (uint16_t)(i << 8);
        for (int k = 0; k < 8; ++k) w = (uint16_t)((w & 0x8000) ? (w << 1) ^ 0x8005 : w << 1);
        t16[i] = w;
    }
    static uint32_t ring[kRingWords];
    static int32_t tile[2][kChunk][32];
    Hand hand;
    ReaderState st;
    st.n = fr.block;
    st.order = 0;
    st.err = 0;
    st.part_end = st.n;
    st.rice = false, st.raw = 0, st.cval = 0, st.k = 0, st.pbits = 4, st.psize = 1, st.part_base = 0, st.fast_until = 0;
    st.r.init(comp, fr.comp_off + fr.hdr_bytes, comp_limit, ring);
    reader_prologue(st, hand, &tile[0][0][0], fr.bps);
    PredictorState p;
    p.n = fr.block;
    p.lo = fr.skip;
    p.hi = fr.skip + fr.take;
    p.out = pcm + (fr.out_off - fr.skip);
    predictor_init(p, hand);
    const int rounds = (fr.block + kChunk - 1) / kChunk;
    // the reader runs one round ahead of the predictor, as on the device (double-buffered tile)
    for (int rd = 0; rd <= rounds; ++rd) {
        if (rd < rounds) reader_chunk(st, &tile[rd & 1][0][0], rd * kChunk);
        if (rd > 0) predictor_chunk(p, &tile[(rd - 1) & 1][0][0], (rd - 1) * kChunk);
    }
    const int64_t total = reader_finish(st, fr);
    if (st.err) errors[fr.utt] |= st.err;
    else if (verify_crc) {
        const unsigned char* const b = comp + fr.comp_off;
        if (crc16_bytes(t16, b, b + total - 2) != (uint32_t)(b[total - 2] << 8 | b[total - 1])) errors[fr.utt] |= kErrCrc;
    }
}
#endif

}  // namespace oe_flacgpu
