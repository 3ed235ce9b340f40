// oe_fbank2_kernel: second-generation ragged-batch Kaldi fbank for sm_100a, standard 80-bin mel matrix (torchaudio's
// values, baked) -- included by oe_frontend.cu; oe_fbank_kernel stays for other matrices (same structure with other
// weights: compile-time structure, parameter weights; anything else: table-driven).
//
// ncu on the first-generation kernel showed the shared-memory data pipe at 72 % of peak (162 wavefronts per
// frame) and 31 % of the issued instructions spent on integer address arithmetic; HBM sat at 3 %.  This
// kernel keeps the same arithmetic (kaldi.py:183-211, 616-633) and the same tile / descriptor interface
// but is laid out around the shared-memory pipe:
//   * no fp32 staging pass: stage A reads the raw int16 PCM straight from the cp.async buffer (LDS.32 =
//     two samples), converts, applies pre-emphasis (neighbour sample by one warp shuffle) and the frame
//     mean in registers;
//   * every shared-memory access is `per-lane base + compile-time immediate` (padded rows instead of XOR
//     swizzles), so the FFT stages carry no integer instructions;
//   * radix-4 16-point FFTs (oe_fft.h: fft16_r4), per-lane twiddles as scalar FMAs on the register halves;
//   * the conjugate-partner exchange moves HALF of Z: lane k1 keeps k2 = 0..7, publishes k2 = 8..15 and the
//     pair untangle (oe_fft.h: untangle_pair) yields bins k and 256 - k from one (P, Q) pair;
//   * every 16-thread group owns a private power-spectrum slice inside its own exchange area, so power is
//     stored straight from the untangle (no register staging, one CTA barrier fewer);
//   * no constant loads on the tile's critical path: mel weights and resampler taps are FFMA immediates (baked
//     torchaudio tables, oe_mel80.h / oe_rs_coefs.h), twiddles of the FFT stages are compile-time constants.
// Measured per frame (ncu, profiles/): 106 shared-memory wavefronts and 374 warp instructions (was 162 and 459);
// 98 us per 150 k frames (was 128 us).
//
// Work split: CTA = 256 threads = 16 groups of 16; 32-frame tile; warp w owns frames 4w .. 4w+3; group
// g = 0/1 of the warp packs frames (4w + g, 4w + g + 2) into the two halves of f32x2 registers (the +2
// pairing puts the two groups' LDS.32 on disjoint banks).
#pragma once

namespace oe {
namespace k2 {

constexpr int kXGroup = 4608;      // bytes of exchange area per group
constexpr int kXRow = 144;         // exchange 1: row k1 = 16 cells of 8 B (frame pair) + 16 B pad -> conflict-free LDS.128
constexpr int kXPlane = 2304;      // exchange 1: imaginary plane
constexpr int kPubRow = 72;        // published half rows: 8 cells + 1 pad cell (odd cell stride: conflict-free both ways)
constexpr int kPubPlane = 1152;
constexpr int kPwBase = 2304;      // power slice [257 bins][2 frames] behind the published rows
constexpr int kOutRow = 81;        // floats per output-tile row (80 + 1 pad)
constexpr int kF = 80;

// slice offset: the 32 (group, half) columns of a tile land on 32 distinct banks for the mel reads
__host__ __device__ constexpr int slice_off(int grp) { return 64 * (grp & 1) + 8 * (grp >> 1); }
static_assert(kPwBase + slice_off(15) + 257 * 8 <= kXGroup, "power slice must fit the group area");
static_assert(2 * 16 * kPubRow <= kPwBase, "published rows must stay below the power slice");

template <bool kF32, bool kRs>
struct Smem {
    static constexpr int X = 0;                                        // 16 group areas (rs: fp32 resampler input)
    static constexpr int Raw = 16 * kXGroup;                           // cp.async landing buffer
    static constexpr int RawBytes = kF32 ? (5376 + 8) * 4 : kRsPieces * 16;
    static constexpr int Tile = Raw + RawBytes;                        // output tile [32][81]; rs: resampled fp32 tile first
    static constexpr int TileBytes = kRs ? 5376 * 4 : 32 * kOutRow * 4;
    static constexpr int TwA = Tile + TileBytes;                       // float2 (cos, sin) W256^(tau k1): 16 rows x 144 B
    static constexpr int TwU = TwA + 16 * 144;                         // float2 (cos, sin)(2 pi (k1 + 16 k2) / 512), k2 < 8: 16 rows x 80 B
    static constexpr int Mask = TwU + 16 * 80;                         // uchar rowmask[32], colmask[96]
    static constexpr int Desc = Mask + 128;                            // TileDesc[2]
    static constexpr int Acc = Desc + 96;                              // long long acc[3][2][80]: fixed-point CMVN statistics
    static constexpr int Flag = Acc + 3 * 2 * kF * 8;                  // int[0]: "this is the last CTA", int[1]: claimed tile, +8: mbarrier
    static constexpr int End = Flag + 16;
    static_assert(Raw % 16 == 0 && Tile % 16 == 0 && TwA % 16 == 0 && TwU % 16 == 0 && Desc % 16 == 0 && Acc % 8 == 0, "align");
    static_assert(End + 1024 <= 116224, "two CTAs per SM");
    static_assert(!kRs || kRsMaxIn * 4 <= 16 * kXGroup, "fp32 resampler input must fit the exchange areas");
    static_assert(32 * kOutRow * 4 <= TileBytes, "out tile");
};

// V2 <-> its 64-bit register image (the host compilation pass only needs these to parse)
__device__ __forceinline__ V2 v2_from_bits(unsigned long long b) {
    V2 r;
#if defined(__CUDA_ARCH__)
    r.v = b;
#else
    r.lo = r.hi = (float)b;
#endif
    return r;
}
__device__ __forceinline__ unsigned long long v2_bits(V2 a) {
#if defined(__CUDA_ARCH__)
    return a.v;
#else
    return (unsigned long long)a.lo;
#endif
}
__device__ __forceinline__ V2 lds_v2(const unsigned char* p) { return v2_from_bits(*reinterpret_cast<const unsigned long long*>(p)); }
__device__ __forceinline__ void sts_v2(unsigned char* p, V2 v) { *reinterpret_cast<unsigned long long*>(p) = v2_bits(v); }

// Standard mel projection of one warp's bin group from the per-group power slices (pcol[2 k] = 4 |X[k]|^2 of this
// lane's frame).  Structure AND weights are compile-time (oe_mel80.h: torchaudio's matrix, x 1/4): every weight is a
// 32-bit immediate of its FFMA.  Kernel-parameter constants cost a uniform constant load for every third weight, each
// worth ~15 cycles of the tile's critical path: immediates made the kernel 4 % faster.  Handles whose matrix has the
// structure but other values run the first-generation kernel, which takes the weights as parameters.
template <int G>
__device__ __forceinline__ void mel_group2(const float* __restrict__ pcol, const FbankParams& P,
                                           float* __restrict__ orow, float log_floor) {
    constexpr int b0 = mel80::kGroup[G], b1 = mel80::kGroup[G + 1];
    float acc[b1 - b0];
    static_for<b0, b1>([&](auto bb) {
        constexpr int b = decltype(bb)::value;
        constexpr int k0 = mel80::kStart[b], off = mel80::kOff[b];
        float a = 0.f;
        static_for<0, mel80::kLen[b]>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            a = fmaf(mel80::weight(off + i), pcol[2 * (k0 + i)], a);
        });
        acc[b - b0] = a;
    });
    static_for<b0, b1>([&](auto bb) {
        constexpr int b = decltype(bb)::value;
        orow[b] = fast_ln(fmaxf(acc[b - b0], log_floor));
    });
}

// Fixed-point scales of the CMVN-statistics accumulators (compute_cmvn_stats): integer adds commute, so the call's sums
// are bitwise reproducible however the tiles are spread over CTAs, and the CTAs can add straight into one global
// accumulator (no per-CTA partials, no reduction kernel).  Range per call: |sum| < 2^63 / 2^28 = 3.4e10 (log-mel
// magnitudes <= 40: 8e8 frames), sum of squares < 2^63 / 2^24 = 5.5e11 (3e8 frames = 950 h of audio in one batch).
constexpr double kFxSum = 268435456.0;      // 2^28
constexpr double kFxSq = 16777216.0;        // 2^24
__device__ __forceinline__ long long fx(double v, double scale) { return __double2ll_rn(v * scale); }
// release / acquire fence at GPU scope
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// kDither: kaldi.fbank's dither (kaldi.py:179-181), one independent N(0,1) per frame element before DC removal; a
// separate instantiation because the Philox + Box-Muller code roughly triples the front.
// ---- bulk-copy staging of interior tiles: one thread hands the whole contiguous PCM window of the next tile to the
// copy unit (cp.async.bulk, completion counted in bytes on an mbarrier) instead of 256 threads issuing three 16-byte
// cp.async each.  Tiles that touch an utterance's first or last samples keep the cp.async path (its src-size form
// zero-fills what lies outside the utterance). ----
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar,
                                          unsigned long long policy) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // earlier generic-proxy accesses of the buffer are ordered first
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy) : "memory");
}
// Stages the next tile's samples; returns true when the bulk path was taken (every thread computes the same answer from
// the descriptor): the caller then waits on the mbarrier instead of its own cp.async group.
template <bool kF32>
__device__ __forceinline__ bool prefetch_tile2(unsigned char* raw, const void* wav, const TileDesc* d, unsigned long long* bar, int tid) {
    const int in_first = d->in_first, in_len = d->in_len;
    const int pieces = kF32 ? 673 * 2 : (d->rs ? kRsPieces : 673);
    const int per = kF32 ? 4 : 8;                                      // samples per 16-byte piece
    if (in_first >= 0 && in_first + per * pieces <= in_len) {
        if (tid == 0) {
            const unsigned char* const src = reinterpret_cast<const unsigned char*>(wav) + (d->wav_utt + in_first) * (kF32 ? 4 : 2);
            bulk_load(raw, src, 16u * (unsigned)pieces, bar, l2_policy_evict_first());
        }
        return true;
    }
    prefetch_tile<kF32>(raw, wav, d, tid);
    return false;
}

template <bool kF32, bool kRs, bool kDither = false>
__global__ void __launch_bounds__(kThreads, 2) oe_fbank2_kernel(const FbankParams P) {
    using S = Smem<kF32, kRs>;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char* const sRaw = smem + S::Raw;
    float* const sTile = reinterpret_cast<float*>(smem + S::Tile);
    unsigned char* const sRowMask = smem + S::Mask;
    unsigned char* const sColMask = sRowMask + 32;
    TileDesc* const sDesc = reinterpret_cast<TileDesc*>(smem + S::Desc);
    long long* const sAcc = reinterpret_cast<long long*>(smem + S::Acc);
    int* const sFlag = reinterpret_cast<int*>(smem + S::Flag);
    unsigned long long* const sBar = reinterpret_cast<unsigned long long*>(smem + S::Flag + 8);    // bulk-copy completion

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tau = tid & 15, grp = tid >> 4;
    const int fA = 4 * warp + (grp & 1);                 // packed frames: lo = fA, hi = fA + 2
    const DevTables* __restrict__ tab = P.tab;
    constexpr int F = kF, rowO = kOutRow;

    if (P.stat_acc != nullptr)
        for (int i = tid; i < 3 * 2 * F; i += kThreads) sAcc[i] = 0;

    int tile = blockIdx.x;
    int slot = 0;
    // ---- one-time table staging: twiddles as (cos, +sin) ----
    {
        float2* const twA = reinterpret_cast<float2*>(smem + S::TwA);
        float2* const twU = reinterpret_cast<float2*>(smem + S::TwU);
        {
            const int r = tid >> 4, c = tid & 15;
            const float2 t = tab->twA[r * kRowE + c];    // (cos, -sin)
            twA[r * 18 + c] = make_float2(t.x, -t.y);
        }
        if (tid < 128) {
            const int r = tid >> 3, c = tid & 7;
            twU[r * 10 + c] = tab->twU[r * kRowE + c];   // (cos, sin)
        }
    }
    float wv0[13], wv1[13];                    // window taps of this lane: w[32 n1 + 2 tau (+1)] (0 beyond 399)
#pragma unroll
    for (int n1 = 0; n1 < 13; ++n1) {
        wv0[n1] = tab->window[32 * n1 + 2 * tau];
        wv1[n1] = tab->window[32 * n1 + 2 * tau + 1];
    }
    // ---- everything above only read launch constants: with a programmatic dependent launch it overlaps the tail of
    // the descriptor kernel; the tile descriptors are touched from here on ----
    if (tid == 0) mbar_init(sBar, 1);
    unsigned bulk_parity = 0;                  // phase of the mbarrier the next bulk copy completes
    bool bulk_pending = false;
    grid_dep_wait();
    if (tile < P.total_tiles) {                // first tile: descriptor, then its waveform
        prefetch_desc(sDesc, P.tiles + tile, tid);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        if (sDesc[0].nvalid > 0) bulk_pending = prefetch_tile2<kF32>(sRaw, P.wav, sDesc, sBar, tid);
        cp_async_commit();
    }
    const float preemph = tab->preemph;
    const float dc_scale = (1.0f - preemph) * (1.0f / (float)kWin);
    const float log_floor = tab->log_floor;
    const bool fused = (P.n_tmask | P.n_fmask) != 0;
    // two-phase pipeline: the raw rows are re-read by the finalize kernel two launches later -> keep them in L2
    const unsigned long long out_policy = P.keep_out_in_l2 ? l2_policy_evict_last() : l2_policy_evict_normal();
    const bool is15 = tau == 15;
    const int src_lane = (lane & 16) | ((tau + 15) & 15);
    const float m12 = tau < 8 ? 1.f : 0.f;     // chunk n1 = 12 holds samples 384 + 2 tau (+1): only tau < 8 are inside the frame

    unsigned char* const gx = smem + S::X + grp * kXGroup;
    unsigned char* const xw = gx + tau * 8;                               // stage A stores: + k1 * kXRow (+ kXPlane)
    const unsigned char* const xr = gx + tau * kXRow;                     // stage B row loads: + 16 c (+ kXPlane)
    unsigned char* const pubw = gx + tau * kPubRow;                       // publish: + (k2 - 8) * 8 (+ kPubPlane)
    const unsigned char* const pubr = gx + ((16 - tau) & 15) * kPubRow;                     // partner: + (7 - k2) * 8
    unsigned char* const pwA = gx + kPwBase + slice_off(grp) + tau * 8;           // bin k1 + 16 k2: + 128 k2
    unsigned char* const pwB = gx + kPwBase + slice_off(grp) + (256 - tau) * 8;   // bin 256 - k:  - 128 k2

    // Dynamic tile order: a CTA's first two tiles are blockIdx.x and blockIdx.x + gridDim.x, every further one is claimed
    // from a counter (P.sched[1], zeroed by the descriptor kernel) two iterations ahead of its use -- one iteration for
    // the claim to come back, one for the prefetch of the tile.  Tiles with a fused speed perturb cost ~1.4x a plain
    // tile: a fixed stride left the slowest CTA 12 % behind the mean, claimed tiles 3 % (host simulation of the benchmark
    // batch).  Nothing depends on which CTA runs a tile: tile statistics are per tile, the CMVN sums are integer adds.
    // Measured (same box, A/B): plain instantiation 99.0 -> 95.9 us, LibriSpeech-shape pass -2 %, streaming windows -4 %;
    // the instantiation with the fused speed perturb got SLOWER (123.9 -> 128.0 us in the benchmark step -- with the claim
    // compiled in but switched off as well: the loop restructuring alone moves ptxas' register allocation of that
    // 128-register body) and keeps the fixed stride.
    // (kDyn is a compile-time switch so that the fused-speed-perturb instantiation compiles to exactly the fixed-stride loop)
    constexpr bool kDyn = !kRs;
    int next_dyn = tile + gridDim.x, claim = P.total_tiles;
    const bool dyn = kDyn && P.sched != nullptr;
    for (; tile < P.total_tiles; slot ^= 1) {
        const int next = kDyn ? next_dyn : tile + (int)gridDim.x;
        cp_async_wait_all();
        if (bulk_pending) {
            mbar_wait(sBar, bulk_parity);
            bulk_parity ^= 1;
            bulk_pending = false;
        }
        __syncthreads();               // (1) raw samples + descriptor visible; previous tile's rows are out of smem
        const TileDesc* const dp = sDesc + slot;
        const int b = dp->b, t0 = dp->t0;
        const int nvalid = dp->nvalid;                             // <= 0: padding-only tile
        const int rows_here = dp->rows_here;
        const long long out_start = dp->out_start;
        if (next < P.total_tiles) prefetch_desc(sDesc + (slot ^ 1), P.tiles + next, tid);
        cp_async_commit();
        // inline PTX: atomicAdd() is turned into a warp-aggregated atomic whose result is shuffled out at once -- warp 0
        // would sit out the L2 round trip at the top of every tile; this way the wait comes where `claim` is stored
        if (kDyn && tid == 0) {
            claim = dyn ? 0 : next + (int)gridDim.x - 2 * (int)gridDim.x;
            if (dyn && next < P.total_tiles) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(claim) : "l"(P.sched + 1) : "memory");
            else if (dyn) claim = P.total_tiles;
        }
        if (fused) {
            if (tid < 32) {
                const int t = t0 + tid;
                bool m = false;
                for (int j = 0; j < P.n_tmask; ++j) {
                    const int32_t* r = P.tmask + ((int64_t)b * P.n_tmask + j) * 2;
                    m |= (t >= r[0]) & (t < r[1]);
                }
                sRowMask[tid] = m;
            } else if (tid < 32 + F) {
                const int f = tid - 32;
                bool m = false;
                for (int j = 0; j < P.n_fmask; ++j) {
                    const int32_t* r = P.fmask + ((int64_t)b * P.n_fmask + j) * 2;
                    m |= (f >= r[0]) & (f < r[1]);
                }
                sColMask[f] = m;
            }
        }
        if (nvalid > 0) {
            const int rs = kRs ? dp->rs : 0;
            // ---- [fused speed perturb: raw int16 -> fp32 -> polyphase sinc -> resampled fp32 tile] ----
            if (kRs && rs != 0) {
                float* const xin = reinterpret_cast<float*>(smem + S::X);
                const int16_t* const r16 = reinterpret_cast<const int16_t*>(sRaw);
                for (int i = tid; i < kRsPieces * 2; i += kThreads) {       // 4 samples per thread-iteration
                    const int2 v = *reinterpret_cast<const int2*>(r16 + 4 * i);
                    float4 f;
                    f.x = (float)(int16_t)(v.x & 0xffff); f.y = (float)(v.x >> 16);
                    f.z = (float)(int16_t)(v.y & 0xffff); f.w = (float)(v.y >> 16);
                    *reinterpret_cast<float4*>(xin + 4 * i) = f;
                }
                __syncthreads();
                const int shift = rs_first_input(rs, t0) - dp->in_first;   // 0..7
                for (int mi = tid + 1; mi < kRsBlocks; mi += kThreads) {   // block mi -> tile samples 10 mi - 10 .. 10 mi - 1
                    float y[10];
                    if (rs == 1) rs_block_baked<9>(xin + shift + mi * 9, y);
                    else rs_block_baked<11>(xin + shift + mi * 11, y);
                    float2* const dst = reinterpret_cast<float2*>(sTile + 10 * mi - 10);
#pragma unroll
                    for (int j = 0; j < 5; ++j)
                        if (10 * mi - 10 + 2 * j < 5376) dst[j] = make_float2(y[2 * j], y[2 * j + 1]);
                }
                __syncthreads();
            }
            // ---- front: samples -> pre-emphasis (kaldi.py:193-198) -> DC removal (183-186) -> window (201-204) ----
            V2 zr[16], zi[16];
            {
                V2 acc;
                if (kF32 || (kRs && rs != 0)) {
                    const float* const tb = (kF32 ? reinterpret_cast<const float*>(sRaw) + 8 : sTile) + kShift * fA + 2 * tau;
                    float sa = 0.f, sb = 0.f, pa = 0.f, pb = 0.f;
#pragma unroll
                    for (int n1 = 0; n1 < 13; ++n1) {
                        float2 a = *reinterpret_cast<const float2*>(tb + 32 * n1);
                        float2 c = *reinterpret_cast<const float2*>(tb + 2 * kShift + 32 * n1);
                        if (kDither) {
                            const float2 na = dither_normals(b, t0 + fA, 16 * n1 + tau, P.dither_seed);
                            const float2 nc = dither_normals(b, t0 + fA + 2, 16 * n1 + tau, P.dither_seed);
                            a.x = fmaf(P.wav_dither, na.x, a.x); a.y = fmaf(P.wav_dither, na.y, a.y);
                            c.x = fmaf(P.wav_dither, nc.x, c.x); c.y = fmaf(P.wav_dither, nc.y, c.y);
                        }
                        if (n1 < 12) {
                            sa += a.x + a.y;
                            sb += c.x + c.y;
                        } else {
                            sa = fmaf(a.x + a.y, m12, sa);
                            sb = fmaf(c.x + c.y, m12, sb);
                        }
                        const float qa = __shfl_sync(0xffffffffu, is15 ? pa : a.y, src_lane);
                        const float qb = __shfl_sync(0xffffffffu, is15 ? pb : c.y, src_lane);
                        zr[n1] = v2_make(fmaf(-preemph, qa, a.x), fmaf(-preemph, qb, c.x));
                        zi[n1] = v2_make(fmaf(-preemph, a.x, a.y), fmaf(-preemph, c.x, c.y));
                        pa = a.y;
                        pb = c.y;
                    }
                    acc = v2_make(sa, sb);
                } else {
                    // word (s + 8) / 2 of the raw buffer holds samples s, s + 1 (s = tile-relative, even)
                    const uint32_t* const rw = reinterpret_cast<const uint32_t*>(sRaw) + 4 + (kShift / 2) * fA + tau;
                    const V2 npre = vbcast(-preemph);
                    acc = vbcast(0.f);
                    float pa = 0.f, pb = 0.f;
#pragma unroll
                    for (int n1 = 0; n1 < 13; ++n1) {
                        const uint32_t wa = rw[16 * n1], wb = rw[kShift + 16 * n1];
                        float ea = (float)(int16_t)(wa & 0xffffu), oa = (float)((int32_t)wa >> 16);
                        float eb = (float)(int16_t)(wb & 0xffffu), ob = (float)((int32_t)wb >> 16);
                        if (kDither) {
                            const float2 na = dither_normals(b, t0 + fA, 16 * n1 + tau, P.dither_seed);
                            const float2 nb = dither_normals(b, t0 + fA + 2, 16 * n1 + tau, P.dither_seed);
                            ea = fmaf(P.wav_dither, na.x, ea); oa = fmaf(P.wav_dither, na.y, oa);
                            eb = fmaf(P.wav_dither, nb.x, eb); ob = fmaf(P.wav_dither, nb.y, ob);
                        }
                        const V2 xe = v2_make(ea, eb), xo = v2_make(oa, ob);
                        if (n1 < 12) acc = vadd(acc, vadd(xe, xo));
                        else acc = vfma(vadd(xe, xo), vbcast(m12), acc);
                        const float qa = __shfl_sync(0xffffffffu, is15 ? pa : oa, src_lane);
                        const float qb = __shfl_sync(0xffffffffu, is15 ? pb : ob, src_lane);
                        zr[n1] = vfma(v2_make(qa, qb), npre, xe);
                        zi[n1] = vfma(xe, npre, xo);
                        pa = oa;
                        pb = ob;
                    }
                }
                float m0 = v2_lo(acc), m1 = v2_hi(acc);
#pragma unroll
                for (int o = 8; o >= 1; o >>= 1) {
                    m0 += __shfl_xor_sync(0xffffffffu, m0, o);
                    m1 += __shfl_xor_sync(0xffffffffu, m1, o);
                }
                const V2 dc = v2_make(m0 * dc_scale, m1 * dc_scale);      // (1 - preemph) * mean
#pragma unroll
                for (int n1 = 0; n1 < 13; ++n1) {
                    const V2 dr = vsub(zr[n1], dc), di = vsub(zi[n1], dc);
                    zr[n1] = v2_make(v2_lo(dr) * wv0[n1], v2_hi(dr) * wv0[n1]);
                    zi[n1] = v2_make(v2_lo(di) * wv1[n1], v2_hi(di) * wv1[n1]);
                }
            }
            // ---- stage A: 16-point FFT over n1 (z[16 n1 + tau]), twiddle W256^(tau k1), row k1 cell tau ----
            fft16_r4<true, V2>(zr, zi);
            {
                const float4* const tw4 = reinterpret_cast<const float4*>(smem + S::TwA + tau * 144);
                static_for<0, 8>([&](auto ii) {
                    constexpr int i = decltype(ii)::value;
                    constexpr int p0 = r4pos(2 * i), p1 = r4pos(2 * i + 1);
                    const float4 t = tw4[i];                       // k1 = 2i: (t.x, t.y), 2i + 1: (t.z, t.w)
                    V2 r0 = zr[p0], i0 = zi[p0], r1 = zr[p1], i1 = zi[p1];
                    if constexpr (i != 0) cmul_lane(r0, i0, t.x, t.y);
                    cmul_lane(r1, i1, t.z, t.w);
                    sts_v2(xw + (2 * i) * kXRow, r0);
                    sts_v2(xw + (2 * i) * kXRow + kXPlane, i0);
                    sts_v2(xw + (2 * i + 1) * kXRow, r1);
                    sts_v2(xw + (2 * i + 1) * kXRow + kXPlane, i1);
                });
            }
            __syncwarp();
            // ---- stage B: lane k1 = tau transforms row k1 of both frames ----
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                const int c = cc < 4 ? 2 * cc : 2 * cc - 7;        // chunks 0 2 4 6 (butterflies j = 0, 1 of the first pass) first
                const ulonglong2 vr = *reinterpret_cast<const ulonglong2*>(xr + 16 * c);
                const ulonglong2 vi = *reinterpret_cast<const ulonglong2*>(xr + kXPlane + 16 * c);
                zr[2 * c] = v2_from_bits(vr.x);
                zr[2 * c + 1] = v2_from_bits(vr.y);
                zi[2 * c] = v2_from_bits(vi.x);
                zi[2 * c + 1] = v2_from_bits(vi.y);
            }
            __syncwarp();                                          // every lane has its row: the area is reused
            fft16_r4<false, V2>(zr, zi);                           // position p holds Z[tau + 16 r4pos(p)]
            static_for<8, 16>([&](auto kk) {                       // publish k2 = 8..15 for the partner lane
                constexpr int k2 = decltype(kk)::value, pos = r4pos(k2);
                sts_v2(pubw + (k2 - 8) * 8, zr[pos]);
                sts_v2(pubw + (k2 - 8) * 8 + kPubPlane, zi[pos]);
            });
            __syncwarp();
            // ---- pair untangle: P = Z[k1 + 16 k2], Q = Z[256 - k] = row (16 - k1) mod 16, index 15 - k2.  Row 0 is its own
            // partner with index 16 - k2: lane 0 takes Q from its own registers (any shared-memory placement of its row
            // collides with exactly one other lane of the half-warp; measured: +9 wavefronts per frame) and its k2 = 0
            // pair only feeds the weightless bins 0 / 256 ----
            {
                const float4* const tu4 = reinterpret_cast<const float4*>(smem + S::TwU + tau * 80);
                const bool lane0 = tau == 0;
                static_for<0, 4>([&](auto ii) {
                    constexpr int i = decltype(ii)::value;
                    const float4 t = tu4[i];                       // k2 = 2i: (t.x, t.y), 2i + 1: (t.z, t.w)
                    static_for<0, 2>([&](auto jj) {
                        constexpr int k2 = 2 * i + decltype(jj)::value, pos = r4pos(k2), own = r4pos((16 - k2) & 15);
                        V2 qr = zr[own], qi = zi[own];
                        if (!lane0) {
                            qr = lds_v2(pubr + (7 - k2) * 8);
                            qi = lds_v2(pubr + (7 - k2) * 8 + kPubPlane);
                        }
                        V2 pk, pq;
                        untangle_pair<V2>(zr[pos], zi[pos], qr, qi, (k2 & 1) ? t.z : t.x, (k2 & 1) ? t.w : t.y, pk, pq);
                        sts_v2(pwA + 128 * k2, pk);
                        sts_v2(pwB - 128 * k2, pq);
                    });
                });
                constexpr int p8 = r4pos(8);                       // bin 128: X[128] = conj(Z[128]) (row 0, k2 = 8)
                const V2 z2 = vfma(zr[p8], zr[p8], vmul(zi[p8], zi[p8]));
                if (lane0) sts_v2(pwA + 128 * 8, vmul(z2, vbcast(4.f)));
            }
            if (kDyn && tid == 0) sFlag[1] = claim + 2 * (int)gridDim.x;
            cp_async_wait_all();       // next tile's descriptor has landed (this thread's pieces; (4) publishes them)
            __syncthreads();           // (4) power slices complete; the raw buffer has been consumed
            if (next < P.total_tiles) {
                const TileDesc* const dn = sDesc + (slot ^ 1);
                if (dn->nvalid > 0) bulk_pending = prefetch_tile2<kF32>(sRaw, P.wav, dn, sBar, tid);
            }
            cp_async_commit();

            // ---- sparse mel + log: warp = mel-bin group, lane = frame column (group lane >> 1, half lane & 1) ----
            {
                const int gl = lane >> 1, h = lane & 1;
                const int fr = 4 * (gl >> 1) + (gl & 1) + 2 * h;
                const float* const pcol = reinterpret_cast<const float*>(smem + S::X + gl * kXGroup + kPwBase + slice_off(gl) + 4 * h);
                float* const orow = sTile + fr * rowO;
                switch (warp) {
                    case 0: mel_group2<0>(pcol, P, orow, log_floor); break;
                    case 1: mel_group2<1>(pcol, P, orow, log_floor); break;
                    case 2: mel_group2<2>(pcol, P, orow, log_floor); break;
                    case 3: mel_group2<3>(pcol, P, orow, log_floor); break;
                    case 4: mel_group2<4>(pcol, P, orow, log_floor); break;
                    case 5: mel_group2<5>(pcol, P, orow, log_floor); break;
                    case 6: mel_group2<6>(pcol, P, orow, log_floor); break;
                    default: mel_group2<7>(pcol, P, orow, log_floor); break;
                }
            }
            __syncthreads();                                       // (5) output tile complete
            if (P.tile_stats != nullptr || P.stat_acc != nullptr) {
                for (int idx = tid; idx < 3 * F; idx += kThreads) {
                    const int rg = idx / F, f = idx - rg * F;
                    const int n = stats_rows(nvalid, rg);
                    float s = 0.f, m2 = 0.f;
                    if (n > 0) {
                        const float* col = sTile + (11 * rg) * rowO + f;
                        float x[11];
#pragma unroll
                        for (int r = 0; r < 11; ++r) x[r] = r < n ? col[r * rowO] : 0.f;
#pragma unroll
                        for (int r = 0; r < 11; ++r) s += x[r];            // + 0.f beyond n: exact
                        const float mean = s / (float)n;
#pragma unroll
                        for (int r = 0; r < 11; ++r) {
                            const float d = r < n ? x[r] - mean : 0.f;
                            m2 = fmaf(d, d, m2);
                        }
                        if (P.stat_acc != nullptr) {           // sum x^2 = M2 + n mean^2, fixed point: order-independent
                            sAcc[(rg * 2 + 0) * F + f] += fx((double)s, kFxSum);
                            sAcc[(rg * 2 + 1) * F + f] += fx((double)m2 + (double)s * (double)mean, kFxSq);
                        }
                    }
                    if (P.tile_stats != nullptr) {
                        float* const st = P.tile_stats + ((int64_t)tile * 3 + rg) * 2 * F;
                        st[f] = s;
                        st[F + f] = m2;
                    }
                }
            }
        } else {
            if (kDyn && tid == 0) sFlag[1] = claim + 2 * (int)gridDim.x;
            cp_async_wait_all();
            __syncthreads();
            if (next < P.total_tiles) {
                const TileDesc* const dn = sDesc + (slot ^ 1);
                if (dn->nvalid > 0) bulk_pending = prefetch_tile2<kF32>(sRaw, P.wav, dn, sBar, tid);
            }
            cp_async_commit();
        }

        // ---- rows out: [mask] -> [CMVN] -> coalesced stores; padding rows are 0 / (0-mean)*istd ----
        if (P.out != nullptr) {
            const bool has_cmvn = P.cmvn_mean != nullptr;
            const int pitch = (int)P.pitch;
            float* const dst0 = P.out + out_start * P.pitch;
            if (P.out_vec && !fused && !has_cmvn) {
                // float4 stores (pitch and base 16-byte aligned).  A warp step covers 4 rows x 8 float4: lane = (row & 3,
                // float4 & 7), so the four scalar reads per lane hit banks 4 (q + rg) + {0, 17, 2, 19}[row & 3] + i: all 32
                // distinct (tile rows are 81 floats apart), and every row gets one 128-byte store segment.
                constexpr int kQ = F / 4;                             // 20 float4 per row: q blocks of 8, 8, 4
                const int rl = lane & 3, ql = lane >> 2;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int it = warp + 8 * i;                      // 24 steps = 8 row groups x 3 q blocks
                    const int rg = it / 3, qb = it - 3 * rg;
                    const int r = 4 * rg + rl, q = 8 * qb + ql;
                    if (q < kQ && r < rows_here) {
                        const float* const src = sTile + r * rowO + 4 * q;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (r < nvalid) v = make_float4(src[0], src[1], src[2], src[3]);
                        float4* const d4 = reinterpret_cast<float4*>(dst0 + (int64_t)r * pitch + 4 * q);
                        st_f4_hint(d4, v, out_policy);
                    }
                }
            } else if (P.out_vec) {
                // the same lane mapping with [mask] -> [CMVN] applied (kept apart from the plain path above: merging them
                // cost 4 % on the plain path through register allocation of the whole kernel)
                constexpr int kQ = F / 4;                             // 20 float4 per row: q blocks of 8, 8, 4
                const int rl = lane & 3, ql = lane >> 2;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int it = warp + 8 * i;                      // 24 steps = 8 row groups x 3 q blocks
                    const int rg = it / 3, qb = it - 3 * rg;
                    const int r = 4 * rg + rl, q = 8 * qb + ql;
                    if (q < kQ && r < rows_here) {
                        const bool real = r < nvalid;
                        const float* const src = sTile + r * rowO + 4 * q;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (real) v = make_float4(src[0], src[1], src[2], src[3]);
                        if (fused && real) {
                            const uchar4 cm = *reinterpret_cast<const uchar4*>(sColMask + 4 * q);
                            const bool rm = sRowMask[r] != 0;
                            if (rm || cm.x) v.x = 0.f;
                            if (rm || cm.y) v.y = 0.f;
                            if (rm || cm.z) v.z = 0.f;
                            if (rm || cm.w) v.w = 0.f;
                        }
                        if (has_cmvn && (real || P.cmvn_on_pad)) {
                            const float* const m = P.cmvn_mean + 4 * q;
                            v = make_float4(v.x - __ldg(m), v.y - __ldg(m + 1), v.z - __ldg(m + 2), v.w - __ldg(m + 3));
                            if (P.cmvn_istd != nullptr) {
                                const float* const sd = P.cmvn_istd + 4 * q;
                                v = make_float4(v.x * __ldg(sd), v.y * __ldg(sd + 1), v.z * __ldg(sd + 2), v.w * __ldg(sd + 3));
                            }
                        }
                        *reinterpret_cast<float4*>(dst0 + (int64_t)r * pitch + 4 * q) = v;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = warp + 8 * i;
                    if (r < rows_here) {
                        const bool real = r < nvalid;
                        const bool rmask = fused && real && sRowMask[r];
                        const bool do_cmvn = has_cmvn && (real || P.cmvn_on_pad);
                        const float* const srow = sTile + r * rowO;
                        float* const dst = dst0 + r * pitch;
                        for (int f = lane; f < F; f += 32) {
                            float v = real ? srow[f] : 0.f;
                            if (rmask || (fused && real && sColMask[f])) v = 0.f;
                            if (do_cmvn) {
                                v = v - __ldg(P.cmvn_mean + f);
                                if (P.cmvn_istd != nullptr) v = v * __ldg(P.cmvn_istd + f);
                            }
                            dst[f] = v;
                        }
                    }
                }
            }
        }
        if (kDyn) {
            tile = next;
            next_dyn = sFlag[1];       // thread 0's claim, published before barrier (4) of this iteration
        } else {
            tile += gridDim.x;
        }
    }
    cp_async_wait_all();
    if (P.stat_acc != nullptr) {
        for (int idx = tid; idx < 3 * F; idx += kThreads) {      // each accumulator is owned by one thread: no barrier needed
            const int rg = idx / F, f = idx - rg * F;
            atomicAdd(P.stat_acc + f, (unsigned long long)sAcc[(rg * 2 + 0) * F + f]);
            atomicAdd(P.stat_acc + F + f, (unsigned long long)sAcc[(rg * 2 + 1) * F + f]);
        }
        // the last CTA to get here converts the call's integer sums and adds them to the caller's accumulator -- unless a
        // completion kernel follows in the stream (d_stats == null here): its first block does the conversion, and this
        // grid's tail is spared the fence / counter / read-back round trips (~3 us at the end of every launch)
        if (P.d_stats != nullptr) {
            __syncthreads();
            if (tid == 0) {
                fence_gpu();
                sFlag[0] = atomicAdd(P.sched, 1) == (int)gridDim.x - 1;
            }
            __syncthreads();
            if (sFlag[0]) {
                fence_gpu();
                for (int i = tid; i < 2 * F; i += kThreads) {
                    const long long a = (long long)__ldcg(P.stat_acc + i);
                    P.d_stats[i] += (double)a * (i < F ? 1.0 / kFxSum : 1.0 / kFxSq);
                }
                if (tid == 0) P.d_stats[2 * F] += P.stat_count;
            }
        }
    }
}

}  // namespace k2
}  // namespace oe
