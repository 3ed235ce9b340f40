// oe_fbank_kernel: ragged-batch Kaldi fbank for sm_100a (included by oe_frontend.cu).
//
// One CTA (256 threads) per 32-frame tile, persistent grid, 2 CTAs / SM.  Per tile:
//   (1) wait for the tile's raw samples (cp.async prefetch issued during the previous tile)
//   (2) convert int16/fp32 -> fp32, fold pre-emphasis (kaldi.py:193-198) into the staged signal,
//       8-sample block sums for the per-frame DC removal (kaldi.py:183-186)
//   (3) stage A  16 threads x 2 frames: mean (half-warp shuffles), window, 16-point DIF FFT in registers,
//       twiddle, half-warp exchange through shared memory
//   (4) stage B  two 16-point FFTs per lane, real-FFT untangle, |X|^2 -> power tile [256 bins][32 frames]
//   (5) sparse mel + log (kaldi.py:621-633): warp = bin group, lane = frame -> output tile
//   (6) [per-tile column statistics]  [mask -> CMVN]  coalesced row stores
// See DESIGN.md for the data layout and the per-phase instruction budget.
#pragma once

namespace oe {

// Philox-4x32-10 (Salmon et al., SC'11): counter-based generator, four 32-bit words per (counter, key).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// Two independent N(0, 1) draws for frame element pair `idx` of frame `frame` of utterance `b` (wav dither):
// Box-Muller on the first two Philox words.
__device__ __forceinline__ float2 dither_normals(unsigned b, unsigned frame, unsigned idx, unsigned long long seed) {
    const uint4 r = philox4x32_10(make_uint4(idx, frame, b, 0x57415644u), make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const float u1 = (float)((r.x >> 8) + 1u) * (1.0f / 16777216.0f);      // (0, 1]
    const float u2 = (float)(r.y >> 8) * (1.0f / 16777216.0f);             // [0, 1)
    const float rad = sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    return make_float2(rad * cs, rad * sn);
}

// Everything one tile needs, self-contained (no dependent loads): built on the device by
// oe_tile_desc_kernel and pulled into shared memory one tile ahead with cp.async.
struct __align__(16) TileDesc {
    int b;                  // utterance
    int t0;                 // first frame of the tile
    int nvalid;             // real frames in the tile (<= 0: padding-only tile)
    int rows_here;          // rows to write (real + padding)
    long long wav_utt;      // element index of the utterance's sample 0 in the waveform buffer
    int in_first;           // utterance-relative index of the first staged input sample (multiple of 8, may be < 0)
    int in_len;             // utterance length in input samples
    long long out_start;    // first output row of the tile
    int rs;                 // fused speed perturb: 0 none, 1 = 9:10 (speed 0.9), 2 = 11:10 (speed 1.1)
    int pad;
};
static_assert(sizeof(TileDesc) == 48, "TileDesc is three 16-byte cp.async pieces");

// Fused speed perturb (polyphase, NEW = 10 outputs per ORIG inputs, width 7).  A tile needs resampled samples
// [160 t0 - 10, 160 t0 + 5380) = 539 polyphase blocks starting at block 16 t0 - 1.
constexpr int kRsBlocks = 539;
__host__ __device__ __forceinline__ int rs_orig(int rs) { return rs == 1 ? 9 : 11; }
__host__ __device__ __forceinline__ int rs_first_input(int rs, int t0) { return (16 * t0 - 1) * rs_orig(rs) - 7; }
constexpr int kRsMaxIn = kRsBlocks * 11 + 14 + 8;          // staged input samples incl. alignment slack (5951)
constexpr int kRsPieces = (kRsMaxIn + 7) / 8;              // 16-byte int16 pieces (744)

struct FbankParams {
    const void* wav;
    const TileDesc* tiles;          // [total_tiles]
    float* out;
    int64_t pitch;
    int total_tiles;
    const int32_t* tmask;
    const int32_t* fmask;
    int n_tmask;
    int n_fmask;
    const float* cmvn_mean;
    const float* cmvn_istd;
    int cmvn_on_pad;
    int out_vec;                    // rows are dense (pitch == F) and `out` is 16-byte aligned: float4 stores allowed
    int keep_out_in_l2;             // two-phase: the raw rows are re-read by oe_finalize_kernel -> L2 evict_last stores
    float wav_dither;               // kaldi.fbank dither (gen-2 kernel, kDither instantiations), 0 = off
    unsigned long long dither_seed;
    float* tile_stats;              // [total_tiles][3][2][F]: per row-group column sum and sum of squared deviations
    double* cta_stats;              // [gridDim.x][3][2][F] or null: this CTA's running sum / sum of squares (fp64)
    const DevTables* tab;
    // ---- gen-2 kernel: CMVN statistics (compute_cmvn_stats) accumulated inside the kernel ----
    unsigned long long* stat_acc;   // [2F] fixed-point sum / sum of squares of this call (zeroed by the descriptor kernel); null = none
    int32_t* sched;                 // [0]: CTAs that have finished (zeroed likewise): the last one converts stat_acc
    double* d_stats;                // caller's accumulator [2F+1]: += the call's sums and frame count
    double stat_count;
    float mel_w[512];               // standard-structure fast path: weights (x 1/4) in mel80::kOff order;
                                    // lives in the kernel-parameter constant bank -> FFMA constant operands
    float rs_coef[2][256];          // fused speed perturb: sinc taps [new = 10][taps] for 9:10 and 11:10
};

// rows of a 32-frame tile covered by statistics row-group rg (0..2): [11 rg, min(nvalid, 11 rg + 11))
__host__ __device__ __forceinline__ int stats_rows(int nvalid, int rg) {
    const int lo = 11 * rg, hi = nvalid < lo + 11 ? nvalid : lo + 11;
    return hi > lo ? hi - lo : 0;
}

__device__ __forceinline__ int find_utt(const int32_t* __restrict__ prefix, int B, int tile) {
    int lo = 0, hi = B;                       // prefix[lo] <= tile < prefix[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (prefix[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------
// shared memory map (bytes)
constexpr int kGroupX = 4096;                              // exchange bytes per 16-thread group: 2 planes x 16 rows x 128 B
constexpr int kRowP = 34;                                  // floats per power-spectrum row: 32 frames + 2 pad (conflict-free STS.64)
constexpr int kSmP = 0;                                    // float p[5376]      | out tile (32 x (F+1))
constexpr int kSmS8 = kSmP + 5376 * 4;                     // float s8[672]
constexpr int kSmE = kSmS8 + kChunks * 4;                  // exchange buffers [16 groups][4096 B] | power tile [256][34]
constexpr int kSmRawF32 = kSmE + kBins * kRowP * 4;        // fp32 input: next tile's raw samples share the exchange area (behind the power tile)
constexpr int kSmRaw = kSmE + 16 * kGroupX;                // int16 input: dedicated prefetch buffer (745 pieces of 16 B)
constexpr int kSmTwA = kSmRaw + 11936;                     // float2 twA[16][18]
constexpr int kSmTwU = kSmTwA + 16 * kRowE * 8;            // float2 twU[16][18]
constexpr int kSmMask = kSmTwU + 16 * kRowE * 8;           // uchar rowmask[32], colmask[128]
constexpr int kSmDesc = kSmMask + 32 + kMaxMel;            // TileDesc[2]
constexpr int kSmAcc = kSmDesc + 2 * 48;                   // double acc[3][2][128]: CMVN-statistics accumulators
constexpr int kSmEdge = kSmAcc + 3 * 2 * kMaxMel * 8;      // float edge[32]: fused-resampler block edges
constexpr int kSmStd = kSmEdge + 32 * 4;                   // end of the standard-mel layout
constexpr int kSmMelIdx = kSmStd;                          // generic mel only: int start/len/off [3][128], group_begin[9] (+pad)
constexpr int kSmMelW = kSmMelIdx + (3 * kMaxMel + 12) * 4;  // generic mel only: float mel_w[nnz]
static_assert(kSmE % 16 == 0, "exchange rows hold 16-byte chunks");
static_assert(kSmRaw % 16 == 0 && kSmRawF32 % 16 == 0 && kSmTwA % 16 == 0 && kSmTwU % 16 == 0 && kSmMelW % 16 == 0 &&
              kSmDesc % 16 == 0 && kSmAcc % 16 == 0, "align");
static_assert(kSmRawF32 + 673 * 32 <= kSmRaw, "fp32 raw prefetch must fit behind the power tile");
static_assert(kRsPieces * 16 <= 11936 && 673 * 16 <= 11936, "int16 raw prefetch buffer too small");
static_assert(kRsMaxIn * 4 <= 16 * kGroupX, "fp32 resampler input must fit the exchange area");
static_assert(32 * (kMaxMel + 1) * 4 <= 5376 * 4, "out tile must fit in the staging area");
static_assert(kSmStd + 1024 <= 116224, "two CTAs per SM");

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
// L2 residency hints (two-phase pipeline): the raw log-mel rows are read again by the completion kernel right after this
// one -> evict_last stores; the waveform is read once -> evict_first loads (prefetch_tile).  Together they keep the raw
// rows in L2 between the two kernels (measured with single-metric ncu passes, profiles/r02_step_traffic.md).
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_normal() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_f4_hint(float4* p, float4 v, unsigned long long policy) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy) : "memory");
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// 16-byte cp.async of data that is read exactly once (the waveform), with an L2 eviction policy
__device__ __forceinline__ void cp_async16_stream(void* smem_dst, const void* gmem_src, int src_bytes, unsigned long long policy) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;\n" ::"r"(d), "l"(gmem_src), "r"(src_bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Asynchronously stages the raw samples [s_base - 8, s_base + 5376) of one utterance into shared memory,
// zero-filling everything outside [0, wlen) (cp.async src-size form), so the waveform's HBM latency
// overlaps the previous tile's FFT.  Chunk j of the buffer holds samples s_base - 8 + 8 j.
// Piece q holds utterance samples in_first + (8 | 4) q ...; pieces are 8-sample aligned, so none straddles 0.
template <bool kF32>
__device__ __forceinline__ void prefetch_tile(unsigned char* raw, const void* wav, const TileDesc* d, int tid) {
    const int in_first = d->in_first, in_len = d->in_len;
    // the waveform is read once: L2 evict_first, so the PCM stream does not push the raw log-mel rows out of L2 before
    // the completion kernel re-reads them (single-metric ncu passes: its DRAM reads fall from 14.7 MB to 2.3 MB per
    // batch, the step's DRAM traffic from 1.6x to 1.28x the algorithmic bytes)
    const unsigned long long pol = l2_policy_evict_first();
    if (kF32) {
        const float* w = reinterpret_cast<const float*>(wav) + d->wav_utt;
        for (int q = tid; q < 673 * 2; q += kThreads) {
            const int s = in_first + 4 * q;
            int valid = in_len - s;
            valid = valid < 0 ? 0 : (valid > 4 ? 4 : valid);
            if (s < 0) valid = 0;
            cp_async16_stream(raw + 16 * q, valid ? (const void*)(w + s) : (const void*)w, 4 * valid, pol);
        }
    } else {
        const int16_t* w = reinterpret_cast<const int16_t*>(wav) + d->wav_utt;
        const int pieces = d->rs ? kRsPieces : 673;
        if (in_first >= 0 && in_first + 8 * pieces <= in_len) {       // interior tile: every piece is whole
            const int16_t* const src = w + in_first;
            for (int q = tid; q < pieces; q += kThreads) cp_async16_stream(raw + 16 * q, src + 8 * q, 16, pol);
            return;
        }
        for (int q = tid; q < pieces; q += kThreads) {
            const int s = in_first + 8 * q;
            int valid = in_len - s;
            valid = valid < 0 ? 0 : (valid > 8 ? 8 : valid);
            if (s < 0) valid = 0;
            cp_async16_stream(raw + 16 * q, valid ? (const void*)(w + s) : (const void*)w, 2 * valid, pol);
        }
    }
}

// One polyphase block (10 outputs) of the fused speed perturb; taps the hann window zeroes are pruned at
// compile time (same predicate as oe_resample_fast_kernel), coefficients are kernel-parameter constants.
OE_CX bool rs_tap_nonzero(int orig, int neu, int width, int p, int q) {
    const double base = (orig < neu ? orig : neu) * 0.99;
    const double t = (-(double)p / neu + (double)(q - width) / orig) * base;
    return t < 6.0 && t > -6.0;
}

template <int ORIG>
__device__ __forceinline__ void rs_block(const float* __restrict__ xin, const float (&coef)[256], float (&y)[10]) {
    constexpr int TAPS = 14 + ORIG;
    float x[TAPS];
#pragma unroll
    for (int q = 0; q < TAPS; ++q) x[q] = xin[q];
    static_for<0, 10>([&](auto pp) {
        constexpr int p = decltype(pp)::value;
        float acc = 0.f;
        static_for<0, TAPS>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            if constexpr (rs_tap_nonzero(ORIG, 10, 7, p, q)) acc = fmaf(coef[p * TAPS + q], x[q], acc);
        });
        y[p] = acc;
    });
}

// rs_block with the baked tables of oe_rs_coefs.h: every coefficient is a 32-bit immediate of its FFMA (a run-time
// table costs a uniform constant load per two taps).  Same taps, same order: bitwise identical to rs_block on a
// table that matches the baked bits (oe_add_resampler only fuses such tables).
template <int ORIG>
__device__ __forceinline__ void rs_block_baked(const float* __restrict__ xin, float (&y)[10]) {
    constexpr int TAPS = 14 + ORIG;
    float x[TAPS];
#pragma unroll
    for (int q = 0; q < TAPS; ++q) x[q] = xin[q];
    static_for<0, 10>([&](auto pp) {
        constexpr int p = decltype(pp)::value;
        float acc = 0.f;
        static_for<0, TAPS>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            if constexpr (rs_tap_nonzero(ORIG, 10, 7, p, q)) {
                constexpr float c = rsbaked::coef<ORIG>(p * TAPS + q);
                acc = fmaf(c, x[q], acc);
            }
        });
        y[p] = acc;
    });
}

__device__ __forceinline__ void prefetch_desc(TileDesc* dst, const TileDesc* src, int tid) {
    if (tid < 3) cp_async16(reinterpret_cast<unsigned char*>(dst) + 16 * tid,
                            reinterpret_cast<const unsigned char*>(src) + 16 * tid, 16);
}

// ln(x) for x >= the log floor (a normal number): lg2.approx.ftz (max abs error 2^-22.6 on the mantissa's
// log2, i.e. ~1.2e-7 absolute on ln) without the denormal guard __logf carries.
__device__ __forceinline__ float fast_ln(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y * 0.693147180559945309f;
}

// Standard-structure mel projection of one warp's bin group: every index is a compile-time constant,
// the weights are kernel-parameter constants, repeated power-spectrum loads are CSE'd by the compiler.
template <int G>
__device__ __forceinline__ void mel_group_std(const float* __restrict__ pcol, const FbankParams& P,
                                              float* __restrict__ orow, float log_floor) {
    constexpr int b0 = mel80::kGroup[G], b1 = mel80::kGroup[G + 1];
    float acc[b1 - b0];
    static_for<b0, b1>([&](auto bb) {
        constexpr int b = decltype(bb)::value;
        constexpr int k0 = mel80::kStart[b], off = mel80::kOff[b];
        float a = 0.f;
        static_for<0, mel80::kLen[b]>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            a = fmaf(P.mel_w[off + i], pcol[(k0 + i) * kRowP], a);
        });
        acc[b - b0] = a;
    });
    static_for<b0, b1>([&](auto bb) {
        constexpr int b = decltype(bb)::value;
        orow[b] = fast_ln(fmaxf(acc[b - b0], log_floor));
    });
}

// kRs: the batch contains utterances with fused speed perturb (int16 input only); compiled out otherwise to keep
// the hot loop small (the kernel is instruction-cache sensitive).
template <bool kF32, bool kStdMel, bool kRs>
__global__ void __launch_bounds__(kThreads, 2) oe_fbank_kernel(const FbankParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* const sp = reinterpret_cast<float*>(smem + kSmP);
    float* const s8 = reinterpret_cast<float*>(smem + kSmS8);
    float* const sPw = reinterpret_cast<float*>(smem + kSmE);
    unsigned char* const sRaw = smem + (kF32 ? kSmRawF32 : kSmRaw);
    float2* const sTwA = reinterpret_cast<float2*>(smem + kSmTwA);
    float2* const sTwU = reinterpret_cast<float2*>(smem + kSmTwU);
    unsigned char* const sRowMask = smem + kSmMask;
    unsigned char* const sColMask = sRowMask + 32;
    int* const sMelStart = reinterpret_cast<int*>(smem + kSmMelIdx);
    int* const sMelLen = sMelStart + kMaxMel;
    int* const sMelOff = sMelLen + kMaxMel;
    int* const sGroup = sMelOff + kMaxMel;
    float* const sMelW = reinterpret_cast<float*>(smem + kSmMelW);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tau = tid & 15, grp = tid >> 4;
    const DevTables* __restrict__ tab = P.tab;
    const int F = kStdMel ? mel80::kBins : tab->n_mel;
    const int rowO = F + 1;

    TileDesc* const sDesc = reinterpret_cast<TileDesc*>(smem + kSmDesc);
    double* const sAcc = reinterpret_cast<double*>(smem + kSmAcc);
    if (P.cta_stats != nullptr)
        for (int i = tid; i < 3 * 2 * kMaxMel; i += kThreads) sAcc[i] = 0.0;

    grid_dep_wait();
    // ---- first tile: descriptor, then its waveform starts moving before anything else ----
    int tile = blockIdx.x;
    int slot = 0;
    if (tile < P.total_tiles) {
        prefetch_desc(sDesc, P.tiles + tile, tid);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        if (sDesc[0].nvalid > 0) prefetch_tile<kF32>(sRaw, P.wav, sDesc, tid);
        cp_async_commit();
    }
    // ---- one-time table staging ----
    for (int i = tid; i < 16 * kRowE; i += kThreads) sTwA[i] = tab->twA[i];
    for (int i = tid; i < 16 * kRowE; i += kThreads) sTwU[i] = tab->twU[i];
    if (!kStdMel) {
        for (int i = tid; i < kMaxMel; i += kThreads) {
            sMelStart[i] = tab->mel_start[i];
            sMelLen[i] = tab->mel_len[i];
            sMelOff[i] = tab->mel_off[i];
        }
        if (tid < 9) sGroup[tid] = tab->group_begin[tid];
        for (int i = tid; i < tab->nnz; i += kThreads) sMelW[i] = tab->mel_w[i];
    }
    float wv0[13], wv1[13];                    // window taps of this lane: w[32 n1 + 2 tau (+1)]
#pragma unroll
    for (int n1 = 0; n1 < 13; ++n1) {
        wv0[n1] = tab->window[32 * n1 + 2 * tau];
        wv1[n1] = tab->window[32 * n1 + 2 * tau + 1];
    }
    const float preemph = tab->preemph;
    const float dc_coef = 1.0f - preemph;
    const float log_floor = tab->log_floor;
    const bool fused = (P.n_tmask | P.n_fmask) != 0;

    for (; tile < P.total_tiles; tile += gridDim.x, slot ^= 1) {
        const int next = tile + gridDim.x;
        cp_async_wait_all();
        __syncthreads();               // (1) raw samples + descriptor visible; previous tile's rows are out of smem
        const TileDesc* const dp = sDesc + slot;
        const int b = dp->b, t0 = dp->t0;
        const int nvalid = dp->nvalid;                             // <= 0: padding-only tile
        const int rows_here = dp->rows_here;
        const long long out_start = dp->out_start;
        if (next < P.total_tiles) prefetch_desc(sDesc + (slot ^ 1), P.tiles + next, tid);
        cp_async_commit();
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
            // ---- [fused speed perturb: raw -> fp32 -> polyphase sinc -> resampled tile in sp] ----
            const int rs = (kF32 || !kRs) ? 0 : dp->rs;
            float* const sEdge = reinterpret_cast<float*>(smem + kSmEdge);
            if (kRs && rs != 0) {
                float* const xin = sPw;                               // fp32 input window, front of the (idle) E area
                const int16_t* const r16 = reinterpret_cast<const int16_t*>(sRaw);
                for (int i = tid; i < kRsPieces * 2; i += kThreads) {       // 4 samples per thread-iteration
                    const int2 v = *reinterpret_cast<const int2*>(r16 + 4 * i);
                    float4 f;
                    f.x = (float)(int16_t)(v.x & 0xffff); f.y = (float)(v.x >> 16);
                    f.z = (float)(int16_t)(v.y & 0xffff); f.w = (float)(v.y >> 16);
                    *reinterpret_cast<float4*>(xin + 4 * i) = f;
                }
                __syncthreads();
                const int orig = rs_orig(rs);
                const int shift = rs_first_input(rs, t0) - dp->in_first;   // 0..7
                for (int mi = tid; mi < kRsBlocks; mi += kThreads) {       // block mi -> tile samples 10 mi - 10 .. 10 mi - 1
                    float y[10];
                    if (rs == 1) rs_block<9>(xin + shift + mi * 9, P.rs_coef[0], y);
                    else rs_block<11>(xin + shift + mi * 11, P.rs_coef[1], y);
                    if (mi == 0) {
                        sEdge[0] = y[9];                                   // sample just before the tile
                    } else {
                        float2* const dst = reinterpret_cast<float2*>(sp + 10 * mi - 10);
#pragma unroll
                        for (int j = 0; j < 5; ++j)
                            if (10 * mi - 10 + 2 * j < 5376) dst[j] = make_float2(y[2 * j], y[2 * j + 1]);
                    }
                }
                __syncthreads();
                if (tid >= 1 && tid < 21) sEdge[tid] = sp[256 * tid - 1];  // block edges, read before the in-place pass
                __syncthreads();
            }
            // ---- convert: p[j] = x[j] - preemph * x[j-1] (kaldi.py:193-198), 8-sample block sums ----
            {
                const bool utt_start = t0 == 0;
                for (int blk = warp; blk < 21; blk += 8) {
                    const int i0 = blk * 256 + 4 * lane, i1 = i0 + 128;
                    float xa[4], xb[4], edge;
                    if (kRs && rs != 0) {                                  // resampled tile, in place in sp
                        const float4 a = *reinterpret_cast<const float4*>(sp + i0);
                        const float4 c = *reinterpret_cast<const float4*>(sp + i1);
                        xa[0] = a.x; xa[1] = a.y; xa[2] = a.z; xa[3] = a.w;
                        xb[0] = c.x; xb[1] = c.y; xb[2] = c.z; xb[3] = c.w;
                        edge = sEdge[blk];
                    } else if (kF32) {
                        const float* r = reinterpret_cast<const float*>(sRaw) + 8;
                        const float4 a = *reinterpret_cast<const float4*>(r + i0);
                        const float4 c = *reinterpret_cast<const float4*>(r + i1);
                        xa[0] = a.x; xa[1] = a.y; xa[2] = a.z; xa[3] = a.w;
                        xb[0] = c.x; xb[1] = c.y; xb[2] = c.z; xb[3] = c.w;
                        edge = r[blk * 256 - 1];
                    } else {
                        const int16_t* r = reinterpret_cast<const int16_t*>(sRaw) + 8;
                        const int2 a = *reinterpret_cast<const int2*>(r + i0);
                        const int2 c = *reinterpret_cast<const int2*>(r + i1);
                        xa[0] = (float)(int16_t)(a.x & 0xffff); xa[1] = (float)(a.x >> 16);
                        xa[2] = (float)(int16_t)(a.y & 0xffff); xa[3] = (float)(a.y >> 16);
                        xb[0] = (float)(int16_t)(c.x & 0xffff); xb[1] = (float)(c.x >> 16);
                        xb[2] = (float)(int16_t)(c.y & 0xffff); xb[3] = (float)(c.y >> 16);
                        edge = (float)r[blk * 256 - 1];
                    }
                    const float upa = __shfl_sync(0xffffffffu, xa[3], (lane + 31) & 31);
                    const float upb = __shfl_sync(0xffffffffu, xb[3], (lane + 31) & 31);
                    const float pva = lane ? upa : ((utt_start && blk == 0) ? xa[0] : edge);
                    const float pvb = lane ? upb : upa;
                    float4 pa, pb;
                    pa.x = fmaf(-preemph, pva, xa[0]);
                    pa.y = fmaf(-preemph, xa[0], xa[1]);
                    pa.z = fmaf(-preemph, xa[1], xa[2]);
                    pa.w = fmaf(-preemph, xa[2], xa[3]);
                    pb.x = fmaf(-preemph, pvb, xb[0]);
                    pb.y = fmaf(-preemph, xb[0], xb[1]);
                    pb.z = fmaf(-preemph, xb[1], xb[2]);
                    pb.w = fmaf(-preemph, xb[2], xb[3]);
                    __syncwarp();                                          // in-place variant: all lanes have read
                    *reinterpret_cast<float4*>(sp + i0) = pa;
                    *reinterpret_cast<float4*>(sp + i1) = pb;
                    float sa = (xa[0] + xa[1]) + (xa[2] + xa[3]);
                    float sb = (xb[0] + xb[1]) + (xb[2] + xb[3]);
                    sa += __shfl_xor_sync(0xffffffffu, sa, 1);
                    sb += __shfl_xor_sync(0xffffffffu, sb, 1);
                    if (!(lane & 1)) {
                        s8[blk * 32 + (lane >> 1)] = sa;
                        s8[blk * 32 + 16 + (lane >> 1)] = sb;
                    }
                }
            }
            if (!kF32) cp_async_wait_all();    // next tile's descriptor has landed (this thread's pieces; (2) publishes them)
            __syncthreads();                                       // (2)
            if (!kF32) {               // int16: the raw buffer has been consumed -> next tile's samples start moving now
                if (next < P.total_tiles) {
                    const TileDesc* const dn = sDesc + (slot ^ 1);
                    if (dn->nvalid > 0) prefetch_tile<kF32>(sRaw, P.wav, dn, tid);
                }
                cp_async_commit();
            }
            // ---- stage A: frame means (kaldi.py:183-186), one packed 16-point FFT of z[16 n1 + tau] for 2 frames ----
            unsigned char* const gx = smem + kSmE + grp * kGroupX;   // this group's exchange area: re plane, im plane (+2048)
            {
                V2 zr[16], zi[16];
                const float* q0 = s8 + 20 * (2 * grp) + tau;
                float m0 = (q0[0] + q0[16]) + q0[32], m1 = (q0[20] + q0[36]) + q0[52];
                if (tau < 2) {
                    m0 += q0[48];
                    m1 += q0[68];
                }
#pragma unroll
                for (int o = 8; o >= 1; o >>= 1) {
                    m0 += __shfl_xor_sync(0xffffffffu, m0, o);
                    m1 += __shfl_xor_sync(0xffffffffu, m1, o);
                }
                const float c0 = dc_coef * (m0 / (float)kWin), c1 = dc_coef * (m1 / (float)kWin);
                const float* base = sp + kShift * (2 * grp) + 2 * tau;
#pragma unroll
                for (int n1 = 0; n1 < 13; ++n1) {
                    const float2 v0 = *reinterpret_cast<const float2*>(base + 32 * n1);
                    const float2 v1 = *reinterpret_cast<const float2*>(base + kShift + 32 * n1);
                    zr[n1] = v2_make((v0.x - c0) * wv0[n1], (v1.x - c1) * wv0[n1]);
                    zi[n1] = v2_make((v0.y - c0) * wv1[n1], (v1.y - c1) * wv1[n1]);
                }
#pragma unroll
                for (int n1 = 13; n1 < 16; ++n1) zr[n1] = zi[n1] = vbcast(0.f);
                fft_dif<16, 13, V2>(zr, zi);
                // twiddle by W256^(tau k1) and store row k1: cell tau of 8 B (frame pair), 16-byte chunks
                // XOR-swizzled by k1 & 7 so that the row reads (LDS.128, one row per lane) are conflict-free
                unsigned off8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) off8[j] = (unsigned)((((tau >> 1) ^ j) << 4) + ((tau & 1) << 3));
                const float4* const tw4 = reinterpret_cast<const float4*>(sTwA + tau * kRowE);
                static_for<0, 8>([&](auto ii) {
                    constexpr int i = decltype(ii)::value;
                    const float4 t = tw4[i];                       // k1 = 2i: (t.x, t.y), 2i+1: (t.z, t.w); (cos, -sin)
                    constexpr int p0 = bitrev<16>(2 * i), p1 = bitrev<16>(2 * i + 1);
                    V2 r0 = zr[p0], i0 = zi[p0];
                    if constexpr (i != 0) {
                        const V2 c = vbcast(t.x), sn = vbcast(t.y);
                        r0 = vfma(zi[p0], vbcast(-t.y), vmul(zr[p0], c));
                        i0 = vfma(zr[p0], sn, vmul(zi[p0], c));
                    }
                    const V2 c1v = vbcast(t.z), s1v = vbcast(t.w);
                    const V2 r1 = vfma(zi[p1], vbcast(-t.w), vmul(zr[p1], c1v));
                    const V2 i1 = vfma(zr[p1], s1v, vmul(zi[p1], c1v));
                    unsigned char* const d0 = gx + (2 * i) * 128 + off8[(2 * i) & 7];
                    unsigned char* const d1 = gx + (2 * i + 1) * 128 + off8[(2 * i + 1) & 7];
                    *reinterpret_cast<V2*>(d0) = r0;
                    *reinterpret_cast<V2*>(d0 + 2048) = i0;
                    *reinterpret_cast<V2*>(d1) = r1;
                    *reinterpret_cast<V2*>(d1 + 2048) = i1;
                });
            }
            __syncwarp();

            // ---- stage B: lane k1 = tau transforms row k1 of both frames; rows exchanged; untangle ----
            V2 pk[16];
            {
                V2 ar[16], ai[16];
                {
                    const unsigned char* const row = gx + tau * 128;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 vr = *reinterpret_cast<const uint4*>(row + ((c ^ (tau & 7)) << 4));
                        const uint4 vi = *reinterpret_cast<const uint4*>(row + 2048 + ((c ^ (tau & 7)) << 4));
                        ar[2 * c] = v2_make(__uint_as_float(vr.x), __uint_as_float(vr.y));
                        ar[2 * c + 1] = v2_make(__uint_as_float(vr.z), __uint_as_float(vr.w));
                        ai[2 * c] = v2_make(__uint_as_float(vi.x), __uint_as_float(vi.y));
                        ai[2 * c + 1] = v2_make(__uint_as_float(vi.z), __uint_as_float(vi.w));
                    }
                }
                __syncwarp();                                      // every lane has its row: the area is reused for Z
                fft_dif<16, 16, V2>(ar, ai);                       // position p holds Z[tau + 16 bitrev(p)]
                // Z[k1][k2] -> slot (k2 + k1) mod 16 of row k1 (8 B cells, additive swizzle: conflict-free both ways)
                {
                    unsigned char* const wrow = gx + tau * 128;
                    const unsigned wk = (unsigned)tau << 3;
                    static_for<0, 16>([&](auto pp) {
                        constexpr int pos = decltype(pp)::value;
                        constexpr int k2 = bitrev<16>(pos);
                        unsigned char* const d = wrow + (((k2 << 3) + wk) & 120);
                        *reinterpret_cast<V2*>(d) = ar[pos];
                        *reinterpret_cast<V2*>(d + 2048) = ai[pos];
                    });
                }
                __syncwarp();
                // conjugate partner Z[256 - k]: row (16 - k1) mod 16, index 15 - k2 (k1 >= 1) or (16 - k2) mod 16 (k1 == 0)
                {
                    const int prow = (16 - tau) & 15;
                    const unsigned char* const rrow = gx + prow * 128;
                    const unsigned rk = (unsigned)((15 + (tau == 0 ? 1 : 0) + prow) << 3);     // slot = (K - k2) mod 16
                    const float4* const tw4 = reinterpret_cast<const float4*>(sTwU + tau * kRowE);
                    static_for<0, 8>([&](auto ii) {
                        constexpr int i = decltype(ii)::value;
                        const float4 t = tw4[i];                   // k = tau + 16 k2: k2 = 2i -> (t.x, t.y), 2i+1 -> (t.z, t.w)
                        {
                            constexpr int k2 = 2 * i, pos = bitrev<16>(k2);
                            const unsigned char* const q = rrow + ((rk - (k2 << 3)) & 120);
                            const V2 qr = *reinterpret_cast<const V2*>(q), qi = *reinterpret_cast<const V2*>(q + 2048);
                            pk[k2] = untangle_power<V2>(ar[pos], ai[pos], qr, qi, vbcast(t.x), vbcast(t.y));
                        }
                        {
                            constexpr int k2 = 2 * i + 1, pos = bitrev<16>(k2);
                            const unsigned char* const q = rrow + ((rk - (k2 << 3)) & 120);
                            const V2 qr = *reinterpret_cast<const V2*>(q), qi = *reinterpret_cast<const V2*>(q + 2048);
                            pk[k2] = untangle_power<V2>(ar[pos], ai[pos], qr, qi, vbcast(t.z), vbcast(t.w));
                        }
                    });
                }
            }
            if (kF32) cp_async_wait_all();   // fp32 input: the next descriptor is needed right after (3)
            __syncthreads();           // (3) every group is done with its exchange area -> power tile [+ fp32 raw prefetch]
            if (kF32) {
                if (next < P.total_tiles) {
                    const TileDesc* const dn = sDesc + (slot ^ 1);
                    if (dn->nvalid > 0) prefetch_tile<kF32>(sRaw, P.wav, dn, tid);
                }
                cp_async_commit();
            }
            {
                float* const prow = sPw + tau * kRowP + 2 * grp;   // bin k = tau + 16 k2, columns (2 grp, 2 grp + 1)
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) *reinterpret_cast<V2*>(prow + 16 * k2 * kRowP) = pk[k2];
            }
            __syncthreads();                                       // (4) power tile complete

            // ---- sparse mel + log: warp = mel-bin group, lane = frame ----
            if (kStdMel) {
                const float* const pcol = sPw + lane;
                float* const orow = sp + lane * rowO;
                switch (warp) {
                    case 0: mel_group_std<0>(pcol, P, orow, log_floor); break;
                    case 1: mel_group_std<1>(pcol, P, orow, log_floor); break;
                    case 2: mel_group_std<2>(pcol, P, orow, log_floor); break;
                    case 3: mel_group_std<3>(pcol, P, orow, log_floor); break;
                    case 4: mel_group_std<4>(pcol, P, orow, log_floor); break;
                    case 5: mel_group_std<5>(pcol, P, orow, log_floor); break;
                    case 6: mel_group_std<6>(pcol, P, orow, log_floor); break;
                    default: mel_group_std<7>(pcol, P, orow, log_floor); break;
                }
            } else {
                const float* const pcol = sPw + lane;
                for (int bin = sGroup[warp]; bin < sGroup[warp + 1]; ++bin) {
                    const int k0 = sMelStart[bin], len = sMelLen[bin];
                    const float* const w = sMelW + sMelOff[bin];
                    float acc = 0.f;
                    for (int i = 0; i < len; ++i) acc = fmaf(w[i], pcol[(k0 + i) * kRowP], acc);
                    sp[lane * rowO + bin] = fast_ln(fmaxf(acc, log_floor));
                }
            }
            __syncthreads();                                       // (5) output tile complete
            if (P.tile_stats != nullptr || P.cta_stats != nullptr) {
                for (int idx = tid; idx < 3 * F; idx += kThreads) {
                    const int rg = idx / F, f = idx - rg * F;
                    const int n = stats_rows(nvalid, rg);
                    float s = 0.f, m2 = 0.f;
                    if (n > 0) {
                        const float* col = sp + (11 * rg) * rowO + f;
                        for (int r = 0; r < n; ++r) s += col[r * rowO];
                        const float mean = s / (float)n;
                        for (int r = 0; r < n; ++r) {
                            const float d = col[r * rowO] - mean;
                            m2 = fmaf(d, d, m2);
                        }
                        if (P.cta_stats != nullptr) {          // sum x^2 = M2 + n mean^2, accumulated in fp64, fixed order
                            sAcc[(rg * 2 + 0) * kMaxMel + f] += (double)s;
                            sAcc[(rg * 2 + 1) * kMaxMel + f] += (double)m2 + (double)s * (double)mean;
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
            cp_async_wait_all();
            __syncthreads();
            if (next < P.total_tiles) {
                const TileDesc* const dn = sDesc + (slot ^ 1);
                if (dn->nvalid > 0) prefetch_tile<kF32>(sRaw, P.wav, dn, tid);
            }
            cp_async_commit();
        }

        // ---- rows out: [mask] -> [CMVN] -> coalesced stores; padding rows are 0 / (0-mean)*istd ----
        if (P.out != nullptr) {
            const bool has_cmvn = P.cmvn_mean != nullptr;
            const int pitch = (int)P.pitch;
            float* const dst0 = P.out + out_start * P.pitch;
            if (kStdMel && P.out_vec && !fused && !has_cmvn) {
                // the tile is one contiguous run of rows_here * 80 floats: flat float4 stores
                constexpr int kQ = mel80::kBins / 4;                  // float4 per row
                const int nq = rows_here * kQ;
                float4* const dst4 = reinterpret_cast<float4*>(dst0);
                for (int c = tid; c < nq; c += kThreads) {
                    const int r = c / kQ, col = (c - r * kQ) * 4;
                    const float* const src = sp + r * rowO + col;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < nvalid) v = make_float4(src[0], src[1], src[2], src[3]);
                    dst4[c] = v;
                }
            } else if (!fused && !has_cmvn) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = warp + 8 * i;
                    if (r < rows_here) {
                        const bool real = r < nvalid;
                        const float* const srow = sp + r * rowO;
                        float* const dst = dst0 + r * pitch;
                        for (int f = lane; f < F; f += 32) dst[f] = real ? srow[f] : 0.f;
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
                        const float* const srow = sp + r * rowO;
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
    }
    cp_async_wait_all();
    if (P.cta_stats != nullptr) {      // each accumulator is owned by one thread: no barrier needed
        for (int idx = tid; idx < 3 * F; idx += kThreads) {
            const int rg = idx / F, f = idx - rg * F;
            double* const dst = P.cta_stats + ((int64_t)blockIdx.x * 3 + rg) * 2 * F;
            dst[f] = sAcc[(rg * 2 + 0) * kMaxMel + f];
            dst[F + f] = sAcc[(rg * 2 + 1) * kMaxMel + f];
        }
    }
}

}  // namespace oe
