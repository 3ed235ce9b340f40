// B200 (sm_100a) acoustic front-end kernels + the C ABI declared in include/openeat_frontend.h.
//
// Reference behaviour being replaced (see DESIGN.md for the full map):
//   openeat/dataset/dataset.py:39-118, 155-239     _extract_feature / audio_collate_func
//   torchaudio/compliance/kaldi.py:514-645          fbank (third-party, called at dataset.py:93-100)
//   openeat/dataset/feature_processor.py:5-64       _normalization / _spec_augmentation / _spec_substitute
//   openeat/dataset/audio_processor.py:19-35        _speed_perturb
//   openeat/modules/cmvn.py:35-46                   GlobalCMVN.forward
//
// Kernel inventory
//   oe_fbank_kernel     ragged batch, one CTA per 32-frame tile, persistent grid:
//                       stage waveform -> smem (pre-emphasis folded in, block sums for DC removal)
//                       -> per-frame 512-pt real FFT as 256-pt complex FFT split 16x16 over 16 threads
//                       (registers + half-warp smem exchange) -> power -> sparse mel -> log
//                       -> [mask, CMVN] -> coalesced rows; optional per-tile column statistics.
//   oe_utt_stats_kernel per-utterance mean/std from tile statistics (Chan merge, fp64).
//   oe_finalize_kernel  per-utt normalisation + spec_sub gather + spec_aug masks + CMVN + padding.
//   oe_global_stats_kernel  += sum / sumsq / count for compute_cmvn_stats.
//   oe_cmvn_kernel      GlobalCMVN.forward.
//   oe_resample_kernel  polyphase sinc resampler (speed perturb).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/openeat_frontend.h"
#include "oe_fft.h"
#include "oe_ingest.h"

namespace oe {

constexpr int kWin = 400;           // frame_length
constexpr int kShift = 160;         // frame_shift
constexpr int kFft = 512;
constexpr int kBins = 256;          // fft bins carrying mel weight (kaldi.py:627 pads Nyquist with 0)
constexpr int kTileFrames = 32;
constexpr int kThreads = 256;
constexpr int kTileSamples = kShift * (kTileFrames - 1) + kWin;   // 5360
constexpr int kChunks = 672;        // 8-sample chunks staged per tile (5376 samples, 16 of slack)
constexpr int kMaxMel = 128;
constexpr int kMaxNnz = 2048;
constexpr int kRowE = 18;           // complex per exchange row: 16 + 2 pad -> 144 B (conflict-free LDS.128)
constexpr int kRowO = 81;           // floats per output-tile row (80 + 1 pad); generic: F + 1

// Programmatic dependent launch: blocks until the preceding grid in the stream has completed and its writes are
// visible (returns immediately when the kernel was launched without the attribute).
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the next grid of the stream be scheduled now (it still waits for this grid's completion in its own
// grid_dep_wait before reading anything): its blocks become resident and run their prologues under this grid.
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct DevTables {
    float window[kFft];             // zero beyond frame_length
    float2 twA[16 * kRowE];         // W256^(tau*k1) = (cos, -sin)
    float2 twU[16 * kRowE];         // (cos, sin)(2 pi k / 512), k = k1 + 16*k2, row k1
    int mel_start[kMaxMel];
    int mel_len[kMaxMel];
    int mel_off[kMaxMel];
    int group_begin[9];
    int n_mel;
    int nnz;
    float log_floor;
    float preemph;
    float mel_w[kMaxNnz];           // already multiplied by 1/4 (see untangle_power)
};

// 1 / rows for rows = 0..11, correctly rounded (rows <= 11: one fp32 rounding on a partial mean)
__constant__ float kInvRows[12] = {0.f, 1.f, 1.f / 2.f, 1.f / 3.f, 1.f / 4.f, 1.f / 5.f, 1.f / 6.f,
                                   1.f / 7.f, 1.f / 8.f, 1.f / 9.f, 1.f / 10.f, 1.f / 11.f};

}  // namespace oe

#include "oe_mel80.h"
#include "oe_rs_coefs.h"
#include "oe_fbank_kernel.cuh"
#include "oe_fbank2_kernel.cuh"
#include "oe_flac_gpu.cuh"

namespace oe {

// ------------------------------------------------------------------------------------------
struct TileDescParams {
    const int32_t* tile_prefix;  // [B+1]
    const int64_t* wav_off;
    const int32_t* wav_len;
    const int32_t* n_frames;
    const int32_t* n_rows;
    const int64_t* out_row;
    const int32_t* rs_mode;      // [B] fused speed perturb: 0 none, 1 = 9:10, 2 = 11:10
    TileDesc* tiles;
    int B, total_tiles;
    int feats;                   // the input already is features: offsets / lengths count rows
    // per-call state of the gen-2 kernel's in-kernel CMVN statistics, reset here (stream-ordered before the fbank kernel)
    int32_t* sched;              // [1] or null: CTAs done
    unsigned long long* stat_acc;    // [n_acc] or null
    int n_acc;
};

// Expands the per-utterance metadata into one self-contained descriptor per 32-frame tile.
constexpr int kDescSmemUtts = 4096;
__device__ __forceinline__ void tile_desc_body(const TileDescParams& P, int32_t* sh_prefix) {
    // the binary search runs on a shared-memory copy of the prefix array (one coalesced read instead of
    // log2(B) dependent global loads per thread)
    const bool staged = P.B <= kDescSmemUtts;
    if (staged) {
        for (int i = threadIdx.x; i <= P.B; i += blockDim.x) sh_prefix[i] = P.tile_prefix[i];
        __syncthreads();
    }
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    {
        const int stride = gridDim.x * blockDim.x;
        if (P.stat_acc != nullptr)
            for (int i = tile; i < P.n_acc; i += stride) P.stat_acc[i] = 0ull;
        if (P.sched != nullptr && tile == 0) P.sched[0] = P.sched[1] = 0;      // CTAs done; tiles claimed (gen-2 kernel)
    }
    if (tile >= P.total_tiles) return;
    const int b = staged ? find_utt(sh_prefix, P.B, tile) : find_utt(P.tile_prefix, P.B, tile);
    const int t0 = (tile - P.tile_prefix[b]) * kTileFrames;
    TileDesc d;
    d.b = b;
    d.t0 = t0;
    d.nvalid = min(kTileFrames, P.n_frames[b] - t0);
    d.rows_here = min(kTileFrames, P.n_rows[b] - t0);
    d.wav_utt = P.wav_off[b];
    d.in_len = P.wav_len[b];
    d.rs = P.feats ? 0 : P.rs_mode[b];
    if (P.feats) d.in_first = t0;
    else if (d.rs == 0) d.in_first = t0 * kShift - 8;
    else d.in_first = rs_first_input(d.rs, t0) & ~7;          // floor to a multiple of 8 (also for negatives)
    d.out_start = P.out_row[b] + t0;
    d.pad = 0;
    P.tiles[tile] = d;
}

__global__ void oe_tile_desc_kernel(const TileDescParams P) {
    __shared__ int32_t sh_prefix[kDescSmemUtts + 1];
    grid_dep_wait();                           // the metadata block (oe_fetch_kernel) has landed
    grid_dep_launch();                         // the fbank kernel's table staging runs under this kernel
    tile_desc_body(P, sh_prefix);
}

// Small batches: the whole metadata block travels INSIDE the kernel parameters (it is a few hundred bytes for a batch of
// 16 utterances), this kernel stores it into the workspace for the kernels behind it and builds the descriptors straight
// from the parameter bank -- no oe_fetch_kernel, one latency-bound launch less in front of a 10 us fbank kernel.
constexpr int kInlineMeta = 3072;
struct TileDescInlineParams {
    TileDescParams p;            // array pointers = byte offsets into `meta`
    uint4* ws_meta;              // workspace copy for the fbank / completion kernels
    int meta_bytes;
    __align__(16) unsigned char meta[kInlineMeta];
};
__global__ void oe_tile_desc_inline_kernel(const __grid_constant__ TileDescInlineParams Q) {
    __shared__ int32_t sh_prefix[kDescSmemUtts + 1];
    grid_dep_launch();
    const int n16 = (Q.meta_bytes + 15) / 16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x)
        Q.ws_meta[i] = reinterpret_cast<const uint4*>(Q.meta)[i];
    TileDescParams P = Q.p;
    const unsigned char* const m = Q.meta;
    P.tile_prefix = reinterpret_cast<const int32_t*>(m + reinterpret_cast<uintptr_t>(Q.p.tile_prefix));
    P.wav_off = reinterpret_cast<const int64_t*>(m + reinterpret_cast<uintptr_t>(Q.p.wav_off));
    P.wav_len = reinterpret_cast<const int32_t*>(m + reinterpret_cast<uintptr_t>(Q.p.wav_len));
    P.n_frames = reinterpret_cast<const int32_t*>(m + reinterpret_cast<uintptr_t>(Q.p.n_frames));
    P.n_rows = reinterpret_cast<const int32_t*>(m + reinterpret_cast<uintptr_t>(Q.p.n_rows));
    P.out_row = reinterpret_cast<const int64_t*>(m + reinterpret_cast<uintptr_t>(Q.p.out_row));
    P.rs_mode = reinterpret_cast<const int32_t*>(m + reinterpret_cast<uintptr_t>(Q.p.rs_mode));
    tile_desc_body(P, sh_prefix);
}

// Small host -> device transfers WITHOUT the copy engine: the source is pinned host memory mapped into the device's
// address space, the SMs read it over PCIe.  A cudaMemcpyAsync of the same bytes queues on the host-to-device DMA engine
// BEHIND whatever bulk copy is in flight there -- in the collate pipeline that is the next batch's 48 MB of PCM, i.e.
// the metadata of batch i (and with it every kernel of batch i) would wait ~0.9 ms for the PCM of batch i+1
// (measured: 1.15 ms per step instead of the 0.87 ms the PCM copy alone takes).
__global__ void __launch_bounds__(256) oe_fetch_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n16) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------
struct UttStatsParams {
    const float* tile_stats;     // null: no per-utterance statistics wanted (global role only)
    const int32_t* tile_prefix;
    const int32_t* n_frames;
    float* utt_mean;     // [B][F]
    float* utt_std;      // [B][F]
    int F;
    int B;
    // compute_cmvn_stats role of the blocks behind the B utterance blocks (gridDim.x = B + global_blocks)
    const double* partial;   // [n_partials][2F]: sum, sum of squares written by the producer CTAs
    double* stats;           // [2F+1] accumulated in place
    double count;
    int n_partials;
};

constexpr int kUttSlices = 8;
constexpr int kUttMaxPart = 12;          // partial sums per slice held in registers: covers 32 tiles = 1 024 frames per utterance
constexpr int kGlobStats = 8;            // statistics per global-role block

// Block role 1 (blockIdx.x < B), feature_processor.py:5-8: mean and population std over the frames of one utterance,
// merged from the per-tile partials in a fixed order (fp64): mean = S/N, M2 = sum_p [M2_p + n_p (mean_p - mean)^2]
// (two passes over the partials: robust for constant features, where a one-pass difference of sums would cancel).
// block = (F columns) x (kUttSlices slices of the partial list), combined through shared memory; the partial sums stay
// in registers between the passes; loads use clamped 32-bit indices (no branches).  The kernel is a latency chain
// (five dependent round trips plus an fp64 division and a square root), so it is sized for one wave: F <= 80 runs
// 640-thread blocks at <= 51 registers, two per SM.
// Block role 2 (blockIdx.x >= B), compute_cmvn_stats: += sum, sum of squares and frame count into the caller's
// accumulator.  The producers' partials are combined in a fixed order (strided sums per lane, then in lane order;
// the tile -> CTA assignment is static), so the result is bitwise reproducible run to run.
template <int MAX_THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(MAX_THREADS, MIN_BLOCKS) oe_utt_stats_kernel(const UttStatsParams P) {
    __shared__ double sh[kUttSlices][kMaxMel];
    const int f = threadIdx.x, y = threadIdx.y;
    grid_dep_launch();                         // the finalize blocks become resident while this latency chain runs
    grid_dep_wait();
    if ((int)blockIdx.x >= P.B) {
        const int nthr = blockDim.x * kUttSlices, lanes = nthr / kGlobStats;   // lanes per statistic
        const int t = y * blockDim.x + f, slot = t / lanes, l = t - slot * lanes;
        const int stat = ((int)blockIdx.x - P.B) * kGlobStats + slot;
        double* const flat = &sh[0][0];
        double s = 0.0;
        if (slot < kGlobStats && stat < 2 * P.F) {
            const double* __restrict__ src = P.partial + stat;
#pragma unroll 8
            for (int g = l; g < P.n_partials; g += lanes) s += src[(int64_t)g * 2 * P.F];
        }
        flat[t] = s;
        __syncthreads();
        if (l == 0 && slot < kGlobStats) {
            if (stat < 2 * P.F) {
                double tot = 0.0;
                for (int j = 0; j < lanes; ++j) tot += flat[t + j];
                P.stats[stat] += tot;
            } else if (stat == 2 * P.F) {
                P.stats[2 * P.F] += P.count;
            }
        }
        return;
    }
    const int b = blockIdx.x;
    const int nfr = P.n_frames[b];
    const int np = 3 * ((nfr + kTileFrames - 1) / kTileFrames);
    if (np == 0) {                                                  // dropped utterance (no frames): nothing reads these
        if (y == 0) P.utt_mean[(int64_t)b * P.F + f] = P.utt_std[(int64_t)b * P.F + f] = 0.f;
        return;
    }
    const int F2 = 2 * P.F;
    const float* __restrict__ base = P.tile_stats + (int64_t)P.tile_prefix[b] * 3 * F2 + f;
    float ps[kUttMaxPart];
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < kUttMaxPart; ++u) {
        const int i = y + u * kUttSlices;
        const float v = base[min(i, np - 1) * F2];
        ps[u] = i < np ? v : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kUttMaxPart; ++u) acc += (double)ps[u];
    for (int i = y + kUttMaxPart * kUttSlices; i < np; i += kUttSlices) acc += (double)base[i * F2];
    sh[y][f] = acc;
    __syncthreads();
    double S = 0.0;
#pragma unroll
    for (int j = 0; j < kUttSlices; ++j) S += sh[j][f];
    const double mean = S / (double)nfr;
    __syncthreads();
    acc = 0.0;
    const float* __restrict__ base2 = base + P.F;
#pragma unroll
    for (int u = 0; u < kUttMaxPart; ++u) {
        const int i = y + u * kUttSlices;
        const float m2p = base2[min(i, np - 1) * F2];
        const int tile = i / 3, rg = i - 3 * tile;
        const int rows = i < np ? stats_rows(min(kTileFrames, nfr - tile * kTileFrames), rg) : 0;
        const double d = (double)(ps[u] * kInvRows[rows]) - mean;
        acc += rows > 0 ? (double)m2p + (double)rows * d * d : 0.0;
    }
    for (int i = y + kUttMaxPart * kUttSlices; i < np; i += kUttSlices) {
        const int tile = i / 3, rg = i - 3 * tile;
        const int rows = stats_rows(min(kTileFrames, nfr - tile * kTileFrames), rg);
        const double d = (double)(base[i * F2] * kInvRows[rows]) - mean;
        acc += rows > 0 ? (double)base2[i * F2] + (double)rows * d * d : 0.0;
    }
    sh[y][f] = acc;
    __syncthreads();
    if (y == 0) {
        double m2 = 0.0;
#pragma unroll
        for (int j = 0; j < kUttSlices; ++j) m2 += sh[j][f];
        P.utt_mean[(int64_t)b * P.F + f] = (float)mean;
        P.utt_std[(int64_t)b * P.F + f] = (float)sqrt(m2 / (double)nfr);
    }
}

struct FeatStatsParams {
    const float* feats;          // ragged rows, pitch F
    const TileDesc* tiles;
    float* tile_stats;
    double* cta_stats;           // [gridDim.x][3][2][F] or null
    int F, total_tiles;
};

// Same per-tile column statistics as the fbank kernel's epilogue, for batches that arrive as
// features (data_type != 'wav', dataset.py:190-191, or the numpy-level processor mirrors).
__global__ void oe_feat_tile_stats_kernel(const FeatStatsParams P) {
    grid_dep_wait();
    const int f = threadIdx.x;
    double as[3] = {0.0, 0.0, 0.0}, aq[3] = {0.0, 0.0, 0.0};
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const TileDesc d = P.tiles[tile];
        if (f >= P.F || d.nvalid <= 0) continue;
#pragma unroll
        for (int rg = 0; rg < 3; ++rg) {
            const int n = stats_rows(d.nvalid, rg);
            const float* src = P.feats + (d.wav_utt + d.t0 + 11 * rg) * P.F + f;   // feats mode: offsets count rows
            float s = 0.f, m2 = 0.f;
            if (n > 0) {
                for (int r = 0; r < n; ++r) s += src[(int64_t)r * P.F];
                const float mean = s / (float)n;
                for (int r = 0; r < n; ++r) {
                    const float dd = src[(int64_t)r * P.F] - mean;
                    m2 = fmaf(dd, dd, m2);
                }
                as[rg] += (double)s;
                aq[rg] += (double)m2 + (double)s * (double)mean;
            }
            float* const st = P.tile_stats + ((int64_t)tile * 3 + rg) * 2 * P.F;
            st[f] = s;
            st[P.F + f] = m2;
        }
    }
    if (P.cta_stats != nullptr && f < P.F) {
#pragma unroll
        for (int rg = 0; rg < 3; ++rg) {
            double* const dst = P.cta_stats + ((int64_t)blockIdx.x * 3 + rg) * 2 * P.F;
            dst[f] = as[rg];
            dst[P.F + f] = aq[rg];
        }
    }
}

struct FinalizeParams {
    const float* raw;            // ragged raw log-mel, pitch F
    const int64_t* frame_prefix; // [B] first raw row of each utterance
    const int64_t* row_prefix;   // [B+1] output rows
    const int32_t* n_frames;
    const int64_t* out_row;
    float* out;
    int64_t pitch;
    const float* utt_mean;
    const float* utt_std;
    const int32_t* frame_map;
    const int64_t* map_off;
    const int32_t* tmask;
    const int32_t* fmask;
    int n_tmask, n_fmask;
    const float* cmvn_mean;
    const float* cmvn_istd;
    int cmvn_on_pad;
    int F;
    float dither_a;              // feature dither amplitude (dataset.py:199-201), 0 = off
    unsigned long long dither_seed;
};

// dataset.py:195-218 on the device: normalise -> substitute -> mask -> pad, then GlobalCMVN.
// grid = (utterances, row chunks).  VEC = 4: thread = (float4 column chunk, row lane); the per-column
// constants (mean, 1/std, CMVN, frequency mask) live in registers for the whole chunk of rows.  One row per thread
// in flight is deliberate: the raw rows are L2 hits, the kernel is bound by the HBM writes, and a 4-row unrolled
// variant (64 registers) measured 9 us slower in the warm pipeline (tools/ab_step.py) although ncu's cold-cache
// replay showed it faster.
constexpr int kFinRows = 128;       // rows per block
// DITHER is a template flag: the Philox code would otherwise raise the register count (40 -> 63) and cost the
// plain path a third of its occupancy.
template <int VEC, bool DITHER>
__global__ void __launch_bounds__(256) oe_finalize_kernel(const FinalizeParams P) {
    grid_dep_wait();
    const int b = blockIdx.x;
    const int r0 = blockIdx.y * kFinRows;
    const int nrows = (int)(P.row_prefix[b + 1] - P.row_prefix[b]);
    if (r0 >= nrows) return;
    const int F = P.F;
    const int ncol = (F + VEC - 1) / VEC;                 // column chunks per row
    const int lanes = 256 / ncol;                         // row lanes per block
    const int cc = threadIdx.x % ncol, rl = threadIdx.x / ncol;
    const int c = cc * VEC;
    const int nfr = P.n_frames[b];
    const bool norm = P.utt_mean != nullptr, has_cm = P.cmvn_mean != nullptr, has_ci = P.cmvn_istd != nullptr;
    // per-column constants: computed once per block by F threads, shared through shared memory (every thread loading
    // its own 16 constants and the mask ranges made the prologue 45 loads for 11 rows of work)
    __shared__ __align__(16) float sh_c[4][kMaxMel + 4];
    __shared__ __align__(4) unsigned char sh_m[kMaxMel + 4];
    if ((int)threadIdx.x < F) {
        const int f = threadIdx.x;
        sh_c[0][f] = norm ? P.utt_mean[(int64_t)b * F + f] : 0.f;
        sh_c[1][f] = norm ? 1.0f / P.utt_std[(int64_t)b * F + f] : 1.f;      // 0 variance: 1/0 = inf, 0 * inf = NaN like x/0
        sh_c[2][f] = has_cm ? P.cmvn_mean[f] : 0.f;
        sh_c[3][f] = has_ci ? P.cmvn_istd[f] : 1.f;
        bool m = false;
        for (int j = 0; j < P.n_fmask; ++j) {
            const int32_t* r = P.fmask + ((int64_t)b * P.n_fmask + j) * 2;
            m |= (f >= r[0]) & (f < r[1]);
        }
        sh_m[f] = m;
    }
    __syncthreads();
    if (rl >= lanes) return;
    float mean[VEC], rstd[VEC], cm[VEC], ci[VEC];
    bool cmask[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        const int f = c + e < F ? c + e : F - 1;
        mean[e] = sh_c[0][f];
        rstd[e] = sh_c[1][f];
        cm[e] = sh_c[2][f];
        ci[e] = sh_c[3][f];
        cmask[e] = sh_m[f] != 0;
    }
    // Lean row loop (the kernel was issue-bound at 118 instructions per row chunk: 70 % issue-slot utilisation at 34 %
    // of the DRAM peak): 64-bit bases once, 32-bit row offsets, the first four time masks in registers, identities
    // instead of branches ((x - 0) * 1 is exact), and the padding value precomputed per thread.
    const float* const rsrc = P.raw + P.frame_prefix[b] * F + c;
    float* const odst = P.out + P.out_row[b] * P.pitch + c;
    const int pitch = (int)P.pitch;
    const int32_t* const fmap = P.frame_map ? P.frame_map + P.map_off[b] : nullptr;
    const int32_t* const tm = P.tmask + (int64_t)b * P.n_tmask * 2;
    constexpr int kFinTm = 4;
    int tm_lo[kFinTm], tm_hi[kFinTm];
#pragma unroll
    for (int j = 0; j < kFinTm; ++j) {
        tm_lo[j] = j < P.n_tmask ? tm[2 * j] : 0;
        tm_hi[j] = j < P.n_tmask ? tm[2 * j + 1] : 0;
    }
    float padv[VEC];                                      // padding rows: 0, or (0 - mean) * istd with CMVN on padding
#pragma unroll
    for (int e = 0; e < VEC; ++e) padv[e] = (has_cm && P.cmvn_on_pad) ? (0.f - cm[e]) * ci[e] : 0.f;
    const int r_end = min(nrows, r0 + kFinRows);
#pragma unroll 1
    for (int t = r0 + rl; t < r_end; t += lanes) {
        float v[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[e] = padv[e];
        if (t < nfr) {
            const int ts = fmap ? fmap[t] : t;
            const float* const src = rsrc + ts * F;
            if (VEC == 4) {
                const float4 x = *reinterpret_cast<const float4*>(src);
                v[0] = x.x; v[1 % VEC] = x.y; v[2 % VEC] = x.z; v[3 % VEC] = x.w;
            } else {
                v[0] = src[0];
            }
            bool rmask = false;
#pragma unroll
            for (int j = 0; j < kFinTm; ++j) rmask |= (t >= tm_lo[j]) & (t < tm_hi[j]);
            for (int j = kFinTm; j < P.n_tmask; ++j) rmask |= (t >= tm[2 * j]) & (t < tm[2 * j + 1]);
            float du[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) du[e] = 0.f;
            if (DITHER) {                  // keyed by the SOURCE frame: substituted rows carry their noise along
                const uint4 r = philox4x32_10(make_uint4((unsigned)(c / VEC), (unsigned)ts, (unsigned)b, 0u),
                                              make_uint2((unsigned)P.dither_seed, (unsigned)(P.dither_seed >> 32)));
                const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int e = 0; e < VEC; ++e) du[e] = ((float)(w[e] >> 8) * (1.0f / 16777216.0f) - 0.5f) * P.dither_a;
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                float y = (v[e] - mean[e]) * rstd[e];                  // mean = 0, rstd = 1 without normalisation: exact
                if (DITHER) y += du[e];
                if (rmask || cmask[e]) y = 0.f;
                v[e] = (y - cm[e]) * ci[e];                            // cm = 0, ci = 1 without CMVN: exact
            }
        }
        float* const dst = odst + t * pitch;
        if (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
        else dst[0] = v[0];
    }
}

// ------------------------------------------------------------------------------------------
// In-place completion of the padded tensor (gen-2 path without spec_sub): the fbank kernel has written the RAW log-mel
// rows straight to their final place (L2 evict_last) together with the per-tile column statistics; this kernel
//   1. merges the utterance's tile statistics itself (feature_processor.py:5-8: mean and population std per bin, fp64,
//      fixed order: row group y of every tile in tile order, then y = 0, 1, 2; two passes, robust for constant features)
//      -- no separate statistics kernel, no mean / std round trip through global memory;
//   2. rewrites the rows in place: (x - mean) * (1 / std) -> [feature dither] -> [time / frequency masks -> 0] ->
//      [(y - cmvn_mean) * cmvn_istd], and fills the padding rows [frames, nrows) with 0 or (0 - mean) * istd
//      (dataset.py:195-218, cmvn.py:43-46).
// No raw scratch: the step's DRAM traffic is the PCM in and the padded tensor out (the re-read hits L2).
// grid = (utterance, part): every part redoes the (small) merge of its utterance and handles a contiguous share of the
// rows; block = 512 threads = 25 row lanes x 20 float4 columns, four rows in flight per thread.
struct Finalize2Params {
    float* out;
    int64_t pitch;
    const int64_t* out_row;      // [B]
    const int64_t* row_prefix;   // [B+1] output rows (frames + padding)
    const int32_t* n_frames;
    const int32_t* tile_prefix;  // [B+1]
    const float* tile_stats;     // [tiles][3][2][F], null without normalisation
    const int32_t* tmask;
    const int32_t* fmask;
    int n_tmask, n_fmask;
    const float* cmvn_mean;
    const float* cmvn_istd;
    int cmvn_on_pad;
    int parts;
    float dither_a;
    unsigned long long dither_seed;
    // spec_sub (feature_processor.py:44-64) as a composed frame map y[t] = x[map[t]], map[t] <= t (every substitution copies
    // from EARLIER frames); null = none.  Handled in place by walking the utterance from its last rows to its first:
    // whatever a row reads lies at or below it and is still raw.  One block per utterance (parts == 1).
    const int32_t* frame_map;
    const int64_t* map_off;      // [B]
    // compute_cmvn_stats: the fbank kernel in front of this one has added the batch's fixed-point sums into stat_acc; block
    // (0, 0) converts them into the caller's fp64 accumulator (null = nothing to do)
    const unsigned long long* stat_acc;      // [2F]
    double* d_stats;                         // [2F+1]
    double stat_count;
};
constexpr int kFin2Threads = 512;
template <bool DITHER, bool SUB = false>
__global__ void __launch_bounds__(kFin2Threads, 2) oe_finalize2_kernel(const Finalize2Params P) {
    constexpr int F = 80;
    __shared__ double shD[3][F];
    __shared__ float shC[2][F];
    grid_dep_wait();
    const int b = blockIdx.x, tid = threadIdx.x;
    if (P.d_stats != nullptr && b == 0 && blockIdx.y == 0) {
        if (tid < 2 * F) P.d_stats[tid] += (double)(long long)__ldcg(P.stat_acc + tid) * (tid < F ? 1.0 / k2::kFxSum : 1.0 / k2::kFxSq);
        if (tid == 2 * F) P.d_stats[2 * F] += P.stat_count;
    }
    const int nfr = P.n_frames[b];
    const int nrows = (int)(P.row_prefix[b + 1] - P.row_prefix[b]);
    const int per = (nrows + P.parts - 1) / P.parts;
    const int r_lo = blockIdx.y * per, r_hi = min(nrows, r_lo + per);
    if (r_lo >= r_hi) return;
    const bool norm = P.tile_stats != nullptr && r_lo < nfr;           // a part that only holds padding needs no statistics
    if (norm) {
        const int f = tid % F, y = tid / F;                            // y < 3: row group y of every tile
        const int tile0 = P.tile_prefix[b], ntiles = P.tile_prefix[b + 1] - tile0;
        const float* const base = P.tile_stats + ((int64_t)tile0 * 3 + (y < 3 ? y : 0)) * 2 * F + f;
        const int nt = y < 3 ? ntiles : 0;
        double acc = 0.0;
#pragma unroll 8
        for (int tl = 0; tl < nt; ++tl) acc += (double)__ldcg(base + tl * 6 * F);
        if (y < 3) shD[y][f] = acc;
        __syncthreads();
        const double S = shD[0][f] + shD[1][f] + shD[2][f];
        const double mean = S / (double)nfr;
        __syncthreads();
        acc = 0.0;
#pragma unroll 8
        for (int tl = 0; tl < nt; ++tl) {
            const int rows = stats_rows(min(kTileFrames, nfr - tl * kTileFrames), y);
            const float sp = __ldcg(base + tl * 6 * F), m2p = __ldcg(base + tl * 6 * F + F);
            const double d = (double)(sp * kInvRows[rows]) - mean;
            acc += rows > 0 ? (double)m2p + (double)rows * d * d : 0.0;
        }
        if (y < 3) shD[y][f] = acc;
        __syncthreads();
        if (y == 0) {
            const double m2 = shD[0][f] + shD[1][f] + shD[2][f];
            shC[0][f] = (float)mean;
            shC[1][f] = 1.0f / (float)sqrt(m2 / (double)nfr);          // 0 variance: 1/0 = inf, 0 * inf = NaN like x/0
        }
        __syncthreads();
    }
    const int q = tid % 20, rl = tid / 20;                             // 25 row lanes x 20 float4 columns
    constexpr int kLanes = 25;
    const bool lane_ok = rl < kLanes;
    if (!SUB && !lane_ok) return;
    const int c = 4 * q;
    const bool has_cm = P.cmvn_mean != nullptr, has_ci = P.cmvn_istd != nullptr;
    float mean4[4], rstd4[4], cm[4], ci[4], padv[4];
    bool cmask[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        mean4[e] = norm ? shC[0][c + e] : 0.f;
        rstd4[e] = norm ? shC[1][c + e] : 1.f;
        cm[e] = has_cm ? __ldg(P.cmvn_mean + c + e) : 0.f;
        ci[e] = has_ci ? __ldg(P.cmvn_istd + c + e) : 1.f;
        padv[e] = (has_cm && P.cmvn_on_pad) ? (0.f - cm[e]) * ci[e] : 0.f;
        bool m = false;
        for (int j = 0; j < P.n_fmask; ++j) {
            const int32_t* r = P.fmask + ((int64_t)b * P.n_fmask + j) * 2;
            m |= (c + e >= __ldg(r)) & (c + e < __ldg(r + 1));
        }
        cmask[e] = m;
    }
    const int32_t* const tm = P.tmask + (int64_t)b * P.n_tmask * 2;
    constexpr int kTm = 4;                                             // time masks held in registers
    int tm_lo[kTm], tm_hi[kTm];
#pragma unroll
    for (int j = 0; j < kTm; ++j) {
        tm_lo[j] = j < P.n_tmask ? __ldg(tm + 2 * j) : 0;
        tm_hi[j] = j < P.n_tmask ? __ldg(tm + 2 * j + 1) : 0;
    }
    const int pitch = (int)P.pitch;
    float* const odst = P.out + P.out_row[b] * P.pitch + c;
    if constexpr (!SUB) {
        const int real_hi = min(r_hi, nfr);
        constexpr int kU = 4;
#pragma unroll 1
        for (int t0 = r_lo + rl; t0 < real_hi; t0 += kLanes * kU) {
            float4 x[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (t0 + kLanes * u < real_hi) x[u] = __ldcg(reinterpret_cast<const float4*>(odst + (int64_t)(t0 + kLanes * u) * pitch));
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int t = t0 + kLanes * u;
                if (t < real_hi) {
                    bool rmask = false;
#pragma unroll
                    for (int j = 0; j < kTm; ++j) rmask |= (t >= tm_lo[j]) & (t < tm_hi[j]);
                    for (int j = kTm; j < P.n_tmask; ++j) rmask |= (t >= __ldg(tm + 2 * j)) & (t < __ldg(tm + 2 * j + 1));
                    float v[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
                    float du[4] = {0.f, 0.f, 0.f, 0.f};
                    if (DITHER) {
                        const uint4 rn = philox4x32_10(make_uint4((unsigned)q, (unsigned)t, (unsigned)b, 0u),
                                                       make_uint2((unsigned)P.dither_seed, (unsigned)(P.dither_seed >> 32)));
                        const unsigned w[4] = {rn.x, rn.y, rn.z, rn.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) du[e] = ((float)(w[e] >> 8) * (1.0f / 16777216.0f) - 0.5f) * P.dither_a;
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float yv = (v[e] - mean4[e]) * rstd4[e];           // mean = 0, rstd = 1 without normalisation: exact
                        if (DITHER) yv += du[e];
                        if (rmask || cmask[e]) yv = 0.f;
                        v[e] = (yv - cm[e]) * ci[e];                       // cm = 0, ci = 1 without CMVN: exact
                    }
                    *reinterpret_cast<float4*>(odst + (int64_t)t * pitch) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
        const float4 pv = make_float4(padv[0], padv[1], padv[2], padv[3]);
#pragma unroll 4
        for (int t = max(r_lo, nfr) + rl; t < r_hi; t += kLanes) *reinterpret_cast<float4*>(odst + (int64_t)t * pitch) = pv;
    } else {
        const int real_hi = min(r_hi, nfr);
        constexpr int kU = 4;
        constexpr int kChunk = kLanes * kU;
        const int32_t* const fmap = SUB ? P.frame_map + P.map_off[b] : nullptr;
        // without a frame map: chunks in ascending order; with one: from the LAST chunk down, loads of a chunk separated from
        // its stores by a block barrier (a row of the chunk may be the source of a later row of the same chunk)
        const int n_chunks = real_hi > r_lo ? (real_hi - r_lo + kChunk - 1) / kChunk : 0;
#pragma unroll 1
        for (int ch = 0; ch < n_chunks; ++ch) {
            const int t0 = r_lo + (SUB ? n_chunks - 1 - ch : ch) * kChunk + rl;
            float4 x[kU];
            int ts[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int t = t0 + kLanes * u;
                ts[u] = t;
                if (lane_ok && t < real_hi) {
                    if (SUB) ts[u] = __ldg(fmap + t);
                    x[u] = __ldcg(reinterpret_cast<const float4*>(odst + (int64_t)ts[u] * pitch));
                }
            }
            if (SUB) __syncthreads();
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int t = t0 + kLanes * u;
                if (lane_ok && t < real_hi) {
                    bool rmask = false;
#pragma unroll
                    for (int j = 0; j < kTm; ++j) rmask |= (t >= tm_lo[j]) & (t < tm_hi[j]);
                    for (int j = kTm; j < P.n_tmask; ++j) rmask |= (t >= __ldg(tm + 2 * j)) & (t < __ldg(tm + 2 * j + 1));
                    float v[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
                    float du[4] = {0.f, 0.f, 0.f, 0.f};
                    if (DITHER) {              // keyed by the SOURCE frame: substituted rows carry their noise along
                        const uint4 rn = philox4x32_10(make_uint4((unsigned)q, (unsigned)ts[u], (unsigned)b, 0u),
                                                       make_uint2((unsigned)P.dither_seed, (unsigned)(P.dither_seed >> 32)));
                        const unsigned w[4] = {rn.x, rn.y, rn.z, rn.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) du[e] = ((float)(w[e] >> 8) * (1.0f / 16777216.0f) - 0.5f) * P.dither_a;
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float yv = (v[e] - mean4[e]) * rstd4[e];           // mean = 0, rstd = 1 without normalisation: exact
                        if (DITHER) yv += du[e];
                        if (rmask || cmask[e]) yv = 0.f;
                        v[e] = (yv - cm[e]) * ci[e];                       // cm = 0, ci = 1 without CMVN: exact
                    }
                    *reinterpret_cast<float4*>(odst + (int64_t)t * pitch) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
        if (lane_ok) {
            const float4 pv = make_float4(padv[0], padv[1], padv[2], padv[3]);
#pragma unroll 4
            for (int t = max(r_lo, nfr) + rl; t < r_hi; t += kLanes) *reinterpret_cast<float4*>(odst + (int64_t)t * pitch) = pv;
        }
    }
}

// Padding rows of the single-pass layout: the fbank kernel writes whole 32-frame tiles (real frames plus the padding
// rows that share the last tile); rows [32 ceil(frames / 32), out_nrows) of every utterance are filled here with
// 0 or (0 - mean) * istd (dataset.py:214-218 pad_sequence, then GlobalCMVN on the padded tensor).  Reads only the
// launch metadata, so under a programmatic dependent launch it runs in the shadow of the fbank kernel's tail: the
// dependency wait comes LAST (work first, wait before exiting), which keeps the chain transitive -- this grid cannot
// complete before the fbank grid has completed and flushed, so whatever waits on this grid also waits on that one.
struct PadFillParams {
    const int32_t* n_frames;
    const int32_t* n_rows;
    const int64_t* out_row;
    float* out;
    int64_t pitch;
    const float* cmvn_mean;      // null: zeros
    const float* cmvn_istd;
    int F;
};
constexpr int kPadRows = 128;        // rows per block
__global__ void __launch_bounds__(256) oe_pad_fill_kernel(const PadFillParams P) {
    const int b = blockIdx.x;
    const int first = (P.n_frames[b] + kTileFrames - 1) / kTileFrames * kTileFrames + blockIdx.y * kPadRows;
    const int last = min(P.n_rows[b], first + kPadRows);
    if (first >= last) {
        grid_dep_wait();
        return;
    }
    float* const dst = P.out + (P.out_row[b] + first) * P.pitch;
    const int F = P.F;
    if (P.pitch == F && F % 4 == 0 && !(reinterpret_cast<uintptr_t>(P.out) & 15)) {      // dense rows: one flat float4 run
        const int q = F / 4, n4 = (last - first) * q;
        for (int i = threadIdx.x; i < n4; i += 256) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (P.cmvn_mean != nullptr) {
                const int c = (i % q) * 4;
                const float* const m = P.cmvn_mean + c;            // caller buffers: no alignment assumed
                v = make_float4(0.f - m[0], 0.f - m[1], 0.f - m[2], 0.f - m[3]);
                if (P.cmvn_istd != nullptr) {
                    const float* const s = P.cmvn_istd + c;
                    v = make_float4(v.x * s[0], v.y * s[1], v.z * s[2], v.w * s[3]);
                }
            }
            reinterpret_cast<float4*>(dst)[i] = v;
        }
    } else {
        for (int i = threadIdx.x; i < (last - first) * F; i += 256) {
            const int r = i / F, f = i - r * F;
            float v = 0.f;
            if (P.cmvn_mean != nullptr) {
                v = 0.f - P.cmvn_mean[f];
                if (P.cmvn_istd != nullptr) v = v * P.cmvn_istd[f];
            }
            dst[(int64_t)r * P.pitch + f] = v;
        }
    }
    grid_dep_wait();
}

// ------------------------------------------------------------------------------------------
// The first consumer of the (B, T, F) feature tensor, fused with the step in front of it:
//   encoder.py:221-222   xs = global_cmvn(xs)                    (x - mean) * istd on the padded batch, cmvn.py:43-46
//   subsampling.py:76-78 Conv2d(1, odim, 3, 2) + ReLU on xs.unsqueeze(1)   (first layer of Conv2dSubsampling4, :110-111)
// y[b][c][t][f] = relu(bias[c] + sum_{i,j<3} w[c][i][j] * cmvn(x)[b][2t + i][2f + j]),  t < (T-3)/2+1, f < (F-3)/2+1.
// GlobalCMVN alone is a full read + write of the batch; here it costs two FMAs per staged input value.  The kernel is
// bound by the stores of its output, odim/4 x (F'/F) ~ 31x the input (5 GB for a 256 x 1000 x 80 batch at odim 256):
// thread = one output position (t, f) of the tile with its 9 inputs in registers, looping over the channels; the weights
// are read as shared-memory broadcasts; a warp's stores of one channel are 32 consecutive floats of the (t, f) plane.
struct ConvSubParams {
    const float* x;              // (B, T, F) fp32, row pitch `pitch`
    int64_t pitch;
    int B, T, F, odim, T1, F1;
    const float* w;              // [odim][9]
    const float* bias;           // [odim]
    const float* cmvn_mean;      // [F] or null
    const float* cmvn_istd;      // [F] or null
    float* y;                    // (B, odim, T1, F1)
};
constexpr int kConvTT = 8;       // output rows per block
constexpr int kConvThreads = 320;
__global__ void __launch_bounds__(kConvThreads) oe_conv_sub1_kernel(const ConvSubParams P) {
    extern __shared__ __align__(16) float csm[];
    float* const sw = csm;                                   // [odim][12]: 9 weights, bias, 2 pad -> three LDS.128 per channel
    float* const sx = csm + (size_t)P.odim * 12;             // [2 kConvTT + 1][F] staged (CMVN'd) input rows
    const int b = blockIdx.y, t0 = blockIdx.x * kConvTT;
    const int rows = min(2 * kConvTT + 1, P.T - 2 * t0);
    for (int i = threadIdx.x; i < P.odim; i += kConvThreads) {
#pragma unroll
        for (int k = 0; k < 9; ++k) sw[i * 12 + k] = P.w[i * 9 + k];
        sw[i * 12 + 9] = P.bias ? P.bias[i] : 0.f;
        sw[i * 12 + 10] = sw[i * 12 + 11] = 0.f;
    }
    const float* const src = P.x + ((int64_t)b * P.T + 2 * t0) * P.pitch;
    for (int i = threadIdx.x; i < rows * P.F; i += kConvThreads) {
        const int r = i / P.F, f = i - r * P.F;
        float v = src[(int64_t)r * P.pitch + f];
        if (P.cmvn_mean) {
            v = v - P.cmvn_mean[f];
            if (P.cmvn_istd) v = v * P.cmvn_istd[f];
        }
        sx[i] = v;
    }
    __syncthreads();
    const int npos = min(kConvTT, P.T1 - t0) * P.F1;
    float* const ybase = P.y + (int64_t)b * P.odim * P.T1 * P.F1 + (int64_t)t0 * P.F1;
    const int64_t plane = (int64_t)P.T1 * P.F1;
    for (int pos = threadIdx.x; pos < npos; pos += kConvThreads) {
        const int tl = pos / P.F1, f = pos - tl * P.F1;
        float in[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) in[3 * i + j] = sx[(2 * tl + i) * P.F + 2 * f + j];
        float* out = ybase + pos;
#pragma unroll 4
        for (int c = 0; c < P.odim; ++c) {
            const float4 w0 = *reinterpret_cast<const float4*>(sw + c * 12);
            const float4 w1 = *reinterpret_cast<const float4*>(sw + c * 12 + 4);
            const float4 w2 = *reinterpret_cast<const float4*>(sw + c * 12 + 8);
            float a = w2.y;                                      // bias
            a = fmaf(w0.x, in[0], a); a = fmaf(w0.y, in[1], a); a = fmaf(w0.z, in[2], a);
            a = fmaf(w0.w, in[3], a); a = fmaf(w1.x, in[4], a); a = fmaf(w1.y, in[5], a);
            a = fmaf(w1.z, in[6], a); a = fmaf(w1.w, in[7], a); a = fmaf(w2.x, in[8], a);
            __stcs(out, fmaxf(a, 0.f));                          // written once, read by the next layer: streaming store
            out += plane;
        }
    }
}

// openeat/modules/cmvn.py:43-46
__global__ void __launch_bounds__(256) oe_cmvn_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                      int64_t n, int dim, const float* __restrict__ mean,
                                                      const float* __restrict__ istd) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int f = (int)(i % dim);
        float v = x[i] - mean[f];
        if (istd) v = v * istd[f];
        y[i] = v;
    }
}

// ------------------------------------------------------------------------------------------
struct RsTable {
    int orig, neu, taps, width, coef_off;
};
constexpr int kRsDirectCoefs = 1 << 16;     // registered tables above this many coefficients are evaluated on the fly

struct ResampleParams {
    const void* in;
    float* out;
    const int64_t* in_off;
    const int32_t* in_len;
    const int32_t* table_id;
    const int64_t* out_off;
    const int32_t* out_len;
    const RsTable* tables;
    const float* coefs;
    const int32_t* orig;         // [B] rates of the utterances with table_id == OE_RS_DIRECT (others ignored), or null
    const int32_t* neu;
};

// torchaudio functional.py:1401-1432: y[m*new + p] = sum_q k[p][q] * xpad[m*orig + q], xpad = x shifted by width.
// Generic table-driven kernel (any ratio); `skip_fast` leaves utterances to the specialised kernel below.
template <bool kF32>
__global__ void __launch_bounds__(256) oe_resample_kernel(const ResampleParams P, int fast_a, int fast_b) {
    const int b = blockIdx.y;
    const int n_out = P.out_len[b];
    const int n_in = P.in_len[b];
    const int tid_ = P.table_id[b];
    if (tid_ >= 0 && (tid_ == fast_a || tid_ == fast_b)) return;
    float* const out = P.out + P.out_off[b];
    const int64_t ioff = P.in_off[b];
    const int stride = gridDim.x * blockDim.x;
    auto sample = [&](int xi) -> float {
        return kF32 ? reinterpret_cast<const float*>(P.in)[ioff + xi] : (float)reinterpret_cast<const int16_t*>(P.in)[ioff + xi];
    };
    if (tid_ == OE_RS_DIRECT) {
        // Ratios whose polyphase table would be huge (speeds drawn from a continuous range, 44.1 kHz sources): every
        // coefficient is evaluated where it is used, with torchaudio's formula (functional.py:1343-1398: hann-windowed
        // sinc, lowpass_filter_width 6, rolloff 0.99); only the ~13 orig/min(orig, new) taps inside the window are
        // visited (the table holds exact zeros everywhere else).  The filter argument is formed from the exact integer
        // j new - phase orig: torch's own fp32 grid (p/new + idx/orig, a difference of two numbers near 1) is noisy for
        // long periods -- its resampled 953:1000 waveform sits 0.8 (int16 scale) from the float64 evaluation, this
        // kernel 1e-3 (tests/test_gpu_mirrors.py::test_long_ratio_and_kaiser_resamplers).
        const int orig = P.orig[b], neu = P.neu[b];
        const float base = (float)(orig < neu ? orig : neu) * 0.99f;
        const float halfw = 6.0f * (float)orig / base;                  // taps with |t| < 6 around the output instant
        const float scale = base / (float)orig;
        const float tscale = base / ((float)orig * (float)neu);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
            const int m = i / neu, ph = i - m * neu;
            const float c = (float)ph * (float)orig / (float)neu;       // output instant, in input samples past m * orig
            const int j_lo = (int)ceilf(c - halfw) - 1, j_hi = (int)floorf(c + halfw) + 1;
            float acc = 0.f;
            for (int j = j_lo; j <= j_hi; ++j) {
                const int xi = m * orig + j;
                if (xi < 0 || xi >= n_in) continue;
                float t = (float)((long long)j * neu - (long long)ph * orig) * tscale;
                t = fminf(6.0f, fmaxf(-6.0f, t));
                const float w = cosf(t * 3.14159265358979323846f / 6.0f / 2.0f);
                const float a = t * 3.14159265358979323846f;
                const float k = (a == 0.f ? 1.0f : sinf(a) / a) * (w * w) * scale;
                acc = fmaf(k, sample(xi), acc);
            }
            out[i] = acc;
        }
        return;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
        float acc = 0.f;
        if (tid_ < 0) {
            acc = sample(i);
        } else {
            const RsTable t = P.tables[tid_];
            const int m = i / t.neu, ph = i - m * t.neu;
            const float* const k = P.coefs + (size_t)t.coef_off + (size_t)ph * t.taps;
            const int x0 = m * t.orig - t.width;
            for (int q = 0; q < t.taps; ++q) {
                const int xi = x0 + q;
                if (xi >= 0 && xi < n_in) acc = fmaf(k[q], sample(xi), acc);
            }
        }
        out[i] = acc;
    }
}

// ---- specialised polyphase kernel for the speed-perturb ratios (9:10 and 11:10, width 7) ----
// hann-windowed sinc taps vanish where the window argument is clamped (functional.py:1370-1381):
// tap (p, q) is non-zero only if |(-p/new + (q - width)/orig) * 0.99 min(orig, new)| < 6.
struct RsFastParams {
    ResampleParams r;
    int table_id;
    float coef[10 * 25];            // [new][taps] in the kernel-parameter constant bank
};

constexpr int kRsM = 256;           // polyphase blocks (of `new` outputs) per CTA, one per thread

template <bool kF32, int ORIG, int NEW, int WIDTH>
__global__ void __launch_bounds__(kRsM) oe_resample_fast_kernel(const RsFastParams P) {
    constexpr int TAPS = 2 * WIDTH + ORIG;
    constexpr int XIN = kRsM * ORIG + 2 * WIDTH;
    __shared__ float sx[XIN + 8];
    __shared__ float sy[kRsM * NEW];
    const int b = blockIdx.y;
    if (P.r.table_id[b] != P.table_id) return;
    const int n_out = P.r.out_len[b];
    const int m0 = blockIdx.x * kRsM;
    if (m0 * NEW >= n_out) return;
    const int n_in = P.r.in_len[b];
    const int64_t ioff = P.r.in_off[b];
    const int tid = threadIdx.x;
    const int x_base = m0 * ORIG - WIDTH;                       // first input sample this CTA needs
    for (int i = tid; i < XIN; i += kRsM) {
        const int xi = x_base + i;
        float v = 0.f;
        if (xi >= 0 && xi < n_in)
            v = kF32 ? __ldg(reinterpret_cast<const float*>(P.r.in) + ioff + xi)
                     : (float)__ldg(reinterpret_cast<const int16_t*>(P.r.in) + ioff + xi);
        sx[i] = v;
    }
    __syncthreads();
    {
        float x[TAPS];
#pragma unroll
        for (int q = 0; q < TAPS; ++q) x[q] = sx[tid * ORIG + q];
        static_for<0, NEW>([&](auto pp) {
            constexpr int p = decltype(pp)::value;
            float acc = 0.f;
            static_for<0, TAPS>([&](auto qq) {
                constexpr int q = decltype(qq)::value;
                if constexpr (rs_tap_nonzero(ORIG, NEW, WIDTH, p, q)) acc = fmaf(P.coef[p * TAPS + q], x[q], acc);
            });
            sy[tid * NEW + p] = acc;
        });
    }
    __syncthreads();
    float* const out = P.r.out + P.r.out_off[b] + (int64_t)m0 * NEW;
    const int n_here = min(kRsM * NEW, n_out - m0 * NEW);
    for (int i = tid; i < n_here; i += kRsM) out[i] = sy[i];
}

}  // namespace oe

// ==========================================================================================
// C ABI
// ==========================================================================================
struct oe_frontend {
    oe_config cfg;
    int device;
    int sm_count;
    oe::DevTables* d_tab;
    std::vector<float> window, mel;
    std::vector<oe::RsTable> rs;
    std::vector<char> rs_builtin;      // table i is the library's own hann sinc (found again by ratio), not a caller's kernel
    std::vector<float> rs_coefs;
    int rs_fast_9_10, rs_fast_11_10;   // table ids served by oe_resample_fast_kernel, or -1
    oe::RsTable* d_rs;
    float* d_rs_coefs;
    size_t d_rs_cap, d_rs_coefs_cap;   // device capacities (tables / floats); grown in oe_add_resampler
    size_t fbank_smem;
    bool std_mel;                  // the mel matrix has the baked mel80 structure -> kernels with a compile-time mel structure
    bool mel_baked;                // ... and exactly torchaudio's weights (oe_mel80.h) -> gen-2 kernel, weights as FFMA immediates
    bool force_v1;                 // OE_FBANK_V1=1: first-generation kernel (A/B timing only)
    int fin2_parts;                // OE_FIN2_PARTS=n: blocks per utterance of oe_finalize2_kernel (tuning only; 0 = automatic)
    bool no_inline_meta;           // OE_NO_INLINE_META=1: small batches send their metadata through oe_fetch_kernel too (A/B timing only)
    bool no_coalesce;              // OE_NO_COALESCE=1: consecutive windows of one recording stay separate utterances (A/B timing only)
    bool no_inplace_sub;           // OE_NO_INPLACE_SUB=1: spec_sub batches keep the ragged scratch + statistics kernel + out-of-place finalize (A/B timing only)
    bool no_inplace;               // OE_NO_INPLACE=1: raw scratch + statistics kernel + out-of-place finalize behind the gen-2 kernel (A/B timing only)
    // pinned staging ring for the per-call metadata block (a pageable source would make cudaMemcpyAsync wait for the
    // stream); a slot is reused once the copy that read it has completed
    static constexpr int kMetaSlots = 4;
    unsigned char* h_meta[kMetaSlots];
    unsigned char* h_meta_dev[kMetaSlots];     // the same slots as the device sees them (mapped pinned memory)
    size_t h_meta_cap[kMetaSlots];
    cudaEvent_t h_meta_ev[kMetaSlots];
    int h_meta_next;
    long long launches;            // kernels launched through this handle
    bool timing;                   // oe_frontend_set_kernel_timing
    bool timed;                    // the events below bracket a kernel of the most recent call
    cudaEvent_t ev_begin, ev_end;
    cudaEvent_t ev_step[2];        // timing on: the whole launch sequence of the most recent call
    float mel_w_std[512];
};

namespace {

thread_local char g_err[512] = "";

// Launch with programmatic stream serialization (PDL): the grid may be scheduled while its predecessor in the stream
// drains; every kernel launched this way executes grid_dep_wait() before it touches anything a predecessor wrote
// (and every kernel of the chain does so, which keeps the ordering transitive).  OE_NO_PDL=1 disables it.
template <class Params>
cudaError_t launch_dep(void (*kern)(const Params), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const Params& prm) {
    // one flag per instantiation, shared by every host thread that launches through this library: atomic, relaxed (a
    // thread that still sees `true` after another one switched it off merely repeats the fallback below)
    static std::atomic<bool> pdl_flag([] { const char* e = getenv("OE_NO_PDL"); return !(e && e[0] == '1'); }());
    const bool pdl = pdl_flag.load(std::memory_order_relaxed);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, prm);
    if (e != cudaSuccess && pdl && (e == cudaErrorNotSupported || e == cudaErrorInvalidValue)) {
        (void)cudaGetLastError();             // a driver / device without programmatic launches: plain stream order from now on
        pdl_flag.store(false, std::memory_order_relaxed);
        cfg.numAttrs = 0;
        e = cudaLaunchKernelEx(&cfg, kern, prm);
    }
    return e;
}


int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define OE_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(OE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// kaldi.py:98-100 in double, rounded once
void default_window(int n, std::vector<float>& w) {
    w.assign(n, 0.f);
    for (int j = 0; j < n; ++j) {
        double h = 0.5 - 0.5 * std::cos(2.0 * oe::kPi * j / (n - 1));
        if (j == 0 || j == n - 1) h = 0.0;
        w[j] = (float)std::pow(h, 0.85);
    }
}

// kaldi.py:436-511 (vtln_warp 1.0): scalar limits in double, per-bin arithmetic in fp32 like torch
void default_mel(const oe_config& c, std::vector<float>& m) {
    const int nb = c.num_mel_bins, nf = c.fft_size / 2;
    if (nb == oe::mel80::kBins && c.sample_rate == 16000 && c.fft_size == 512 && c.low_freq == 20.0f && c.high_freq <= 0.0f &&
        c.high_freq + 8000.0f == 8000.0f) {
        // the standard configuration gets torchaudio's own matrix, bit for bit (oe_mel80.h): same values as
        // kaldi.get_mel_banks, and the handle qualifies for the kernel that carries the weights as immediates
        m.assign((size_t)nb * nf, 0.f);
        for (int b = 0; b < nb; ++b)
            for (int i = 0; i < oe::mel80::kLen[b]; ++i)
                m[(size_t)b * nf + oe::mel80::kStart[b] + i] = 4.0f * oe::mel80::weight(oe::mel80::kOff[b] + i);
        return;
    }
    const double nyq = 0.5 * c.sample_rate;
    double high = c.high_freq;
    if (high <= 0.0) high += nyq;
    const double mel_lo = 1127.0 * std::log(1.0 + (double)c.low_freq / 700.0);
    const double mel_hi = 1127.0 * std::log(1.0 + high / 700.0);
    const double delta = (mel_hi - mel_lo) / (nb + 1);
    const float bw = (float)((double)c.sample_rate / c.fft_size);
    m.assign((size_t)nb * nf, 0.f);
    for (int b = 0; b < nb; ++b) {
        const float left = (float)mel_lo + (float)b * (float)delta;
        const float center = (float)mel_lo + ((float)b + 1.0f) * (float)delta;
        const float right = (float)mel_lo + ((float)b + 2.0f) * (float)delta;
        for (int k = 0; k < nf; ++k) {
            const float mel = 1127.0f * std::log(1.0f + (bw * (float)k) / 700.0f);
            const float up = (mel - left) / (center - left);
            const float down = (right - mel) / (right - center);
            m[(size_t)b * nf + k] = std::max(0.0f, std::min(up, down));
        }
    }
}

struct Meta {               // device-side metadata block layout (byte offsets into the workspace)
    size_t wav_off, out_row, frame_prefix, row_prefix, map_off;          // int64 arrays
    size_t wav_len, n_frames, n_rows, tile_prefix, rs_mode, tmask, fmask, fmap, tiles;   // int32 arrays
    size_t meta_bytes;
    size_t raw, tile_stats, utt_mean, utt_std, stat_partial, sched, stat_acc, total;
    int max_rows;
    int64_t total_frames, total_rows, total_map;
    int64_t pad_rows;           // output rows behind the last 32-frame tile of each utterance
    int total_tiles;
    bool two_phase, need_stats, feats;
    bool k2;                    // the gen-2 kernel (standard 80-bin mel matrix, waveform input) runs this batch
    bool inplace_ok;            // ... writes the raw rows to their final place, oe_finalize2_kernel completes them there
};

int plan(const oe_frontend* fe, const oe_batch* bt, Meta& M, std::vector<int32_t>* frames_out) {
    if (!fe || !bt) return fail(OE_ERR_INVALID, "null frontend or batch");
    if (bt->batch < 0) return fail(OE_ERR_INVALID, "negative batch");
    if (bt->batch > 0 && (!bt->wav_offsets || !bt->wav_lens || !bt->out_rows))
        return fail(OE_ERR_INVALID, "wav_offsets, wav_lens and out_rows are required");
    if (bt->wav_dtype != OE_WAV_I16 && bt->wav_dtype != OE_WAV_F32 && bt->wav_dtype != OE_FEATS_F32)
        return fail(OE_ERR_INVALID, "bad wav_dtype");
    const bool feats = bt->wav_dtype == OE_FEATS_F32;
    if (bt->n_tmask < 0 || bt->n_fmask < 0 || (bt->n_tmask > 0 && !bt->tmask) || (bt->n_fmask > 0 && !bt->fmask))
        return fail(OE_ERR_INVALID, "mask counts / pointers inconsistent");
    if (bt->frame_map && !bt->frame_map_offsets) return fail(OE_ERR_INVALID, "frame_map without offsets");
    const int B = bt->batch, F = fe->cfg.num_mel_bins;
    const int64_t pitch = bt->out_pitch ? bt->out_pitch : F;
    if (pitch < F) return fail(OE_ERR_INVALID, "out_pitch smaller than num_mel_bins");
    if (bt->wav_dither != 0.f) {                         // validated here, before anything is launched
        if (feats) return fail(OE_ERR_INVALID, "wav_dither needs waveform input");
        if (!fe->mel_baked || fe->force_v1) return fail(OE_ERR_UNSUPPORTED, "wav_dither is built for the standard 80-bin kernel (torchaudio's mel matrix) only");
        for (int b = 0; bt->resample_ids && b < B; ++b)
            if (bt->resample_ids[b] >= 0)
                return fail(OE_ERR_UNSUPPORTED, "wav_dither cannot be combined with a fused speed perturb: resample first (oe_resample)");
    }
    M.two_phase = feats || bt->norm_mode != OE_NORM_NONE || bt->frame_map != nullptr || bt->feature_dither != 0.f;
    M.feats = feats;
    M.need_stats = bt->norm_mode != OE_NORM_NONE || (feats && bt->d_stats != nullptr);
    M.k2 = !feats && fe->mel_baked && !fe->force_v1;
    M.inplace_ok = M.k2 && M.two_phase && !fe->no_inplace && F % 4 == 0 && pitch % 4 == 0;
    bool map_from_past = true;
    M.total_frames = M.total_rows = M.total_map = 0;
    M.pad_rows = 0;
    M.max_rows = 0;
    int64_t tiles = 0;
    if (frames_out) frames_out->resize(B);
    for (int b = 0; b < B; ++b) {
        if (bt->wav_lens[b] < 0) return fail(OE_ERR_INVALID, "negative wav_len at %d", b);
        if (bt->wav_offsets[b] < 0 || (!feats && (bt->wav_offsets[b] & 7)))
            return fail(OE_ERR_INVALID, "wav_offsets[%d] must be a non-negative multiple of 8 samples", b);
        int64_t eff_len = bt->wav_lens[b];
        if (bt->resample_ids && bt->resample_ids[b] >= 0) {
            const int id = bt->resample_ids[b];
            if (feats || bt->wav_dtype != OE_WAV_I16) return fail(OE_ERR_UNSUPPORTED, "fused speed perturb takes int16 PCM input");
            if (id != fe->rs_fast_9_10 && id != fe->rs_fast_11_10)
                return fail(OE_ERR_UNSUPPORTED, "resample_ids[%d]=%d: only the 9:10 and 11:10 (width 7) tables can be fused; "
                                                "use oe_resample for other ratios", b, id);
            eff_len = oe_resample_out_len(eff_len, fe->rs[id].orig, fe->rs[id].neu);
        }
        const int nfr = feats ? bt->wav_lens[b] : oe_num_frames(fe, eff_len);
        const int nrows = bt->out_nrows ? bt->out_nrows[b] : nfr;
        if (nrows < nfr) return fail(OE_ERR_INVALID, "out_nrows[%d]=%d < frames %d", b, nrows, nfr);
        if (frames_out) (*frames_out)[b] = nfr;
        M.total_frames += nfr;
        M.total_rows += nrows;
        M.max_rows = std::max(M.max_rows, nrows);
        tiles += (nfr + oe::kTileFrames - 1) / oe::kTileFrames;      // single pass: padding rows behind the last tile: oe_pad_fill_kernel
        M.pad_rows += std::max(0, nrows - (nfr + oe::kTileFrames - 1) / oe::kTileFrames * oe::kTileFrames);
        if (bt->frame_map) {
            M.total_map = std::max<int64_t>(M.total_map, bt->frame_map_offsets[b] + nfr);
            // spec_sub only ever copies from EARLIER frames (feature_processor.py:57-63): such a map can be applied in place
            const int32_t* const mp = bt->frame_map + bt->frame_map_offsets[b];
            for (int t = 0; t < nfr && map_from_past; ++t) map_from_past = mp[t] <= t;
        }
    }
    if (tiles > INT32_MAX) return fail(OE_ERR_INVALID, "batch too large");
    if (bt->frame_map && (!map_from_past || fe->no_inplace_sub)) M.inplace_ok = false;      // a general gather keeps the raw scratch
    M.total_tiles = (int)tiles;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 16); return r; };
    M.wav_off = take(8 * (size_t)B);
    M.out_row = take(8 * (size_t)B);
    M.frame_prefix = take(8 * (size_t)(B + 1));
    M.row_prefix = take(8 * (size_t)(B + 1));
    M.map_off = take(8 * (size_t)B);
    M.wav_len = take(4 * (size_t)B);
    M.n_frames = take(4 * (size_t)B);
    M.n_rows = take(4 * (size_t)B);
    M.tile_prefix = take(4 * (size_t)(B + 1));
    M.rs_mode = take(4 * (size_t)B);
    M.tmask = take(8 * (size_t)B * bt->n_tmask);
    M.fmask = take(8 * (size_t)B * bt->n_fmask);
    M.fmap = take(4 * (size_t)M.total_map);
    M.meta_bytes = o;
    M.tiles = take(sizeof(oe::TileDesc) * (size_t)M.total_tiles);
    // raw log-mel scratch: not when the rows are finished in place
    M.raw = take(M.two_phase && !feats && !M.inplace_ok ? 4 * (size_t)M.total_frames * F : 0);
    M.tile_stats = take(M.need_stats ? 4 * (size_t)M.total_tiles * 3 * 2 * F : 0);
    M.utt_mean = take(bt->norm_mode != OE_NORM_NONE && !M.inplace_ok ? 4 * (size_t)B * F : 0);
    M.utt_std = take(bt->norm_mode != OE_NORM_NONE && !M.inplace_ok ? 4 * (size_t)B * F : 0);
    M.stat_partial = take(bt->d_stats && !M.k2 ? 8 * (size_t)(fe->sm_count * 8) * 3 * 2 * F : 0);
    M.sched = take(M.k2 ? 16 : 0);
    M.stat_acc = take(M.k2 && bt->d_stats ? 8 * (size_t)2 * F : 0);
    M.total = o;
    return OE_OK;
}

}  // namespace

extern "C" {

const char* oe_last_error(void) { return g_err; }
int oe_abi_version(void) { return OE_ABI_VERSION; }

int oe_config_default(oe_config* cfg) {
    if (!cfg) return fail(OE_ERR_INVALID, "null cfg");
    cfg->sample_rate = 16000;
    cfg->frame_length = 400;
    cfg->frame_shift = 160;
    cfg->fft_size = 512;
    cfg->num_mel_bins = 80;
    cfg->preemph = 0.97f;
    cfg->low_freq = 20.0f;
    cfg->high_freq = 0.0f;
    cfg->log_floor = 1.1920928955078125e-07f;
    return OE_OK;
}

int32_t oe_num_frames(const oe_frontend* fe, int64_t n) {
    const int win = fe ? fe->cfg.frame_length : oe::kWin, shift = fe ? fe->cfg.frame_shift : oe::kShift;
    if (n < win) return 0;
    return (int32_t)(1 + (n - win) / shift);
}

int oe_frontend_create(const oe_config* cfg, const float* window, const float* mel, int device, oe_frontend** out) {
    if (!cfg || !out) return fail(OE_ERR_INVALID, "null cfg/out");
    *out = nullptr;
    if (cfg->frame_length != oe::kWin || cfg->frame_shift != oe::kShift || cfg->fft_size != oe::kFft)
        return fail(OE_ERR_UNSUPPORTED, "only 25 ms / 10 ms framing at 16 kHz (400/160/512 samples) is built; got %d/%d/%d",
                    cfg->frame_length, cfg->frame_shift, cfg->fft_size);
    if (cfg->num_mel_bins < 4 || cfg->num_mel_bins > oe::kMaxMel)
        return fail(OE_ERR_UNSUPPORTED, "num_mel_bins must be in [4, %d]", oe::kMaxMel);
    if (!(cfg->preemph >= 0.f && cfg->preemph <= 1.f)) return fail(OE_ERR_INVALID, "preemph must be in [0,1]");
    if (device < 0) OE_CUDA(cudaGetDevice(&device));
    OE_CUDA(cudaSetDevice(device));
    oe_frontend* fe = new oe_frontend();
    fe->cfg = *cfg;
    fe->device = device;
    fe->d_tab = nullptr;
    fe->launches = 0;
    fe->timing = fe->timed = false;
    fe->ev_begin = fe->ev_end = nullptr;
    fe->ev_step[0] = fe->ev_step[1] = nullptr;
    fe->d_rs = nullptr;
    fe->d_rs_coefs = nullptr;
    fe->d_rs_cap = fe->d_rs_coefs_cap = 0;
    fe->rs_fast_9_10 = fe->rs_fast_11_10 = -1;
    fe->h_meta_next = 0;
    for (int i = 0; i < oe_frontend::kMetaSlots; ++i) {
        fe->h_meta[i] = nullptr;
        fe->h_meta_dev[i] = nullptr;
        fe->h_meta_cap[i] = 0;
        fe->h_meta_ev[i] = nullptr;
    }
    const int nb = cfg->num_mel_bins, nf = cfg->fft_size / 2;
    if (window) fe->window.assign(window, window + cfg->frame_length); else default_window(cfg->frame_length, fe->window);
    if (mel) fe->mel.assign(mel, mel + (size_t)nb * nf); else default_mel(*cfg, fe->mel);
    for (int b = 0; b < nb; ++b)
        if (fe->mel[(size_t)b * nf] != 0.f) {
            delete fe;
            return fail(OE_ERR_UNSUPPORTED, "mel weight on fft bin 0 is not supported (Kaldi banks never have one)");
        }
    if (fe->window[0] != 0.f) {
        delete fe;
        return fail(OE_ERR_UNSUPPORTED, "window[0] must be 0 (povey): pre-emphasis is folded into the staged waveform");
    }
    std::vector<oe::DevTables> hv(1);
    oe::DevTables& h = hv[0];
    memset(&h, 0, sizeof(h));
    for (int j = 0; j < cfg->frame_length; ++j) h.window[j] = fe->window[j];
    for (int tau = 0; tau < 16; ++tau)
        for (int k1 = 0; k1 < 16; ++k1) {
            const double a = 2.0 * oe::kPi * (double)((tau * k1) % 256) / 256.0;
            h.twA[tau * oe::kRowE + k1] = make_float2((float)std::cos(a), (float)-std::sin(a));
        }
    for (int u = 0; u < 16; ++u)
        for (int k2 = 0; k2 < 16; ++k2) {
            const double a = 2.0 * oe::kPi * (double)(u + 16 * k2) / 512.0;
            h.twU[u * oe::kRowE + k2] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    int nnz = 0;
    for (int b = 0; b < nb; ++b) {
        int first = -1, last = -1;
        for (int k = 0; k < nf; ++k)
            if (fe->mel[(size_t)b * nf + k] != 0.f) {
                if (first < 0) first = k;
                last = k;
            }
        h.mel_start[b] = first < 0 ? 0 : first;
        h.mel_len[b] = first < 0 ? 0 : last - first + 1;
        h.mel_off[b] = nnz;
        if (nnz + h.mel_len[b] > oe::kMaxNnz) {
            delete fe;
            return fail(OE_ERR_UNSUPPORTED, "mel matrix too dense (> %d stored weights)", oe::kMaxNnz);
        }
        for (int i = 0; i < h.mel_len[b]; ++i) h.mel_w[nnz + i] = 0.25f * fe->mel[(size_t)b * nf + h.mel_start[b] + i];
        nnz += h.mel_len[b];
    }
    h.nnz = nnz;
    h.n_mel = nb;
    h.log_floor = cfg->log_floor;
    h.preemph = cfg->preemph;
    {   // contiguous mel-bin groups for the 8 warps, balanced by stored weights (+2 per bin of overhead)
        const double total = nnz + 2.0 * nb;
        int b = 0;
        double acc = 0.0;
        h.group_begin[0] = 0;
        for (int g = 1; g < 8; ++g) {
            while (b < nb && acc + 0.5 * (h.mel_len[b] + 2) <= total * g / 8.0) acc += h.mel_len[b++] + 2;
            h.group_begin[g] = b;
        }
        h.group_begin[8] = nb;
    }
    cudaError_t e = cudaMalloc(&fe->d_tab, sizeof(oe::DevTables));
    if (e == cudaSuccess) e = cudaMemcpy(fe->d_tab, &h, sizeof(h), cudaMemcpyHostToDevice);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    fe->std_mel = nb == oe::mel80::kBins && nnz == oe::mel80::kNnz;
    for (int b = 0; fe->std_mel && b < nb; ++b)
        fe->std_mel = h.mel_start[b] == oe::mel80::kStart[b] && h.mel_len[b] == oe::mel80::kLen[b];
    fe->mel_baked = fe->std_mel;
    for (int i = 0; fe->mel_baked && i < nnz; ++i) {
        const float w = oe::mel80::weight(i);
        fe->mel_baked = memcmp(&w, &h.mel_w[i], 4) == 0;
    }
    memset(fe->mel_w_std, 0, sizeof(fe->mel_w_std));
    if (fe->std_mel) memcpy(fe->mel_w_std, h.mel_w, sizeof(float) * nnz);
    fe->fbank_smem = fe->std_mel ? align_up((size_t)oe::kSmStd, 16) : align_up((size_t)oe::kSmMelW + 4 * (size_t)nnz, 16);
    const int smem_i = (int)fe->fbank_smem;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(oe::oe_fbank_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_i);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(oe::oe_fbank_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_i);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(oe::oe_fbank_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_i);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(oe::oe_fbank_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_i);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(oe::oe_fbank_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_i);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(oe::oe_fbank_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_i);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(oe::k2::oe_fbank2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, oe::k2::Smem<false, false>::End);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(oe::k2::oe_fbank2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, oe::k2::Smem<false, true>::End);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(oe::k2::oe_fbank2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, oe::k2::Smem<true, false>::End);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(oe::k2::oe_fbank2_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, oe::k2::Smem<false, false>::End);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(oe::k2::oe_fbank2_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, oe::k2::Smem<true, false>::End);
    {   // developer switch for A/B timing: OE_FBANK_V1=1 keeps the first-generation kernel on the standard mel layout
        const char* v1 = getenv("OE_FBANK_V1");
        fe->force_v1 = v1 && v1[0] == '1';
        const char* fp_ = getenv("OE_FIN2_PARTS");
        fe->fin2_parts = fp_ ? atoi(fp_) : 0;
        const char* nf = getenv("OE_NO_INPLACE");
        fe->no_inplace = nf && nf[0] == '1';
        const char* ni = getenv("OE_NO_INLINE_META");
        fe->no_inline_meta = ni && ni[0] == '1';
        const char* nz = getenv("OE_NO_COALESCE");
        fe->no_coalesce = nz && nz[0] == '1';
        const char* ns = getenv("OE_NO_INPLACE_SUB");
        fe->no_inplace_sub = ns && ns[0] == '1';
    }
    if (e != cudaSuccess) {
        if (fe->d_tab) cudaFree(fe->d_tab);
        if (fe->d_rs) cudaFree(fe->d_rs);
        if (fe->d_rs_coefs) cudaFree(fe->d_rs_coefs);
        delete fe;
        return fail(OE_ERR_CUDA, "frontend setup failed: %s", cudaGetErrorString(e));
    }
    fe->sm_count = prop.multiProcessorCount;
    *out = fe;
    return OE_OK;
}

int oe_frontend_destroy(oe_frontend* fe) {
    if (!fe) return OE_OK;
    if (fe->ev_begin) cudaEventDestroy(fe->ev_begin);
    if (fe->ev_end) cudaEventDestroy(fe->ev_end);
    for (int i = 0; i < 2; ++i)
        if (fe->ev_step[i]) cudaEventDestroy(fe->ev_step[i]);
    for (int i = 0; i < oe_frontend::kMetaSlots; ++i) {
        if (fe->h_meta_ev[i]) cudaEventDestroy(fe->h_meta_ev[i]);
        if (fe->h_meta[i]) cudaFreeHost(fe->h_meta[i]);
    }
    cudaFree(fe->d_tab);
    cudaFree(fe->d_rs);
    cudaFree(fe->d_rs_coefs);
    delete fe;
    return OE_OK;
}

int64_t oe_frontend_launch_count(const oe_frontend* fe) { return fe ? fe->launches : 0; }

int oe_frontend_set_kernel_timing(oe_frontend* fe, int32_t on) {
    if (!fe) return fail(OE_ERR_INVALID, "null frontend");
    if (on && !fe->ev_begin) {
        OE_CUDA(cudaSetDevice(fe->device));
        OE_CUDA(cudaEventCreate(&fe->ev_begin));
        OE_CUDA(cudaEventCreate(&fe->ev_end));
        OE_CUDA(cudaEventCreate(&fe->ev_step[0]));
        OE_CUDA(cudaEventCreate(&fe->ev_step[1]));
    }
    fe->timing = on != 0;
    fe->timed = false;
    return OE_OK;
}

int oe_frontend_fbank_kernel_ms(oe_frontend* fe, float* ms) {
    if (!fe || !ms) return fail(OE_ERR_INVALID, "null pointer");
    if (!fe->timed) return fail(OE_ERR_INVALID, "no timed fbank kernel: enable oe_frontend_set_kernel_timing before oe_fbank_batch");
    OE_CUDA(cudaEventSynchronize(fe->ev_end));
    OE_CUDA(cudaEventElapsedTime(ms, fe->ev_begin, fe->ev_end));
    return OE_OK;
}

int oe_frontend_step_ms(oe_frontend* fe, float* ms) {
    if (!fe || !ms) return fail(OE_ERR_INVALID, "null pointer");
    if (!fe->timed) return fail(OE_ERR_INVALID, "no timed call: enable oe_frontend_set_kernel_timing before oe_fbank_batch");
    OE_CUDA(cudaEventSynchronize(fe->ev_step[1]));
    OE_CUDA(cudaEventElapsedTime(ms, fe->ev_step[0], fe->ev_step[1]));
    return OE_OK;
}

int oe_frontend_get_tables(const oe_frontend* fe, float* window, float* mel) {
    if (!fe) return fail(OE_ERR_INVALID, "null frontend");
    if (window) memcpy(window, fe->window.data(), fe->window.size() * sizeof(float));
    if (mel) memcpy(mel, fe->mel.data(), fe->mel.size() * sizeof(float));
    return OE_OK;
}

// ---- window coalescing: streaming front-ends feed a long recording as many short windows (BASELINE config 5: 7 500
// windows of 16 frames, 2 800 samples each, 2 560 apart); a 16-frame window fills half a 32-frame tile.  When
// consecutive utterances of a batch are CONTINUATIONS of each other -- the next one starts exactly `frames * shift`
// samples behind the previous one in the waveform buffer and right behind its rows in the output -- and nothing in the
// batch is per-utterance (no normalisation, masks, frame maps, dithers, resampling, padding rows), the run is the same
// set of frames as one long utterance: fbank is frame-independent.  The batch is rewritten that way on the host, the
// kernels see full tiles.  Frame counts are still reported per original utterance. ----
struct Coalesced {
    oe_batch b;
    std::vector<int64_t> off, rows;
    std::vector<int32_t> len;
    std::vector<int32_t> frames_orig;
};
static bool coalesce_windows(const oe_frontend* fe, const oe_batch* bt, Coalesced& c) {
    if (!fe || !bt || fe->no_coalesce || bt->batch < 4 || !bt->wav_offsets || !bt->wav_lens || !bt->out_rows) return false;
    if (bt->wav_dtype != OE_WAV_I16 && bt->wav_dtype != OE_WAV_F32) return false;
    if (bt->norm_mode != OE_NORM_NONE || bt->n_tmask || bt->n_fmask || bt->frame_map || bt->resample_ids ||
        bt->feature_dither != 0.f || bt->wav_dither != 0.f)
        return false;
    const int B = bt->batch, shift = fe->cfg.frame_shift;
    c.frames_orig.resize(B);
    for (int b = 0; b < B; ++b) {
        if (bt->wav_lens[b] < 0 || bt->wav_offsets[b] < 0 || (bt->wav_offsets[b] & 7)) return false;    // plan() reports these on the caller's batch
        c.frames_orig[b] = oe_num_frames(fe, bt->wav_lens[b]);
        if (bt->out_nrows && bt->out_nrows[b] != c.frames_orig[b]) return false;
    }
    c.off.clear();
    c.rows.clear();
    c.len.clear();
    for (int b = 0; b < B;) {
        int e = b;                                                      // run [b, e]
        while (e + 1 < B && c.frames_orig[e] > 0 && c.frames_orig[e + 1] > 0 &&
               bt->wav_offsets[e + 1] == bt->wav_offsets[e] + (int64_t)shift * c.frames_orig[e] &&
               bt->out_rows[e + 1] == bt->out_rows[e] + c.frames_orig[e] &&
               bt->wav_offsets[e + 1] + bt->wav_lens[e + 1] - bt->wav_offsets[b] <= INT32_MAX)
            ++e;
        c.off.push_back(bt->wav_offsets[b]);
        c.rows.push_back(bt->out_rows[b]);
        c.len.push_back((int32_t)(bt->wav_offsets[e] + bt->wav_lens[e] - bt->wav_offsets[b]));
        b = e + 1;
    }
    if ((int)c.off.size() * 2 > B) return false;                       // not worth it
    c.b = *bt;
    c.b.batch = (int32_t)c.off.size();
    c.b.wav_offsets = c.off.data();
    c.b.wav_lens = c.len.data();
    c.b.out_rows = c.rows.data();
    c.b.out_nrows = nullptr;
    c.b.out_frames = nullptr;
    return true;
}

int oe_fbank_workspace_bytes(const oe_frontend* fe, const oe_batch* batch, size_t* bytes) {
    if (!bytes) return fail(OE_ERR_INVALID, "null bytes");
    Meta M;
    Coalesced cz;
    const int rc = plan(fe, coalesce_windows(fe, batch, cz) ? &cz.b : batch, M, nullptr);
    if (rc != OE_OK) return rc;
    *bytes = M.total + 256;
    return OE_OK;
}

// Next slot of the handle's pinned metadata ring, at least `bytes` large; waits (normally not at all) until the copy
// that last read the slot has completed.
static int meta_slot(oe_frontend* fe, size_t bytes, unsigned char** out, int* slot_out) {
    const int slot = fe->h_meta_next;
    fe->h_meta_next = (slot + 1) % oe_frontend::kMetaSlots;
    if (!fe->h_meta_ev[slot]) OE_CUDA(cudaEventCreateWithFlags(&fe->h_meta_ev[slot], cudaEventDisableTiming));
    else OE_CUDA(cudaEventSynchronize(fe->h_meta_ev[slot]));
    if (fe->h_meta_cap[slot] < bytes) {
        if (fe->h_meta[slot]) OE_CUDA(cudaFreeHost(fe->h_meta[slot]));
        fe->h_meta[slot] = nullptr;
        fe->h_meta_cap[slot] = 0;
        const size_t cap = align_up(bytes + bytes / 2 + 4096, 4096);
        OE_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&fe->h_meta[slot]), cap, cudaHostAllocMapped));
        OE_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&fe->h_meta_dev[slot]), fe->h_meta[slot], 0));
        fe->h_meta_cap[slot] = cap;
    }
    *out = fe->h_meta[slot];
    *slot_out = slot;
    return OE_OK;
}

// Scalars and device pointers of an oe_batch that the launch sequence needs (no host arrays): what a prepared batch keeps.
struct LaunchInfo {
    int B;
    int wav_dtype, norm_mode, n_tmask, n_fmask, cmvn_on_padding;
    int64_t out_pitch;
    bool has_map, any_rs;
    int max_pad;                 // single pass: most padding rows behind an utterance's last tile
    const float* d_cmvn_mean;
    const float* d_cmvn_istd;
    double* d_stats;
    float feature_dither, wav_dither;
    unsigned long long dither_seed;
};

// Fills the metadata block `hm` (M.meta_bytes) and the launch scalars from the caller's batch description.
static int pack_meta(const oe_frontend* fe, const oe_batch* bt, const Meta& M, const std::vector<int32_t>& frames,
                     unsigned char* hm, LaunchInfo& L) {
    const int B = bt->batch;
    auto i64 = [&](size_t off) { return reinterpret_cast<int64_t*>(hm + off); };
    auto i32 = [&](size_t off) { return reinterpret_cast<int32_t*>(hm + off); };
    int64_t fp = 0, rp = 0;
    int tp = 0;
    L.max_pad = 0;
    L.any_rs = false;
    for (int b = 0; b < B; ++b) {
        const int nfr = frames[b];
        const int nrows = bt->out_nrows ? bt->out_nrows[b] : nfr;
        i64(M.wav_off)[b] = bt->wav_offsets[b];
        i64(M.out_row)[b] = bt->out_rows[b];
        i64(M.frame_prefix)[b] = fp;
        i64(M.row_prefix)[b] = rp;
        i64(M.map_off)[b] = bt->frame_map ? bt->frame_map_offsets[b] : 0;
        i32(M.wav_len)[b] = bt->wav_lens[b];
        i32(M.n_frames)[b] = nfr;
        i32(M.n_rows)[b] = M.two_phase ? nfr : nrows;
        i32(M.tile_prefix)[b] = tp;
        const bool rs = bt->resample_ids && bt->resample_ids[b] >= 0;
        i32(M.rs_mode)[b] = rs ? (bt->resample_ids[b] == fe->rs_fast_9_10 ? 1 : 2) : 0;
        L.any_rs |= rs;
        L.max_pad = std::max(L.max_pad, nrows - (nfr + oe::kTileFrames - 1) / oe::kTileFrames * oe::kTileFrames);
        fp += nfr;
        rp += nrows;
        tp += (nfr + oe::kTileFrames - 1) / oe::kTileFrames;
    }
    i64(M.frame_prefix)[B] = fp;
    i64(M.row_prefix)[B] = rp;
    i32(M.tile_prefix)[B] = tp;
    if (bt->n_tmask) memcpy(i32(M.tmask), bt->tmask, 8 * (size_t)B * bt->n_tmask);
    if (bt->n_fmask) memcpy(i32(M.fmask), bt->fmask, 8 * (size_t)B * bt->n_fmask);
    if (bt->frame_map) {
        for (int b = 0; b < B; ++b)
            for (int t = 0; t < frames[b]; ++t) {
                const int32_t s = bt->frame_map[bt->frame_map_offsets[b] + t];
                if (s < 0 || s >= frames[b]) return fail(OE_ERR_INVALID, "frame_map[%d][%d]=%d out of range", b, t, s);
            }
        memcpy(i32(M.fmap), bt->frame_map, 4 * (size_t)M.total_map);
    }
    L.B = B;
    L.wav_dtype = bt->wav_dtype;
    L.norm_mode = bt->norm_mode;
    L.n_tmask = bt->n_tmask;
    L.n_fmask = bt->n_fmask;
    L.cmvn_on_padding = bt->cmvn_on_padding;
    L.out_pitch = bt->out_pitch;
    L.has_map = bt->frame_map != nullptr;
    L.d_cmvn_mean = bt->d_cmvn_mean;
    L.d_cmvn_istd = bt->d_cmvn_istd;
    L.d_stats = bt->d_stats;
    L.feature_dither = bt->feature_dither;
    L.wav_dither = bt->wav_dither;
    L.dither_seed = bt->dither_seed;
    return OE_OK;
}

static int launch_batch(oe_frontend* fe, const Meta& M, const LaunchInfo& L, const unsigned char* meta, const void* d_wav,
                        float* d_out, void* d_ws, size_t ws_bytes, cudaStream_t stream);

// ring slot -> device memory through oe_fetch_kernel (see there), then the slot's reuse event
static cudaError_t fetch_small(oe_frontend* fe, int slot, void* d_dst, size_t bytes, cudaStream_t stream) {
    const int n16 = (int)((bytes + 15) / 16);
    if (n16 > 0) {
        oe::oe_fetch_kernel<<<std::min(8, (n16 + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<const uint4*>(fe->h_meta_dev[slot]), reinterpret_cast<uint4*>(d_dst), n16);
        ++fe->launches;
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaEventRecord(fe->h_meta_ev[slot], stream);
    return e;
}

int oe_upload_small(oe_frontend* fe, const void* host, size_t bytes, void* d_dst, oe_stream stream) {
    if (!fe || (bytes && (!host || !d_dst))) return fail(OE_ERR_INVALID, "null pointer");
    if (reinterpret_cast<uintptr_t>(d_dst) & 15) return fail(OE_ERR_INVALID, "d_dst must be 16-byte aligned");
    if (bytes == 0) return OE_OK;
    OE_CUDA(cudaSetDevice(fe->device));
    unsigned char* hm = nullptr;
    int slot = 0;
    const int rc = meta_slot(fe, (bytes + 15) / 16 * 16, &hm, &slot);
    if (rc != OE_OK) return rc;
    memcpy(hm, host, bytes);
    OE_CUDA(fetch_small(fe, slot, d_dst, bytes, (cudaStream_t)stream));
    return OE_OK;
}

int oe_fbank_batch(oe_frontend* fe, const oe_batch* bt, const void* d_wav, float* d_out, void* d_ws,
                   size_t ws_bytes, oe_stream stream_) {
    Meta M;
    std::vector<int32_t> frames;
    Coalesced cz;
    const oe_batch* const orig = bt;
    if (coalesce_windows(fe, bt, cz)) bt = &cz.b;
    int rc = plan(fe, bt, M, &frames);
    if (rc != OE_OK) return rc;
    const int B = bt->batch;
    if (orig->out_frames) for (int b = 0; b < orig->batch; ++b) orig->out_frames[b] = bt == orig ? frames[b] : cz.frames_orig[b];
    if (B == 0) return OE_OK;
    // the metadata is packed straight into a slot of the handle's pinned ring
    OE_CUDA(cudaSetDevice(fe->device));
    unsigned char* hm = nullptr;
    int hslot = 0;
    rc = meta_slot(fe, M.meta_bytes, &hm, &hslot);
    if (rc != OE_OK) return rc;
    LaunchInfo L;
    rc = pack_meta(fe, bt, M, frames, hm, L);
    if (rc != OE_OK) return rc;
    return launch_batch(fe, M, L, nullptr, d_wav, d_out, d_ws, ws_bytes, (cudaStream_t)stream_) ;
}

// ---- prepared batches: validate + pack once, launch many times (or later) with no host work beyond one memcpy ----
struct oe_prepared {
    Meta M;
    LaunchInfo L;
    std::vector<int32_t> frames;
    std::vector<unsigned char> meta;
};

int oe_batch_prepare(oe_frontend* fe, const oe_batch* bt, oe_prepared** out) {
    if (!out) return fail(OE_ERR_INVALID, "null out");
    *out = nullptr;
    oe_prepared* p = new oe_prepared();
    Coalesced cz;
    const bool merged = coalesce_windows(fe, bt, cz);
    if (merged) bt = &cz.b;
    std::vector<int32_t> frames;
    int rc = plan(fe, bt, p->M, &frames);
    if (rc == OE_OK) {
        p->meta.assign(p->M.meta_bytes, 0);
        rc = pack_meta(fe, bt, p->M, frames, p->meta.data(), p->L);
        p->frames = merged ? cz.frames_orig : frames;                  // reported per ORIGINAL utterance
    }
    if (rc != OE_OK) {
        delete p;
        return rc;
    }
    *out = p;
    return OE_OK;
}

int oe_prepared_destroy(oe_prepared* p) {
    delete p;
    return OE_OK;
}

size_t oe_prepared_workspace_bytes(const oe_prepared* p) { return p ? p->M.total + 256 : 0; }
const int32_t* oe_prepared_frames(const oe_prepared* p) { return p ? p->frames.data() : nullptr; }

int oe_fbank_run(oe_frontend* fe, const oe_prepared* p, const void* d_wav, float* d_out, void* d_ws, size_t ws_bytes,
                 oe_stream stream) {
    if (!fe || !p) return fail(OE_ERR_INVALID, "null frontend or prepared batch");
    if (p->L.B == 0) return OE_OK;
    OE_CUDA(cudaSetDevice(fe->device));
    return launch_batch(fe, p->M, p->L, p->meta.data(), d_wav, d_out, d_ws, ws_bytes, (cudaStream_t)stream);
}

// `meta` null: the block has just been packed into the ring slot handed out last (oe_fbank_batch); otherwise it is
// copied into the next slot first.
static int launch_batch_impl(oe_frontend* fe, const Meta& M, const LaunchInfo& L, const unsigned char* meta, const void* d_wav,
                             float* d_out, void* d_ws, size_t ws_bytes, cudaStream_t stream);
static int launch_batch(oe_frontend* fe, const Meta& M, const LaunchInfo& L, const unsigned char* meta, const void* d_wav,
                        float* d_out, void* d_ws, size_t ws_bytes, cudaStream_t stream) {
    if (fe->timing) OE_CUDA(cudaEventRecord(fe->ev_step[0], stream));
    const int rc = launch_batch_impl(fe, M, L, meta, d_wav, d_out, d_ws, ws_bytes, stream);
    if (fe->timing && rc == OE_OK) OE_CUDA(cudaEventRecord(fe->ev_step[1], stream));
    return rc;
}
static int launch_batch_impl(oe_frontend* fe, const Meta& M, const LaunchInfo& L, const unsigned char* meta, const void* d_wav,
                             float* d_out, void* d_ws, size_t ws_bytes, cudaStream_t stream) {
    const LaunchInfo* const bt = &L;            // same field names as oe_batch below
    const int B = L.B, F = fe->cfg.num_mel_bins;
    if (!d_wav) return fail(OE_ERR_INVALID, "null d_wav");
    if (!d_out && !bt->d_stats) return fail(OE_ERR_INVALID, "nothing to produce: d_out and d_stats are both null");
    if (!d_ws || ws_bytes < M.total) return fail(OE_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", M.total, ws_bytes);
    if ((reinterpret_cast<uintptr_t>(d_ws) & 15) || (reinterpret_cast<uintptr_t>(d_wav) & 15))
        return fail(OE_ERR_INVALID, "d_wav and d_workspace must be 16-byte aligned");
    const bool inplace = M.inplace_ok && d_out != nullptr;
    if (inplace && (reinterpret_cast<uintptr_t>(d_out) & 15))
        return fail(OE_ERR_INVALID, "d_out must be 16-byte aligned (rows are finished in place with 16-byte accesses)");
    unsigned char* hm = nullptr;
    int hslot = 0;
    if (meta) {
        const int rc = meta_slot(fe, M.meta_bytes, &hm, &hslot);
        if (rc != OE_OK) return rc;
        memcpy(hm, meta, M.meta_bytes);
    } else {
        hslot = (fe->h_meta_next + oe_frontend::kMetaSlots - 1) % oe_frontend::kMetaSlots;
        hm = fe->h_meta[hslot];
    }
    unsigned char* ws = reinterpret_cast<unsigned char*>(d_ws);
    // small batches: the metadata rides in the descriptor kernel's parameters (oe_tile_desc_inline_kernel)
    const bool inline_meta = M.meta_bytes <= (size_t)oe::kInlineMeta && M.total_tiles > 0 && !fe->no_inline_meta;
    if (inline_meta) OE_CUDA(cudaEventRecord(fe->h_meta_ev[hslot], stream));      // the ring slot is free again at once
    else OE_CUDA(fetch_small(fe, hslot, ws, M.meta_bytes, stream));

    const int64_t pitch = bt->out_pitch ? bt->out_pitch : F;
    oe::FbankParams P;
    memset(&P, 0, sizeof(P));
    P.wav = d_wav;
    const int64_t* d_wav_off = reinterpret_cast<const int64_t*>(ws + M.wav_off);
    const int32_t* d_n_frames = reinterpret_cast<const int32_t*>(ws + M.n_frames);
    P.tiles = reinterpret_cast<const oe::TileDesc*>(ws + M.tiles);
    const int32_t* d_tile_prefix = reinterpret_cast<const int32_t*>(ws + M.tile_prefix);
    P.total_tiles = M.total_tiles;
    if (fe->std_mel) memcpy(P.mel_w, fe->mel_w_std, sizeof(P.mel_w));
    for (int which = 0; which < 2; ++which) {
        const int id = which == 0 ? fe->rs_fast_9_10 : fe->rs_fast_11_10;
        if (id >= 0) memcpy(P.rs_coef[which], fe->rs_coefs.data() + fe->rs[id].coef_off, sizeof(float) * fe->rs[id].neu * fe->rs[id].taps);
    }
    P.tab = fe->d_tab;
    P.tile_stats = M.need_stats ? reinterpret_cast<float*>(ws + M.tile_stats) : nullptr;
    const int64_t* d_tile_out_row;
    if (inplace) {
        // raw rows go straight to their final place; oe_finalize2_kernel completes them there
        P.out = d_out;
        P.keep_out_in_l2 = 1;
        d_tile_out_row = reinterpret_cast<const int64_t*>(ws + M.out_row);
        P.pitch = pitch;
    } else if (M.two_phase) {
        P.out = (d_out || !M.k2) ? reinterpret_cast<float*>(ws + M.raw) : nullptr;   // statistics only: no rows at all
        P.keep_out_in_l2 = 1;
        d_tile_out_row = reinterpret_cast<const int64_t*>(ws + M.frame_prefix);
        P.pitch = F;
    } else {
        P.out = d_out;
        d_tile_out_row = reinterpret_cast<const int64_t*>(ws + M.out_row);
        P.pitch = pitch;
        P.tmask = reinterpret_cast<const int32_t*>(ws + M.tmask);
        P.fmask = reinterpret_cast<const int32_t*>(ws + M.fmask);
        P.n_tmask = bt->n_tmask;
        P.n_fmask = bt->n_fmask;
        P.cmvn_mean = bt->d_cmvn_mean;
        P.cmvn_istd = bt->d_cmvn_istd;
        P.cmvn_on_pad = bt->cmvn_on_padding;
    }
    const int grid = M.feats ? 0 : std::min(M.total_tiles, 2 * fe->sm_count);
    bool stats_by_completion = false;
    if (M.k2) {
        P.sched = reinterpret_cast<int32_t*>(ws + M.sched);
        if (bt->d_stats) {
            P.stat_acc = reinterpret_cast<unsigned long long*>(ws + M.stat_acc);
            // with a completion kernel behind the fbank kernel, that kernel's first block converts the sums
            stats_by_completion = inplace && M.total_rows > 0 && M.total_tiles > 0;
            if (!stats_by_completion) {
                P.d_stats = bt->d_stats;
                P.stat_count = (double)M.total_frames;
            }
        }
    }
    // float4 row stores: gen-1 kernel needs dense rows (pitch == F), gen-2 any 16-byte aligned pitch
    P.out_vec = ((M.k2 ? P.pitch % 4 == 0 : P.pitch == F) && F % 4 == 0 && !(reinterpret_cast<uintptr_t>(P.out) & 15)) ? 1 : 0;
    if (M.total_tiles > 0) {
        oe::TileDescParams T;
        memset(&T, 0, sizeof(T));
        T.tile_prefix = d_tile_prefix;
        T.wav_off = d_wav_off;
        T.wav_len = reinterpret_cast<const int32_t*>(ws + M.wav_len);
        T.n_frames = d_n_frames;
        T.n_rows = reinterpret_cast<const int32_t*>(ws + M.n_rows);
        T.out_row = d_tile_out_row;
        T.rs_mode = reinterpret_cast<const int32_t*>(ws + M.rs_mode);
        T.tiles = reinterpret_cast<oe::TileDesc*>(ws + M.tiles);
        T.B = B;
        T.total_tiles = M.total_tiles;
        T.feats = M.feats ? 1 : 0;
        T.sched = P.sched;
        T.stat_acc = P.stat_acc;
        T.n_acc = 2 * F;
        ++fe->launches;
        if (inline_meta) {
            oe::TileDescInlineParams Q;
            Q.p = T;
            auto rel = [&](const void* dptr) { return reinterpret_cast<uintptr_t>(dptr) - reinterpret_cast<uintptr_t>(ws); };
            Q.p.tile_prefix = reinterpret_cast<const int32_t*>(rel(T.tile_prefix));
            Q.p.wav_off = reinterpret_cast<const int64_t*>(rel(T.wav_off));
            Q.p.wav_len = reinterpret_cast<const int32_t*>(rel(T.wav_len));
            Q.p.n_frames = reinterpret_cast<const int32_t*>(rel(T.n_frames));
            Q.p.n_rows = reinterpret_cast<const int32_t*>(rel(T.n_rows));
            Q.p.out_row = reinterpret_cast<const int64_t*>(rel(T.out_row));
            Q.p.rs_mode = reinterpret_cast<const int32_t*>(rel(T.rs_mode));
            Q.ws_meta = reinterpret_cast<uint4*>(ws);
            Q.meta_bytes = (int)M.meta_bytes;
            memcpy(Q.meta, hm, M.meta_bytes);
            // plain stream order (like oe_fetch_kernel): the workspace may still be read by the previous call's kernels
            oe::oe_tile_desc_inline_kernel<<<dim3((M.total_tiles + 127) / 128), dim3(128), 0, stream>>>(Q);
        } else {
            OE_CUDA(launch_dep(oe::oe_tile_desc_kernel, dim3((M.total_tiles + 127) / 128), dim3(128), 0, stream, T));
        }
        OE_CUDA(cudaGetLastError());
    }
    int n_stat_partials = 0;
    if (M.feats) {
        if ((M.need_stats || bt->d_stats) && M.total_tiles > 0) {
            oe::FeatStatsParams S;
            S.feats = reinterpret_cast<const float*>(d_wav);
            S.tiles = P.tiles;
            S.tile_stats = P.tile_stats;
            S.cta_stats = bt->d_stats ? reinterpret_cast<double*>(ws + M.stat_partial) : nullptr;
            S.F = F;
            S.total_tiles = M.total_tiles;
            n_stat_partials = 3 * std::min(M.total_tiles, fe->sm_count * 8);
            ++fe->launches;
            OE_CUDA(launch_dep(oe::oe_feat_tile_stats_kernel, dim3(std::min(M.total_tiles, fe->sm_count * 8)), dim3(oe::kMaxMel), 0, stream, S));
            OE_CUDA(cudaGetLastError());
        }
    } else if (M.total_tiles > 0) {
        const bool f32 = bt->wav_dtype == OE_WAV_F32;
        if (bt->d_stats && !M.k2) {
            P.cta_stats = reinterpret_cast<double*>(ws + M.stat_partial);
            n_stat_partials = 3 * grid;
        }
        const bool any_rs = bt->any_rs;
        ++fe->launches;
        P.wav_dither = bt->wav_dither;
        P.dither_seed = bt->dither_seed;
        if (fe->timing) OE_CUDA(cudaEventRecord(fe->ev_begin, stream));   // an event between two kernels also ends their PDL overlap
        if (bt->wav_dither != 0.f) {
            if (f32) OE_CUDA(launch_dep(oe::k2::oe_fbank2_kernel<true, false, true>, dim3(grid), dim3(oe::kThreads), oe::k2::Smem<true, false>::End, stream, P));
            else OE_CUDA(launch_dep(oe::k2::oe_fbank2_kernel<false, false, true>, dim3(grid), dim3(oe::kThreads), oe::k2::Smem<false, false>::End, stream, P));
        } else if (M.k2) {
            if (f32) OE_CUDA(launch_dep(oe::k2::oe_fbank2_kernel<true, false>, dim3(grid), dim3(oe::kThreads), oe::k2::Smem<true, false>::End, stream, P));
            else if (any_rs) OE_CUDA(launch_dep(oe::k2::oe_fbank2_kernel<false, true>, dim3(grid), dim3(oe::kThreads), oe::k2::Smem<false, true>::End, stream, P));
            else OE_CUDA(launch_dep(oe::k2::oe_fbank2_kernel<false, false>, dim3(grid), dim3(oe::kThreads), oe::k2::Smem<false, false>::End, stream, P));
        } else if (fe->std_mel) {
            if (f32) OE_CUDA(launch_dep(oe::oe_fbank_kernel<true, true, false>, dim3(grid), dim3(oe::kThreads), fe->fbank_smem, stream, P));
            else if (any_rs) OE_CUDA(launch_dep(oe::oe_fbank_kernel<false, true, true>, dim3(grid), dim3(oe::kThreads), fe->fbank_smem, stream, P));
            else OE_CUDA(launch_dep(oe::oe_fbank_kernel<false, true, false>, dim3(grid), dim3(oe::kThreads), fe->fbank_smem, stream, P));
        } else {
            if (f32) OE_CUDA(launch_dep(oe::oe_fbank_kernel<true, false, false>, dim3(grid), dim3(oe::kThreads), fe->fbank_smem, stream, P));
            else if (any_rs) OE_CUDA(launch_dep(oe::oe_fbank_kernel<false, false, true>, dim3(grid), dim3(oe::kThreads), fe->fbank_smem, stream, P));
            else OE_CUDA(launch_dep(oe::oe_fbank_kernel<false, false, false>, dim3(grid), dim3(oe::kThreads), fe->fbank_smem, stream, P));
        }
        if (fe->timing) {
            OE_CUDA(cudaEventRecord(fe->ev_end, stream));
            fe->timed = true;
        }
        OE_CUDA(cudaGetLastError());
    }
    if (!M.two_phase && d_out && M.pad_rows > 0) {
        const int max_pad = bt->max_pad;
        oe::PadFillParams Q;
        Q.n_frames = d_n_frames;
        Q.n_rows = reinterpret_cast<const int32_t*>(ws + M.n_rows);
        Q.out_row = reinterpret_cast<const int64_t*>(ws + M.out_row);
        Q.out = d_out;
        Q.pitch = pitch;
        Q.cmvn_mean = (bt->d_cmvn_mean && bt->cmvn_on_padding) ? bt->d_cmvn_mean : nullptr;
        Q.cmvn_istd = bt->d_cmvn_istd;
        Q.F = F;
        ++fe->launches;
        OE_CUDA(launch_dep(oe::oe_pad_fill_kernel, dim3(B, (max_pad + oe::kPadRows - 1) / oe::kPadRows), dim3(256), 0, stream, Q));
    }
    if (inplace) {
        if (M.total_rows == 0) return OE_OK;
        oe::Finalize2Params Z;
        memset(&Z, 0, sizeof(Z));
        Z.out = d_out;
        Z.pitch = pitch;
        Z.out_row = reinterpret_cast<const int64_t*>(ws + M.out_row);
        Z.row_prefix = reinterpret_cast<const int64_t*>(ws + M.row_prefix);
        Z.n_frames = d_n_frames;
        Z.tile_prefix = d_tile_prefix;
        Z.tile_stats = bt->norm_mode != OE_NORM_NONE ? P.tile_stats : nullptr;
        Z.tmask = reinterpret_cast<const int32_t*>(ws + M.tmask);
        Z.fmask = reinterpret_cast<const int32_t*>(ws + M.fmask);
        Z.n_tmask = bt->n_tmask;
        Z.n_fmask = bt->n_fmask;
        Z.cmvn_mean = bt->d_cmvn_mean;
        Z.cmvn_istd = bt->d_cmvn_istd;
        Z.cmvn_on_pad = bt->cmvn_on_padding;
        Z.dither_a = bt->feature_dither;
        Z.dither_seed = bt->dither_seed;
        // parts per utterance: every part redoes the statistics merge (two dependent rounds of L2 loads), so as few as
        // fill the machine: <= 1024 rows each, and at least 1.5 blocks per SM in total (benchmark batch, same box:
        // 1 part 163.6 us per step, 2: 164.6, 3: 167.3, 4: 172.5, 8: 186.6)
        int parts = std::max(1, (M.max_rows + 1023) / 1024);
        parts = std::max(parts, (3 * fe->sm_count / 2 + B - 1) / B);
        parts = std::min(parts, std::max(1, (M.max_rows + 31) / 32));
        if (fe->fin2_parts > 0) parts = fe->fin2_parts;
        if (bt->has_map) {                      // spec_sub in place: one block walks the utterance from its last rows down
            parts = 1;
            Z.frame_map = reinterpret_cast<const int32_t*>(ws + M.fmap);
            Z.map_off = reinterpret_cast<const int64_t*>(ws + M.map_off);
        }
        Z.parts = parts;
        if (stats_by_completion) {
            Z.stat_acc = P.stat_acc;
            Z.d_stats = bt->d_stats;
            Z.stat_count = (double)M.total_frames;
        }
        ++fe->launches;
        if (bt->has_map) {
            if (Z.dither_a != 0.f) OE_CUDA(launch_dep(oe::oe_finalize2_kernel<true, true>, dim3(B, parts), dim3(oe::kFin2Threads), 0, stream, Z));
            else OE_CUDA(launch_dep(oe::oe_finalize2_kernel<false, true>, dim3(B, parts), dim3(oe::kFin2Threads), 0, stream, Z));
        } else if (Z.dither_a != 0.f) OE_CUDA(launch_dep(oe::oe_finalize2_kernel<true, false>, dim3(B, parts), dim3(oe::kFin2Threads), 0, stream, Z));
        else OE_CUDA(launch_dep(oe::oe_finalize2_kernel<false, false>, dim3(B, parts), dim3(oe::kFin2Threads), 0, stream, Z));
        OE_CUDA(cudaGetLastError());
        return OE_OK;
    }
    {
        const bool want_utt = bt->norm_mode != OE_NORM_NONE && d_out != nullptr;
        const bool want_glob = bt->d_stats && n_stat_partials > 0;
        if (want_utt || want_glob) {
            oe::UttStatsParams U;
            memset(&U, 0, sizeof(U));
            U.tile_stats = P.tile_stats;
            U.tile_prefix = d_tile_prefix;
            U.n_frames = d_n_frames;
            U.utt_mean = reinterpret_cast<float*>(ws + M.utt_mean);
            U.utt_std = reinterpret_cast<float*>(ws + M.utt_std);
            U.F = F;
            U.B = want_utt ? B : 0;
            int glob_blocks = 0;
            if (want_glob) {
                U.partial = reinterpret_cast<const double*>(ws + M.stat_partial);
                U.stats = bt->d_stats;
                U.count = (double)M.total_frames;
                U.n_partials = n_stat_partials;
                glob_blocks = (2 * F + 1 + oe::kGlobStats - 1) / oe::kGlobStats;
            }
            ++fe->launches;
            if (F <= 80) OE_CUDA(launch_dep(oe::oe_utt_stats_kernel<640, 2>, dim3(U.B + glob_blocks), dim3(F, oe::kUttSlices), 0, stream, U));
            else OE_CUDA(launch_dep(oe::oe_utt_stats_kernel<oe::kMaxMel * oe::kUttSlices, 1>, dim3(U.B + glob_blocks), dim3(F, oe::kUttSlices), 0, stream, U));
            OE_CUDA(cudaGetLastError());
        }
    }
    if (M.two_phase && d_out && M.total_rows > 0) {
        oe::FinalizeParams Z;
        memset(&Z, 0, sizeof(Z));
        Z.raw = M.feats ? reinterpret_cast<const float*>(d_wav) : reinterpret_cast<const float*>(ws + M.raw);
        Z.frame_prefix = reinterpret_cast<const int64_t*>(ws + (M.feats ? M.wav_off : M.frame_prefix));
        Z.row_prefix = reinterpret_cast<const int64_t*>(ws + M.row_prefix);
        Z.n_frames = d_n_frames;
        Z.out_row = reinterpret_cast<const int64_t*>(ws + M.out_row);
        Z.out = d_out;
        Z.pitch = pitch;
        if (bt->norm_mode != OE_NORM_NONE) {
            Z.utt_mean = reinterpret_cast<const float*>(ws + M.utt_mean);
            Z.utt_std = reinterpret_cast<const float*>(ws + M.utt_std);
        }
        if (bt->has_map) {
            Z.frame_map = reinterpret_cast<const int32_t*>(ws + M.fmap);
            Z.map_off = reinterpret_cast<const int64_t*>(ws + M.map_off);
        }
        Z.tmask = reinterpret_cast<const int32_t*>(ws + M.tmask);
        Z.fmask = reinterpret_cast<const int32_t*>(ws + M.fmask);
        Z.n_tmask = bt->n_tmask;
        Z.n_fmask = bt->n_fmask;
        Z.cmvn_mean = bt->d_cmvn_mean;
        Z.cmvn_istd = bt->d_cmvn_istd;
        Z.cmvn_on_pad = bt->cmvn_on_padding;
        Z.dither_a = bt->feature_dither;
        Z.dither_seed = bt->dither_seed;
        Z.F = F;
        const bool vec = (F % 4 == 0) && (pitch % 4 == 0) && !(reinterpret_cast<uintptr_t>(Z.raw) & 15) &&
                         !(reinterpret_cast<uintptr_t>(d_out) & 15);
        dim3 zgrid((unsigned)B, (unsigned)((M.max_rows + oe::kFinRows - 1) / oe::kFinRows));
        ++fe->launches;
        const bool dith = Z.dither_a != 0.f;
        if (vec) OE_CUDA(dith ? launch_dep(oe::oe_finalize_kernel<4, true>, zgrid, dim3(256), 0, stream, Z)
                              : launch_dep(oe::oe_finalize_kernel<4, false>, zgrid, dim3(256), 0, stream, Z));
        else OE_CUDA(dith ? launch_dep(oe::oe_finalize_kernel<1, true>, zgrid, dim3(256), 0, stream, Z)
                          : launch_dep(oe::oe_finalize_kernel<1, false>, zgrid, dim3(256), 0, stream, Z));
        OE_CUDA(cudaGetLastError());
    }
    return OE_OK;
}

int oe_cmvn_apply(const float* d_x, float* d_y, int64_t rows, int32_t dim, const float* d_mean,
                  const float* d_istd, oe_stream stream) {
    if (rows < 0 || dim <= 0) return fail(OE_ERR_INVALID, "bad shape");
    if (rows == 0) return OE_OK;
    if (!d_x || !d_y || !d_mean) return fail(OE_ERR_INVALID, "null pointer");
    const int64_t n = rows * dim;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    oe::oe_cmvn_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_x, d_y, n, dim, d_mean, d_istd);
    OE_CUDA(cudaGetLastError());
    return OE_OK;
}

int oe_cmvn_conv_subsample(const float* d_x, int64_t pitch, int32_t B, int32_t T, int32_t F, const float* d_mean,
                           const float* d_istd, const float* d_w, const float* d_bias, int32_t odim, float* d_y,
                           oe_stream stream) {
    if (B < 0 || T < 0 || F < 3 || odim <= 0) return fail(OE_ERR_INVALID, "bad shape");
    if (pitch == 0) pitch = F;
    if (pitch < F) return fail(OE_ERR_INVALID, "pitch smaller than F");
    const int T1 = T >= 3 ? (T - 3) / 2 + 1 : 0, F1 = (F - 3) / 2 + 1;
    if (B == 0 || T1 == 0) return OE_OK;
    if (!d_x || !d_w || !d_y) return fail(OE_ERR_INVALID, "null pointer");
    if (d_istd && !d_mean) return fail(OE_ERR_INVALID, "istd without mean");
    const size_t smem = ((size_t)odim * 12 + (size_t)(2 * oe::kConvTT + 1) * F) * sizeof(float);
    if (smem > 200 * 1024) return fail(OE_ERR_UNSUPPORTED, "odim * 48 + F * 68 bytes of shared memory exceed 200 KB");
    if (smem > 48 * 1024) OE_CUDA(cudaFuncSetAttribute(oe::oe_conv_sub1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    oe::ConvSubParams P;
    P.x = d_x;
    P.pitch = pitch;
    P.B = B;
    P.T = T;
    P.F = F;
    P.odim = odim;
    P.T1 = T1;
    P.F1 = F1;
    P.w = d_w;
    P.bias = d_bias;
    P.cmvn_mean = d_mean;
    P.cmvn_istd = d_istd;
    P.y = d_y;
    oe::oe_conv_sub1_kernel<<<dim3((T1 + oe::kConvTT - 1) / oe::kConvTT, B), oe::kConvThreads, smem, (cudaStream_t)stream>>>(P);
    OE_CUDA(cudaGetLastError());
    return OE_OK;
}

int64_t oe_resample_out_len(int64_t n, int32_t orig, int32_t neu) {
    if (orig <= 0 || neu <= 0 || n < 0) return -1;
    return (neu * n + orig - 1) / orig;          // ceil(new*n/orig), functional.py:1427
}

int oe_add_resampler(oe_frontend* fe, int32_t orig, int32_t neu, const float* kernel, int32_t taps, int32_t* table_id) {
    if (!fe || !table_id) return fail(OE_ERR_INVALID, "null pointer");
    if (orig <= 0 || neu <= 0) return fail(OE_ERR_INVALID, "rates must be positive");
    for (size_t i = 0; i < fe->rs.size(); ++i)
        if (fe->rs[i].orig == orig && fe->rs[i].neu == neu && !kernel && fe->rs_builtin[i]) {
            *table_id = (int)i;
            return OE_OK;
        }
    // functional.py:1343-1398
    const double lowpass = 6.0, rolloff = 0.99;
    const double base = std::min(orig, neu) * rolloff;
    int width = (int)std::ceil(lowpass * orig / base);
    int ntaps = 2 * width + orig;
    if (kernel) {
        // any polyphase table in torchaudio's layout: kernel[new][2 * width + orig] (e.g. a long Kaiser design of sox quality)
        if (taps < orig || ((taps - orig) & 1)) return fail(OE_ERR_INVALID, "taps must be 2*width+orig with width >= 0; got %d for orig %d", taps, orig);
        ntaps = taps;
        width = (taps - orig) / 2;
    } else if ((size_t)neu * ntaps > (size_t)oe::kRsDirectCoefs) {
        // the built-in hann sinc of a ratio with a long period (441:160, speeds drawn from a continuous range) is not
        // tabulated: oe_resample evaluates it on the fly for utterances marked OE_RS_DIRECT
        return fail(OE_ERR_UNSUPPORTED, "the %d:%d table would hold %zu coefficients: pass table_ids[b] = OE_RS_DIRECT with "
                                        "orig_rates / new_rates instead", orig, neu, (size_t)neu * ntaps);
    }
    oe::RsTable t;
    t.orig = orig;
    t.neu = neu;
    t.taps = ntaps;
    t.width = width;
    if (fe->rs_coefs.size() + (size_t)neu * ntaps > (size_t)INT32_MAX) return fail(OE_ERR_UNSUPPORTED, "resampler tables too large");
    t.coef_off = (int)fe->rs_coefs.size();
    for (int p = 0; p < neu; ++p)
        for (int q = 0; q < ntaps; ++q) {
            float v;
            if (kernel) {
                v = kernel[(size_t)p * ntaps + q];
            } else {
                double tt = (-(double)p / neu + (double)(q - width) / orig) * base;
                tt = std::max(-lowpass, std::min(lowpass, tt));
                const double win = std::pow(std::cos(tt * oe::kPi / lowpass / 2.0), 2.0);
                const double ang = tt * oe::kPi;
                const double sinc = ang == 0.0 ? 1.0 : std::sin(ang) / ang;
                v = (float)(sinc * win * (base / orig));
            }
            fe->rs_coefs.push_back(v);
        }
    fe->rs.push_back(t);
    fe->rs_builtin.push_back(kernel == nullptr);
    if (width == 7 && neu == 10 && (orig == 9 || orig == 11)) {
        // the specialised kernel skips taps the hann-sinc formula makes zero: only valid if this table has them zero
        bool ok = true;
        for (int p = 0; p < neu && ok; ++p)
            for (int q = 0; q < ntaps && ok; ++q)
                if (!oe::rs_tap_nonzero(orig, neu, width, p, q) && std::fabs(fe->rs_coefs[t.coef_off + p * ntaps + q]) > 1e-12f) ok = false;
        // ... and the fused resampler of the gen-2 fbank kernel carries the torchaudio table as immediates: only a
        // table with exactly those bits is fused (anything else still has the stand-alone resampling kernels)
        for (int i = 0; i < neu * ntaps && ok; ++i) {
            const float baked = orig == 9 ? oe::rsbaked::coef<9>(i) : oe::rsbaked::coef<11>(i);
            if (memcmp(&baked, &fe->rs_coefs[t.coef_off + i], 4) != 0 && oe::rs_tap_nonzero(orig, neu, width, i / ntaps, i % ntaps)) ok = false;
        }
        if (ok) (orig == 9 ? fe->rs_fast_9_10 : fe->rs_fast_11_10) = (int)fe->rs.size() - 1;
    }
    OE_CUDA(cudaSetDevice(fe->device));
    // device copies grow by doubling; a setup-time call: waiting for the device before the old block is freed is fine
    if (fe->rs.size() > fe->d_rs_cap || fe->rs_coefs.size() > fe->d_rs_coefs_cap) {
        OE_CUDA(cudaDeviceSynchronize());
        if (fe->d_rs) cudaFree(fe->d_rs);
        if (fe->d_rs_coefs) cudaFree(fe->d_rs_coefs);
        fe->d_rs = nullptr;
        fe->d_rs_coefs = nullptr;
        fe->d_rs_cap = std::max<size_t>(16, 2 * fe->rs.size());
        fe->d_rs_coefs_cap = std::max<size_t>(16384, 2 * fe->rs_coefs.size());
        OE_CUDA(cudaMalloc(&fe->d_rs, sizeof(oe::RsTable) * fe->d_rs_cap));
        OE_CUDA(cudaMalloc(&fe->d_rs_coefs, sizeof(float) * fe->d_rs_coefs_cap));
    }
    OE_CUDA(cudaMemcpy(fe->d_rs, fe->rs.data(), sizeof(oe::RsTable) * fe->rs.size(), cudaMemcpyHostToDevice));
    OE_CUDA(cudaMemcpy(fe->d_rs_coefs, fe->rs_coefs.data(), sizeof(float) * fe->rs_coefs.size(), cudaMemcpyHostToDevice));
    *table_id = (int)fe->rs.size() - 1;
    return OE_OK;
}

int oe_resampler_fusable(const oe_frontend* fe, int32_t table_id) {
    return fe && table_id >= 0 && (table_id == fe->rs_fast_9_10 || table_id == fe->rs_fast_11_10) ? 1 : 0;
}

int oe_mel_is_baked(const oe_frontend* fe) { return fe && fe->mel_baked && !fe->force_v1 ? 1 : 0; }

int oe_resample_workspace_bytes(const oe_frontend* fe, const oe_resample_batch* bt, size_t* bytes) {
    if (!fe || !bt || !bytes) return fail(OE_ERR_INVALID, "null pointer");
    const size_t B = (size_t)std::max(bt->batch, 0);
    *bytes = align_up(8 * B, 16) * 2 + align_up(4 * B, 16) * 5 + 256;
    return OE_OK;
}

int oe_resample(oe_frontend* fe, const oe_resample_batch* bt, const void* d_in, float* d_out, void* d_ws,
                size_t ws_bytes, oe_stream stream_) {
    if (!fe || !bt) return fail(OE_ERR_INVALID, "null pointer");
    const int B = bt->batch;
    if (B < 0) return fail(OE_ERR_INVALID, "negative batch");
    if (B == 0) return OE_OK;
    if (!bt->in_offsets || !bt->in_lens || !bt->table_ids || !bt->out_offsets) return fail(OE_ERR_INVALID, "null metadata");
    if (!d_in || !d_out) return fail(OE_ERR_INVALID, "null buffer");
    size_t need;
    oe_resample_workspace_bytes(fe, bt, &need);
    if (!d_ws || ws_bytes < need) return fail(OE_ERR_WORKSPACE, "workspace too small: need %zu bytes", need);
    const size_t a8 = align_up(8 * (size_t)B, 16), a4 = align_up(4 * (size_t)B, 16);
    const size_t hm_bytes = 2 * a8 + 5 * a4;
    OE_CUDA(cudaSetDevice(fe->device));
    unsigned char* hm = nullptr;
    int hslot = 0;
    {
        const int rc = meta_slot(fe, hm_bytes, &hm, &hslot);
        if (rc != OE_OK) return rc;
    }
    memset(hm, 0, hm_bytes);
    int64_t* in_off = reinterpret_cast<int64_t*>(hm);
    int64_t* out_off = reinterpret_cast<int64_t*>(hm + a8);
    int32_t* in_len = reinterpret_cast<int32_t*>(hm + 2 * a8);
    int32_t* tab = reinterpret_cast<int32_t*>(hm + 2 * a8 + a4);
    int32_t* out_len = reinterpret_cast<int32_t*>(hm + 2 * a8 + 2 * a4);
    int32_t* r_orig = reinterpret_cast<int32_t*>(hm + 2 * a8 + 3 * a4);
    int32_t* r_neu = reinterpret_cast<int32_t*>(hm + 2 * a8 + 4 * a4);
    int max_out = 0;
    for (int b = 0; b < B; ++b) {
        const int id = bt->table_ids[b];
        if (id >= (int)fe->rs.size()) return fail(OE_ERR_INVALID, "table_ids[%d]=%d is not registered", b, id);
        if (bt->in_lens[b] < 0) return fail(OE_ERR_INVALID, "negative length");
        in_off[b] = bt->in_offsets[b];
        out_off[b] = bt->out_offsets[b];
        in_len[b] = bt->in_lens[b];
        tab[b] = id;
        if (id == OE_RS_DIRECT) {
            if (!bt->orig_rates || !bt->new_rates || bt->orig_rates[b] <= 0 || bt->new_rates[b] <= 0)
                return fail(OE_ERR_INVALID, "table_ids[%d] = OE_RS_DIRECT needs positive orig_rates / new_rates", b);
            r_orig[b] = bt->orig_rates[b];
            r_neu[b] = bt->new_rates[b];
            out_len[b] = (int32_t)oe_resample_out_len(bt->in_lens[b], r_orig[b], r_neu[b]);
        } else {
            out_len[b] = id < 0 ? bt->in_lens[b] : (int32_t)oe_resample_out_len(bt->in_lens[b], fe->rs[id].orig, fe->rs[id].neu);
        }
        if (bt->out_lens) bt->out_lens[b] = out_len[b];
        max_out = std::max(max_out, out_len[b]);
    }
    if (max_out == 0) return OE_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned char* ws = reinterpret_cast<unsigned char*>(d_ws);
    OE_CUDA(fetch_small(fe, hslot, ws, hm_bytes, stream));
    oe::ResampleParams P;
    P.in = d_in;
    P.out = d_out;
    P.in_off = reinterpret_cast<const int64_t*>(ws);
    P.out_off = reinterpret_cast<const int64_t*>(ws + a8);
    P.in_len = reinterpret_cast<const int32_t*>(ws + 2 * a8);
    P.table_id = reinterpret_cast<const int32_t*>(ws + 2 * a8 + a4);
    P.out_len = reinterpret_cast<const int32_t*>(ws + 2 * a8 + 2 * a4);
    P.tables = fe->d_rs;
    P.coefs = fe->d_rs_coefs;
    P.orig = reinterpret_cast<const int32_t*>(ws + 2 * a8 + 3 * a4);
    P.neu = reinterpret_cast<const int32_t*>(ws + 2 * a8 + 4 * a4);
    const bool f32 = bt->wav_dtype == OE_WAV_F32;
    bool need_generic = false, need_9 = false, need_11 = false;
    for (int b = 0; b < B; ++b) {
        if (out_len[b] == 0) continue;
        if (tab[b] >= 0 && tab[b] == fe->rs_fast_9_10) need_9 = true;
        else if (tab[b] >= 0 && tab[b] == fe->rs_fast_11_10) need_11 = true;
        else need_generic = true;
    }
    if (need_generic) {
        ++fe->launches;
        dim3 grid((unsigned)std::min((max_out + 1023) / 1024, 1024), (unsigned)B);
        if (f32) oe::oe_resample_kernel<true><<<grid, 256, 0, stream>>>(P, fe->rs_fast_9_10, fe->rs_fast_11_10);
        else oe::oe_resample_kernel<false><<<grid, 256, 0, stream>>>(P, fe->rs_fast_9_10, fe->rs_fast_11_10);
    }
    for (int which = 0; which < 2; ++which) {
        if (!(which == 0 ? need_9 : need_11)) continue;
        const int id = which == 0 ? fe->rs_fast_9_10 : fe->rs_fast_11_10;
        ++fe->launches;
        oe::RsFastParams Q;
        Q.r = P;
        Q.table_id = id;
        memset(Q.coef, 0, sizeof(Q.coef));
        memcpy(Q.coef, fe->rs_coefs.data() + fe->rs[id].coef_off, sizeof(float) * fe->rs[id].neu * fe->rs[id].taps);
        dim3 grid((unsigned)((max_out + oe::kRsM * 10 - 1) / (oe::kRsM * 10)), (unsigned)B);
        if (which == 0) {
            if (f32) oe::oe_resample_fast_kernel<true, 9, 10, 7><<<grid, oe::kRsM, 0, stream>>>(Q);
            else oe::oe_resample_fast_kernel<false, 9, 10, 7><<<grid, oe::kRsM, 0, stream>>>(Q);
        } else {
            if (f32) oe::oe_resample_fast_kernel<true, 11, 10, 7><<<grid, oe::kRsM, 0, stream>>>(Q);
            else oe::oe_resample_fast_kernel<false, 11, 10, 7><<<grid, oe::kRsM, 0, stream>>>(Q);
        }
    }
    OE_CUDA(cudaGetLastError());
    return OE_OK;
}


}  // extern "C"

// ------------------------------------------------------------------------------------------
// Host-side planning: Python-`random`-compatible Mersenne Twister (CPython Modules/_randommodule.c
// semantics: genrand_uint32, random() = (a>>5, b>>6) 53-bit, getrandbits(k<=32) = genrand >> (32-k),
// Random._randbelow_with_getrandbits).
namespace {

struct PyMT {
    uint32_t* mt;     // 624 words
    uint32_t* pos;    // index
    uint32_t next() {
        constexpr int N = 624, M = 397;
        constexpr uint32_t MATRIX_A = 0x9908b0dfU, UPPER = 0x80000000U, LOWER = 0x7fffffffU;
        if (*pos >= (uint32_t)N) {
            int kk;
            uint32_t y;
            for (kk = 0; kk < N - M; kk++) {
                y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
                mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1U) ? MATRIX_A : 0U);
            }
            for (; kk < N - 1; kk++) {
                y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
                mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((y & 1U) ? MATRIX_A : 0U);
            }
            y = (mt[N - 1] & UPPER) | (mt[0] & LOWER);
            mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1U) ? MATRIX_A : 0U);
            *pos = 0;
        }
        uint32_t y = mt[(*pos)++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680U;
        y ^= (y << 15) & 0xefc60000U;
        y ^= (y >> 18);
        return y;
    }
    double random() {
        const uint32_t a = next() >> 5, b = next() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
    uint32_t randbelow(uint32_t n) {            // n >= 1
        int k = 0;
        for (uint32_t v = n; v; v >>= 1) ++k;   // n.bit_length()
        uint32_t r = next() >> (32 - k);
        while (r >= n) r = next() >> (32 - k);
        return r;
    }
    int32_t randint(int32_t a, int32_t b) { return a + (int32_t)randbelow((uint32_t)(b - a + 1)); }
};

}  // namespace

extern "C" {

int oe_plan_speeds(uint32_t* mt_state, int32_t n, double perturb_rate, const double* speeds_cfg,
                   int32_t n_speeds_cfg, const double* item_speeds, const uint8_t* active, double* out_speeds) {
    if (!mt_state || n < 0 || (n > 0 && (!item_speeds || !out_speeds))) return fail(OE_ERR_INVALID, "null pointer");
    double cfg[3] = {0.9, 1.1, 0.1};            // audio_processor.py:6-7
    int ncfg = 3;
    if (n_speeds_cfg > 0) {
        if (!speeds_cfg || (n_speeds_cfg != 1 && n_speeds_cfg < 3)) return fail(OE_ERR_INVALID, "speeds must be [fixed] or [start, end, step]");
        ncfg = n_speeds_cfg;
        for (int i = 0; i < ncfg && i < 3; ++i) cfg[i] = speeds_cfg[i];
    }
    if (ncfg > 1 && !(cfg[1] > cfg[0])) return fail(OE_ERR_INVALID, "speeds is wrong !");   // audio_processor.py:10
    PyMT g{mt_state, mt_state + 624};
    for (int i = 0; i < n; ++i) {
        out_speeds[i] = item_speeds[i];
        if (active && !active[i]) continue;
        if (g.random() < perturb_rate) {         // dataset.py:88
            if (ncfg > 1) {
                if (cfg[2] != 0.0) {             // randrange(int(s0/step), int(s0/step)+1) * step
                    const long long lo = (long long)(cfg[0] / cfg[2]);
                    out_speeds[i] = (double)(lo + (long long)g.randbelow(1)) * cfg[2];
                } else {
                    out_speeds[i] = cfg[0] + g.random() * (cfg[1] - cfg[0]);
                }
            } else {
                out_speeds[i] = cfg[0];
            }
        }
    }
    return OE_OK;
}

int oe_plan_augment(uint32_t* mt_state, int32_t n, const int32_t* frames, int32_t num_freq,
                    int32_t do_sub, int32_t sub_max_t, int32_t sub_num, int32_t do_aug, int32_t n_t,
                    int32_t n_f, int32_t max_t, int32_t max_f, int32_t* frame_map, int32_t* tmask, int32_t* fmask) {
    if (!mt_state || n < 0 || (n > 0 && !frames)) return fail(OE_ERR_INVALID, "null pointer");
    if (do_sub && (!frame_map || sub_max_t < 1 || sub_num < 0)) return fail(OE_ERR_INVALID, "bad spec_sub arguments");
    if (do_aug && ((n_t > 0 && (!tmask || max_t < 1)) || (n_f > 0 && (!fmask || max_f < 1)) || num_freq < 1 || n_t < 0 || n_f < 0))
        return fail(OE_ERR_INVALID, "bad spec_aug arguments");
    for (int i = 0; i < n; ++i)
        if (frames[i] < 1) return fail(OE_ERR_INVALID, "frames[%d] must be >= 1", i);
    PyMT g{mt_state, mt_state + 624};
    if (do_sub) {                                // feature_processor.py:55-64, all utterances first (dataset.py:204-205)
        std::vector<int32_t> tmp;
        int64_t off = 0;
        for (int i = 0; i < n; ++i) {
            const int T = frames[i];
            int32_t* idx = frame_map + off;
            for (int t = 0; t < T; ++t) idx[t] = t;
            for (int j = 0; j < sub_num; ++j) {
                const int start = g.randint(0, T - 1);
                const int length = g.randint(1, sub_max_t);
                const int end = std::min(T, start + length);
                const int pos = g.randint(0, start);
                tmp.assign(idx + start - pos, idx + end - pos);      // numpy reads the source before writing
                std::copy(tmp.begin(), tmp.end(), idx + start);
            }
            off += T;
        }
    }
    if (do_aug) {                                // feature_processor.py:31-41 (dataset.py:208-209)
        for (int i = 0; i < n; ++i) {
            const int T = frames[i];
            for (int j = 0; j < n_t; ++j) {
                const int start = g.randint(0, T - 1);
                const int length = g.randint(1, max_t);
                tmask[((int64_t)i * n_t + j) * 2] = start;
                tmask[((int64_t)i * n_t + j) * 2 + 1] = std::min(T, start + length);
            }
            for (int j = 0; j < n_f; ++j) {
                const int start = g.randint(0, num_freq - 1);
                const int length = g.randint(1, max_f);
                fmask[((int64_t)i * n_f + j) * 2] = start;
                fmask[((int64_t)i * n_f + j) * 2 + 1] = std::min(num_freq, start + length);
            }
        }
    }
    return OE_OK;
}

}  // extern "C"

// ==========================================================================================
// Native PCM ingest (host threads, no CUDA) -- oe_ingest.h
// ==========================================================================================
extern "C" {

int oe_ingest_create(int32_t threads, oe_ingest** out) {
    if (!out) return fail(OE_ERR_INVALID, "null out");
    oe_ingest* g = new oe_ingest();
    g->threads = threads > 0 ? threads : (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    g->pool = new oe_ing::Pool(g->threads);
    {
        const char* d = getenv("OE_INGEST_DIRECT");
        g->direct_read = d && d[0] == '1';
        const char* v = getenv("OE_FLAC_VERIFY_MD5");
        g->flac_verify_md5 = v && v[0] == '1';
    }
    *out = g;
    return OE_OK;
}

static void ingest_close_all(oe_ingest* g) {
    for (int& fd : g->fds)
        if (fd >= 0) {
            close(fd);
            fd = -1;
        }
}

int oe_ingest_destroy(oe_ingest* g) {
    if (!g) return OE_OK;
    if (g->driver.joinable()) {
        {
            std::lock_guard<std::mutex> lk(g->qm);
            g->quit = true;
        }
        g->qcv.notify_all();
        g->driver.join();
    }
    ingest_close_all(g);
    delete g->pool;
    delete g;
    return OE_OK;
}

const char* oe_ingest_error(const oe_ingest* g, int32_t index) {
    if (!g || index < 0 || index >= (int)g->errors.size()) return "";
    return g->errors[index].c_str();
}

int oe_ingest_probe(oe_ingest* g, int32_t n, const char* const* paths, const double* starts, const double* ends,
                    int32_t* n_samples, int32_t* sample_rates, int32_t* status) {
    if (!g || n < 0 || (n > 0 && (!paths || !n_samples || !sample_rates || !status))) return fail(OE_ERR_INVALID, "null pointer");
    ingest_close_all(g);
    g->errors.assign(n, std::string());
    g->fds.assign(n, -1);
    g->infos.assign(n, oe_ing::WavInfo());
    g->first.assign(n, 0);
    g->pool->run(n, [&](int i) {
        n_samples[i] = sample_rates[i] = 0;
        status[i] = OE_ERR_INVALID;
        const int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
        if (fd < 0) {
            g->errors[i] = std::string(paths[i]) + ": " + strerror(errno);
            return;
        }
        oe_ing::WavInfo& w = g->infos[i];
        std::string err = oe_ing::parse_wav(fd, paths[i], w);
        int64_t first = 0, count = 0;
        if (err.empty()) {
            const bool seg = starts && ends && !(starts[i] < 0.0);
            oe_ing::segment(w, seg ? starts[i] : 0.0, seg ? ends[i] : 0.0, seg, first, count);
            if (count > INT32_MAX) err = std::string(paths[i]) + ": more than 2^31 samples";
        }
        if (!err.empty()) {
            close(fd);
            g->errors[i] = err;
            status[i] = OE_ERR_UNSUPPORTED;
            return;
        }
        g->fds[i] = fd;
        g->first[i] = first;
        n_samples[i] = (int32_t)count;
        sample_rates[i] = w.sample_rate;
        status[i] = OE_OK;
    });
    return OE_OK;
}

int oe_ingest_read(oe_ingest* g, int32_t n, const char* const* paths, const double* starts, const double* ends,
                   int16_t* dst, const int64_t* offsets, const int32_t* n_samples, int32_t* status) {
    (void)starts;
    (void)ends;
    if (!g || n < 0 || (n > 0 && (!paths || !dst || !offsets || !n_samples || !status))) return fail(OE_ERR_INVALID, "null pointer");
    if ((int)g->fds.size() != n) return fail(OE_ERR_INVALID, "oe_ingest_read must follow oe_ingest_probe of the same %d entries", n);
    g->pool->run(n, [&](int i) {
        const int fd = g->fds[i];
        if (status[i] != OE_OK || fd < 0) return;                     // failed in the probe: message kept
        const oe_ing::WavInfo& w = g->infos[i];
        const int64_t first = g->first[i], count = n_samples[i];
        std::string err;
        int16_t* const out = dst + offsets[i];
        if (w.flac) {
            // the whole compressed file (a LibriSpeech utterance: ~0.3 MB) -> channel 0 of the segment as int32 -> int16
            // through the bounce buffer into the packed buffer, all on this reader thread
            static thread_local std::vector<unsigned char> file;
            static thread_local std::vector<int32_t> pcm;
            static thread_local std::vector<char> bounce(oe_ing::kBounceBytes);
            struct stat st;
            if (fstat(fd, &st) != 0) err = std::string(paths[i]) + ": " + strerror(errno);
            if (err.empty()) {
                file.resize((size_t)st.st_size);
                int64_t done = 0;
                while (done < st.st_size) {
                    const ssize_t r = pread(fd, file.data() + done, (size_t)(st.st_size - done), done);
                    if (r <= 0) break;
                    done += r;
                }
                if (done != st.st_size) err = std::string(paths[i]) + ": short read";
            }
            if (err.empty()) {
                pcm.resize((size_t)std::max<int64_t>(count, 1));
                oe_flac::Info fi;
                int64_t total = 0;
                const std::string e = oe_flac::decode(file.data(), (int64_t)file.size(), 0, first, count, pcm.data(), g->flac_verify_md5, fi, &total);
                if (!e.empty()) err = std::string(paths[i]) + ": " + e;
                else if (total < first + count) err = std::string(paths[i]) + ": the stream is shorter than its STREAMINFO block announces";
            }
            if (err.empty()) {
                char* const o8 = reinterpret_cast<char*>(out);
                int16_t* const b16 = reinterpret_cast<int16_t*>(bounce.data());
                const int64_t chunk = oe_ing::kBounceBytes / 2;
                for (int64_t s0 = 0; s0 < count; s0 += chunk) {
                    const int64_t m = std::min<int64_t>(chunk, count - s0);
                    for (int64_t k = 0; k < m; ++k) b16[k] = (int16_t)pcm[(size_t)(s0 + k)];
                    oe_ing::stream_copy(o8 + 2 * s0, bounce.data(), (size_t)(2 * m));
                }
            }
        } else if (w.channels == 1 && g->direct_read) {               // straight into the packed (pinned) buffer
            int64_t done = 0;
            const int64_t bytes = 2 * count;
            while (done < bytes) {
                const ssize_t r = pread(fd, reinterpret_cast<char*>(out) + done, (size_t)(bytes - done), w.data_off + 2 * first + done);
                if (r <= 0) break;
                done += r;
            }
            if (done != bytes) err = std::string(paths[i]) + ": short read";
        } else if (w.channels == 1) {
            // through a cache-resident bounce buffer, then streaming (non-temporal) stores into the packed buffer: a pread
            // straight into the destination makes every destination line a read-for-ownership first -- three DRAM transfers
            // per byte (page cache read, destination read, destination write) where two suffice, on a host whose memory the
            // H2D copy engines are reading at full rate at the same time
            static thread_local std::vector<char> bounce(oe_ing::kBounceBytes);
            int64_t done = 0;
            const int64_t bytes = 2 * count;
            char* const o8 = reinterpret_cast<char*>(out);                // 16-byte aligned: offsets are multiples of 8 samples
            while (done < bytes) {
                const int64_t want = std::min<int64_t>(oe_ing::kBounceBytes, bytes - done);
                const ssize_t r = pread(fd, bounce.data(), (size_t)want, w.data_off + 2 * first + done);
                if (r <= 0) break;
                oe_ing::stream_copy(o8 + done, bounce.data(), (size_t)r);
                done += r;
            }
            if (done != bytes) err = std::string(paths[i]) + ": short read";
        } else {                                                      // channel 0 of interleaved frames (torchaudio.load(...)[0])
            std::vector<int16_t> tmp((size_t)65536 * w.channels);
            for (int64_t f0 = 0; f0 < count; f0 += 65536) {
                const int64_t nf = std::min<int64_t>(65536, count - f0);
                const int64_t bytes = nf * 2 * w.channels;
                if (pread(fd, tmp.data(), (size_t)bytes, w.data_off + (first + f0) * 2 * w.channels) != bytes) {
                    err = std::string(paths[i]) + ": short read";
                    break;
                }
                for (int64_t f = 0; f < nf; ++f) out[f0 + f] = tmp[(size_t)f * w.channels];
            }
        }
        if (!err.empty()) {
            g->errors[i] = err;
            status[i] = OE_ERR_INVALID;
        }
    });
    ingest_close_all(g);
    return OE_OK;
}

// ---- FLAC streams in memory (shard-tar members, other sample sizes: the Python mirror scales them like torchaudio) ----
int oe_flac_info(const void* data, int64_t size, int32_t* sample_rate, int32_t* channels, int32_t* bits, int64_t* total_samples) {
    if (!data || size <= 0) return fail(OE_ERR_INVALID, "null / empty FLAC buffer");
    oe_flac::Info fi;
    std::string err = oe_flac::parse_streaminfo(static_cast<const unsigned char*>(data), size, fi);
    if (err.empty() && fi.total == 0) {
        int64_t n = 0;
        err = oe_flac::decode(static_cast<const unsigned char*>(data), size, 0, 0, 0, nullptr, false, fi, &n);
        fi.total = n;
    }
    if (!err.empty()) return fail(OE_ERR_UNSUPPORTED, "FLAC: %s", err.c_str());
    if (sample_rate) *sample_rate = fi.sample_rate;
    if (channels) *channels = fi.channels;
    if (bits) *bits = fi.bits;
    if (total_samples) *total_samples = fi.total;
    return OE_OK;
}

int oe_flac_decode(const void* data, int64_t size, int32_t channel, int64_t first, int64_t count, int32_t* out,
                   int32_t verify_md5, int64_t* decoded) {
    if (!data || size <= 0 || first < 0 || count < 0 || (count > 0 && !out)) return fail(OE_ERR_INVALID, "bad FLAC decode arguments");
    oe_flac::Info fi;
    int64_t n = 0;
    const std::string err = oe_flac::decode(static_cast<const unsigned char*>(data), size, channel, first, count, out, verify_md5 != 0, fi, &n);
    if (!err.empty()) return fail(OE_ERR_UNSUPPORTED, "FLAC: %s", err.c_str());
    if (first + count > n) return fail(OE_ERR_INVALID, "FLAC: samples [%lld, %lld) requested, the stream holds %lld", (long long)first, (long long)(first + count), (long long)n);
    if (decoded) *decoded = n;
    return OE_OK;
}

// ---- FLAC decoded on the GPU: host-side packing (file bytes + frame index), the launch, and the fixture encoder ----
int oe_flac_pack(oe_ingest* g, int32_t n, const char* const* paths, const double* starts, const double* ends,
                 void* comp, int64_t comp_capacity, oe_flac_frame* frames, int64_t frames_capacity,
                 int64_t* comp_offsets, int64_t* pcm_offsets, int32_t* n_samples, int32_t* sample_rates, int32_t* status,
                 int64_t* comp_bytes, int64_t* n_frames, int64_t* total_samples) {
    if (!g || n < 0 || (n > 0 && (!paths || !comp_offsets || !pcm_offsets || !n_samples || !sample_rates || !status)) || !comp_bytes ||
        !n_frames || !total_samples)
        return fail(OE_ERR_INVALID, "null pointer");
    ingest_close_all(g);
    g->errors.assign(n, std::string());
    g->fds.assign(n, -1);
    g->first.assign(n, 0);
    std::vector<int64_t> fsize((size_t)n, 0), fbound((size_t)n, 0);
    std::vector<oe_flac::Info> infos((size_t)n);
    // pass 1: open, STREAMINFO, segment
    g->pool->run(n, [&](int i) {
        n_samples[i] = sample_rates[i] = 0;
        status[i] = OE_ERR_INVALID;
        const int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
        if (fd < 0) {
            g->errors[i] = std::string(paths[i]) + ": " + strerror(errno);
            return;
        }
        std::string err;
        struct stat st;
        unsigned char head[42];
        oe_flac::Info& fi = infos[i];
        if (fstat(fd, &st) != 0) err = strerror(errno);
        else if (pread(fd, head, 42, 0) != 42) err = "not a FLAC stream";
        else if (const int64_t id3 = oe_flac::id3v2_bytes(head, 42); id3 > 0 && pread(fd, head, 42, id3) != 42) err = "not a FLAC stream";
        else if (memcmp(head, "fLaC", 4) != 0) err = "not a FLAC stream";
        else {
            head[4] |= 0x80;
            err = oe_flac::parse_streaminfo(head, 42, fi);
        }
        int code = OE_ERR_INVALID;
        if (err.empty()) {
            code = OE_ERR_UNSUPPORTED;
            if (fi.channels != 1) err = std::to_string(fi.channels) + " channels (the GPU decoder takes mono streams; the host decoder reads channel 0)";
            else if (fi.bits != 16) err = std::to_string(fi.bits) + "-bit FLAC (the GPU decoder fills an int16 buffer with 16-bit streams; read_wav scales the others like torchaudio)";
            else if (fi.total == 0) err = "the stream does not announce its length (host decoder)";
        }
        int64_t first = 0, count = 0;
        if (err.empty()) {
            oe_ing::WavInfo w;
            w.sample_rate = fi.sample_rate, w.channels = 1, w.bits = fi.bits, w.frames = fi.total;
            const bool seg = starts && ends && !(starts[i] < 0.0);
            oe_ing::segment(w, seg ? starts[i] : 0.0, seg ? ends[i] : 0.0, seg, first, count);
            if (count > INT32_MAX) err = "more than 2^31 samples";
        }
        if (!err.empty()) {
            close(fd);
            g->errors[i] = std::string(paths[i]) + ": " + err;
            status[i] = code;
            return;
        }
        g->fds[i] = fd;
        g->first[i] = first;
        fsize[i] = st.st_size;
        fbound[i] = fi.total / std::max(fi.min_block, 1) + 2;        // capacity planning only: checked again after the walk
        n_samples[i] = (int32_t)count;
        sample_rates[i] = fi.sample_rate;
        status[i] = OE_OK;
    });
    int64_t cb = 0, fb = 0, total = 0;
    for (int i = 0; i < n; ++i) {
        comp_offsets[i] = cb;
        cb += (fsize[i] + 15) / 16 * 16;
        fb += fbound[i];
        pcm_offsets[i] = total;
        total += ((int64_t)n_samples[i] + 7) / 8 * 8;
    }
    *comp_bytes = cb;
    *total_samples = total;
    if (cb + 16 > comp_capacity || fb > frames_capacity || !comp || !frames) {
        *n_frames = fb;
        ingest_close_all(g);
        return fail(OE_ERR_WORKSPACE, "compressed buffer / frame table too small: %lld + 16 bytes, %lld frames needed", (long long)cb, (long long)fb);
    }
    // pass 2: file bytes into the (pinned) buffer, frame index per file
    std::vector<std::vector<oe_flac_frame>> per((size_t)n);
    unsigned char* const cbase = static_cast<unsigned char*>(comp);
    g->pool->run(n, [&](int i) {
        const int fd = g->fds[i];
        if (status[i] != OE_OK || fd < 0) return;
        unsigned char* const d = cbase + comp_offsets[i];
        int64_t done = 0;
        while (done < fsize[i]) {
            const ssize_t r = pread(fd, d + done, (size_t)(fsize[i] - done), done);
            if (r <= 0) break;
            done += r;
        }
        close(fd);                                                     // by the thread that read it: 256 serial close() calls cost 0.1 ms
        g->fds[i] = -1;
        std::string err;
        if (done != fsize[i]) err = "short read";
        oe_flac::Info fi;
        if (err.empty()) err = oe_flac::parse_streaminfo(d, fsize[i], fi);
        std::vector<oe_flac::FrameSpan> spans;
        if (err.empty()) err = oe_flac::scan_frames(d, fsize[i], fi, spans);
        const int64_t first = g->first[i], last = first + n_samples[i];
        int64_t covered = first;
        if (err.empty()) {
            for (const oe_flac::FrameSpan& s : spans) {
                const int64_t a = std::max(first, s.first_sample), b = std::min(last, s.first_sample + s.h.block);
                if (a >= b) continue;
                if (s.h.ch_code != 0 || s.h.bps != fi.bits) {
                    err = "channel layout / sample size changes inside the stream";
                    break;
                }
                if (a != covered) break;
                oe_flac_frame f;
                f.comp_off = comp_offsets[i] + s.off;
                f.out_off = pcm_offsets[i] + (a - first);
                f.frame_bytes = (int32_t)s.bytes;
                f.hdr_bytes = s.h.hdr_bytes;
                f.block = s.h.block;
                f.bps = s.h.bps;
                f.skip = (int32_t)(a - s.first_sample);
                f.take = (int32_t)(b - a);
                f.utt = i;
                f.reserved = 0;
                per[i].push_back(f);
                covered = b;
            }
            if (err.empty() && covered != last) err = "the frames hold fewer samples than STREAMINFO announces";
            if (err.empty() && (int64_t)per[i].size() > fbound[i]) err = "more frames than STREAMINFO's minimum block size allows";
        }
        if (!err.empty()) {
            g->errors[i] = std::string(paths[i]) + ": " + err;
            status[i] = OE_ERR_INVALID;
            n_samples[i] = 0;                                          // its slot in the PCM layout stays (unused)
            per[i].clear();
        }
    });
    ingest_close_all(g);
    int64_t nf = 0;
    for (int i = 0; i < n; ++i) {
        if (!per[i].empty()) memcpy(frames + nf, per[i].data(), per[i].size() * sizeof(oe_flac_frame));
        nf += (int64_t)per[i].size();
    }
    memset(cbase + cb, 0, 16);
    *n_frames = nf;
    return OE_OK;
}

int oe_flac_decode_batch(const void* d_comp, int64_t comp_bytes, const oe_flac_frame* d_frames, int64_t n_frames,
                         int16_t* d_pcm, int32_t* d_errors, int32_t verify_crc, oe_stream stream) {
    if (n_frames < 0 || comp_bytes < 0) return fail(OE_ERR_INVALID, "negative size");
    if (n_frames == 0) return OE_OK;
    if (!d_comp || !d_frames || !d_pcm || !d_errors) return fail(OE_ERR_INVALID, "null pointer");
    if ((reinterpret_cast<uintptr_t>(d_comp) & 15) || (comp_bytes & 15)) return fail(OE_ERR_INVALID, "d_comp and comp_bytes must be multiples of 16 (oe_flac_pack's layout)");
    const int64_t blocks = (n_frames + 31) / 32;
    if (blocks > INT32_MAX) return fail(OE_ERR_INVALID, "too many frames");
    oe_flacgpu::oe_flac_decode_kernel<<<(unsigned)blocks, 96, 0, (cudaStream_t)stream>>>(
        static_cast<const unsigned char*>(d_comp), comp_bytes + 16, d_frames, n_frames, d_pcm, d_errors, verify_crc);
    OE_CUDA(cudaGetLastError());
    return OE_OK;
}

int oe_flac_encode(const int16_t* pcm, int64_t n, int32_t sample_rate, int32_t block, int32_t partition_order, void* out,
                   int64_t capacity, int64_t* bytes) {
    if (n < 0 || (n > 0 && !pcm) || !bytes || sample_rate <= 0 || sample_rate >= (1 << 20) || block < 16 || block > 65535 ||
        partition_order < 0 || partition_order > 8)
        return fail(OE_ERR_INVALID, "bad FLAC encode arguments");
    static thread_local std::vector<unsigned char> buf;
    oe_flac::encode(pcm, n, sample_rate, block, partition_order, buf);
    *bytes = (int64_t)buf.size();
    if (!out || capacity < (int64_t)buf.size()) return fail(OE_ERR_WORKSPACE, "FLAC stream needs %lld bytes", (long long)buf.size());
    memcpy(out, buf.data(), buf.size());
    return OE_OK;
}

// ---- asynchronous batches ----
static void ingest_driver(oe_ingest* g) {
    for (;;) {
        oe_ingest_job* job = nullptr;
        {
            std::unique_lock<std::mutex> lk(g->qm);
            g->qcv.wait(lk, [&] { return g->quit || !g->queue.empty(); });
            if (g->queue.empty()) return;                              // quit and nothing left
            job = g->queue.front();
            g->queue.erase(g->queue.begin());
        }
        const int n = (int)job->paths.size();
        if (job->flac) {
            job->rc = oe_flac_pack(g, n, job->cpaths.data(), job->starts.data(), job->ends.data(), job->comp, job->comp_capacity,
                                   job->frames, job->frames_capacity, job->comp_offsets.data(), job->offsets.data(), job->lens.data(),
                                   job->rates.data(), job->status.data(), &job->comp_bytes, &job->n_frames, &job->total);
            if (job->rc != OE_OK) job->message = oe_last_error();
            job->errors = g->errors;
            {
                std::lock_guard<std::mutex> lk(g->qm);
                job->done = true;
            }
            g->dcv.notify_all();
            continue;
        }
        oe_ingest_probe(g, n, job->cpaths.data(), job->starts.data(), job->ends.data(), job->lens.data(), job->rates.data(),
                        job->status.data());
        int64_t total = 0;
        for (int i = 0; i < n; ++i) {                                  // 8-sample aligned packing (include/openeat_frontend.h: wav_offsets)
            job->offsets[i] = total;
            total += ((int64_t)job->lens[i] + 7) / 8 * 8;
        }
        job->total = total;
        if (total > job->capacity) {
            for (int i = 0; i < n; ++i)
                if (job->status[i] == OE_OK) {
                    job->status[i] = OE_ERR_WORKSPACE;
                    g->errors[i] = "destination buffer too small";
                }
            ingest_close_all(g);
        } else {
            oe_ingest_read(g, n, job->cpaths.data(), job->starts.data(), job->ends.data(), job->dst, job->offsets.data(),
                           job->lens.data(), job->status.data());
        }
        job->errors = g->errors;
        {
            std::lock_guard<std::mutex> lk(g->qm);
            job->done = true;
        }
        g->dcv.notify_all();
    }
}

int oe_ingest_submit(oe_ingest* g, int32_t n, const char* const* paths, const double* starts, const double* ends,
                     int16_t* dst, int64_t dst_capacity, oe_ingest_job** out) {
    if (!g || !out || n < 0 || (n > 0 && (!paths || !dst))) return fail(OE_ERR_INVALID, "null pointer");
    oe_ingest_job* job = new oe_ingest_job();
    job->owner = g;
    job->paths.assign(paths, paths + n);
    for (auto& p : job->paths) job->cpaths.push_back(p.c_str());
    job->starts.assign(n, -1.0);
    job->ends.assign(n, 0.0);
    if (starts && ends) {
        job->starts.assign(starts, starts + n);
        job->ends.assign(ends, ends + n);
    }
    job->dst = dst;
    job->capacity = dst_capacity;
    job->offsets.assign(n, 0);
    job->lens.assign(n, 0);
    job->rates.assign(n, 0);
    job->status.assign(n, OE_ERR_INVALID);
    {
        std::lock_guard<std::mutex> lk(g->qm);
        if (!g->driver.joinable()) g->driver = std::thread(ingest_driver, g);
        g->queue.push_back(job);
    }
    g->qcv.notify_all();
    *out = job;
    return OE_OK;
}

int oe_ingest_wait(oe_ingest_job* job, const int64_t** offsets, const int32_t** n_samples, const int32_t** sample_rates,
                   const int32_t** status, int64_t* total) {
    if (!job) return fail(OE_ERR_INVALID, "null job");
    oe_ingest* g = job->owner;
    {
        std::unique_lock<std::mutex> lk(g->qm);
        g->dcv.wait(lk, [&] { return job->done; });
    }
    if (offsets) *offsets = job->offsets.data();
    if (n_samples) *n_samples = job->lens.data();
    if (sample_rates) *sample_rates = job->rates.data();
    if (status) *status = job->status.data();
    if (total) *total = job->total;
    return OE_OK;
}

int oe_flac_submit(oe_ingest* g, int32_t n, const char* const* paths, const double* starts, const double* ends, void* comp,
                   int64_t comp_capacity, oe_flac_frame* frames, int64_t frames_capacity, oe_ingest_job** out) {
    if (!g || !out || n < 0 || (n > 0 && !paths)) return fail(OE_ERR_INVALID, "null pointer");
    oe_ingest_job* job = new oe_ingest_job();
    job->owner = g;
    job->flac = true;
    job->paths.assign(paths, paths + n);
    for (auto& p : job->paths) job->cpaths.push_back(p.c_str());
    job->starts.assign(n, -1.0);
    job->ends.assign(n, 0.0);
    if (starts && ends) {
        job->starts.assign(starts, starts + n);
        job->ends.assign(ends, ends + n);
    }
    job->comp = comp;
    job->comp_capacity = comp_capacity;
    job->frames = frames;
    job->frames_capacity = frames_capacity;
    job->comp_offsets.assign(n, 0);
    job->offsets.assign(n, 0);
    job->lens.assign(n, 0);
    job->rates.assign(n, 0);
    job->status.assign(n, OE_ERR_INVALID);
    {
        std::lock_guard<std::mutex> lk(g->qm);
        if (!g->driver.joinable()) g->driver = std::thread(ingest_driver, g);
        g->queue.push_back(job);
    }
    g->qcv.notify_all();
    *out = job;
    return OE_OK;
}

int oe_flac_wait(oe_ingest_job* job, const int64_t** pcm_offsets, const int32_t** n_samples, const int32_t** sample_rates,
                 const int32_t** status, int64_t* comp_bytes, int64_t* n_frames, int64_t* total_samples) {
    if (!job || !job->flac) return fail(OE_ERR_INVALID, "not a job of oe_flac_submit");
    oe_ingest* g = job->owner;
    {
        std::unique_lock<std::mutex> lk(g->qm);
        g->dcv.wait(lk, [&] { return job->done; });
    }
    if (pcm_offsets) *pcm_offsets = job->offsets.data();
    if (n_samples) *n_samples = job->lens.data();
    if (sample_rates) *sample_rates = job->rates.data();
    if (status) *status = job->status.data();
    if (comp_bytes) *comp_bytes = job->comp_bytes;
    if (n_frames) *n_frames = job->n_frames;
    if (total_samples) *total_samples = job->total;
    if (job->rc != OE_OK) return fail(job->rc, "%s", job->message.c_str());
    return OE_OK;
}

const char* oe_ingest_job_error(const oe_ingest_job* job, int32_t index) {
    if (!job || index < 0 || index >= (int)job->errors.size()) return "";
    return job->errors[index].c_str();
}

int oe_ingest_job_release(oe_ingest_job* job) {
    delete job;
    return OE_OK;
}

int oe_host_pad_rows(oe_ingest* g, const float* src, const int32_t* frames, int32_t B, int32_t tmax, int32_t F,
                     const float* pad_row, float* dst) {
    if (!g || B < 0 || tmax < 0 || F <= 0 || (B > 0 && tmax > 0 && (!src || !frames || !dst))) return fail(OE_ERR_INVALID, "null pointer / bad shape");
    std::vector<int64_t> first((size_t)B + 1, 0);
    for (int b = 0; b < B; ++b) {
        if (frames[b] < 0 || frames[b] > tmax) return fail(OE_ERR_INVALID, "frames[%d]=%d outside [0, %d]", b, frames[b], tmax);
        first[b + 1] = first[b] + frames[b];
    }
    // work items: (utterance, chunk of kChunk padded rows): real rows are copied, rows behind them filled
    constexpr int kChunk = 128;
    const int per = (tmax + kChunk - 1) / kChunk;
    const size_t row = (size_t)F * sizeof(float);
    std::vector<float> zero;
    if (!pad_row) {
        zero.assign((size_t)F, 0.f);
        pad_row = zero.data();
    }
    g->pool->run(B * per, [&](int item) {
        const int b = item / per, t0 = (item - b * per) * kChunk, t1 = std::min(tmax, t0 + kChunk);
        const int nf = frames[b];
        char* const d = reinterpret_cast<char*>(dst + ((size_t)b * tmax + t0) * F);
        const int real = std::max(0, std::min(nf, t1) - t0);
        if (real > 0) oe_ing::stream_copy(d, reinterpret_cast<const char*>(src + (first[b] + t0) * F), (size_t)real * row);
        for (int t = t0 + real; t < t1; ++t) oe_ing::stream_copy(d + (size_t)(t - t0) * row, reinterpret_cast<const char*>(pad_row), row);
    });
    return OE_OK;
}

}  // extern "C"
