// tcgen05 feasibility checks for the third-generation fbank kernel:
//   (1) does a K-major, un-swizzled shared-memory descriptor with LBO = 16 B and SBO = 128 B -- i.e. row m, K chunk c at
//       byte 16 (m + c): consecutive rows ALIAS each other's K chunks, which is exactly the overlap of speech frames
//       (frame m + 1 = frame m shifted by one 160-sample block) -- produce the right GEMM?
//   (2) how many cycles does one tcgen05.mma M 128 x N x K 16 (kind::f16, fp16 operands from shared memory) cost in a
//       long back-to-back sequence, for N = 16 / 32 / 64 / 128, aliased vs. canonical operand layout?
//   (3) how fast are narrow, strided tcgen05.ld (32x32b.x2 every 32 columns) against one wide 32x32b.x32?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_alias mma_alias.cu && ./mma_alias
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// D fp32, A / B fp16 (format 0), both K-major
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

constexpr int kRows = 136;                 // plane rows (128 + K chunks of slack)
constexpr int kPlaneBytes = kRows * 16;    // one aliased plane: row q = 8 fp16
constexpr int kPlanes = 32;
constexpr int kBBytesMax = 128 * 32 * 2;   // N <= 128 rows x K 32
constexpr int kSmem = kPlanes * kPlaneBytes + 128 * 32 * 2 /* canonical A, K = 32 */ + kBBytesMax + 64;

// mode 0: correctness of the aliased layout (one MMA pair, K = 32 = chunks 0..3 of plane 0), result to d_out [128][n]
// mode 1: timing, aliased A (plane p, two K steps), `reps` x 32 planes x 2 MMAs
// mode 2: timing, canonical A layout (core matrices contiguous, LBO 128, SBO 256), same count
// mode 3: tcgen05.ld timing
__global__ void __launch_bounds__(128, 1) k(const __half* __restrict__ planes, const __half* __restrict__ bmat, float* __restrict__ d_out,
                                             long long* __restrict__ cycles, int n, int mode, int reps, int dep) {
    extern __shared__ __align__(128) unsigned char smem[];
    __half* const sP = reinterpret_cast<__half*>(smem);
    __half* const sA = reinterpret_cast<__half*>(smem + kPlanes * kPlaneBytes);
    __half* const sB = reinterpret_cast<__half*>(smem + kPlanes * kPlaneBytes + 128 * 32 * 2);
    uint64_t* const bar = reinterpret_cast<uint64_t*>(smem + kPlanes * kPlaneBytes + 128 * 32 * 2 + kBBytesMax);
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kPlanes * kRows * 8; i += 128) sP[i] = planes[i % (kRows * 8)];
    for (int i = tid; i < 128 * 32; i += 128) sA[i] = __float2half(0.25f);
    for (int i = tid; i < n * 32; i += 128) sB[i] = bmat[i];       // canonical K-major layout for K = 32: [n/8][k/8][8][8]
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = make_idesc(128, n);
    const uint32_t p0 = smem_u32(sP), a0 = smem_u32(sA), b0 = smem_u32(sB);
    // B: canonical, K = 32: core matrix (n8, k8) at ((n8 * 4) + k8) * 128 B -> LBO 128, SBO 512; K step 16 = +256 B
    long long t0 = 0, t1 = 0;
    if (mode <= 2) {
        if (tid == 0) {
            t0 = clock64();
            // the issue loop must not cost more than the MMAs: descriptors are base + small increments, no division
            const uint64_t da_alias = make_desc(p0, 16, 128), da_canon = make_desc(a0, 128, 512), db0 = make_desc(b0, 128, 512);
            const uint32_t nmask = (uint32_t)(512 / n) - 1u;
            if (mode == 0) {
                mma_f16(tmem, da_alias, db0, idesc, 0u);
                mma_f16(tmem, da_alias + 2, db0 + 16, idesc, 1u);              // K step 2: chunks 2, 3 (+32 B); B +256 B
            } else {
                for (int r = 0; r < reps; ++r) {
#pragma unroll 8
                    for (int p = 0; p < kPlanes; ++p) {
                        const uint64_t da = mode == 2 ? da_canon : da_alias + (uint64_t)(p * (kPlaneBytes >> 4));
                        const uint32_t c0 = dep ? ((uint32_t)p & nmask) * n : ((uint32_t)(2 * p) & nmask) * n;
                        const uint32_t c1 = dep ? c0 : ((uint32_t)(2 * p + 1) & nmask) * n;
                        mma_f16(tmem + c0, da, db0, idesc, 1u);
                        mma_f16(tmem + c1, mode == 2 ? da + 16 : da + 2, db0 + 16, idesc, 1u);
                    }
                }
            }
            commit(smem_u32(bar));
        }
        mbar_wait(smem_u32(bar), 0);
        t1 = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (mode == 0) {
            for (int c0 = 0; c0 < n; c0 += 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem + c0 + ((uint32_t)(32 * warp) << 16);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                               "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(taddr)
                             : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int c = 0; c < 16; ++c) d_out[tid * n + c0 + c] = __uint_as_float(r[c]);
            }
        }
        if (tid == 0 && cycles != nullptr) cycles[0] = t1 - t0;
    } else {
        // every warp reads its 32 lanes: (a) 16 x (x2 at stride 32 columns), (b) one x32; `reps` times each
        float keep = 0.f;
        __syncthreads();
        t0 = clock64();
        for (int i = 0; i < reps; ++i) {
            uint32_t r[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t taddr = tmem + 32 * j + 2 * (i & 15) + ((uint32_t)(32 * warp) << 16);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[2 * j]), "=r"(r[2 * j + 1]) : "r"(taddr) : "memory");
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) keep += __uint_as_float(r[j]);
        }
        __syncthreads();
        t1 = clock64();
        if (tid == 0) cycles[0] = t1 - t0;
        __syncthreads();
        t0 = clock64();
        for (int i = 0; i < reps; ++i) {
            uint32_t r[32];
            const uint32_t taddr = tmem + 32 * (i & 15) + ((uint32_t)(32 * warp) << 16);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
                "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) keep += __uint_as_float(r[j]);
        }
        __syncthreads();
        t1 = clock64();
        if (tid == 0) cycles[1] = t1 - t0;
        if (keep == 123.456f) d_out[0] = keep;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    // plane: row q holds 8 values; A[m][8 c + e] = plane[m + c][e]
    std::vector<__half> plane(kRows * 8);
    std::vector<float> pf(kRows * 8);
    srand(3);
    for (int i = 0; i < kRows * 8; ++i) {
        pf[i] = (float)((rand() % 4097) - 2048);
        plane[i] = __float2half(pf[i]);
    }
    __half* dP;
    float* dD;
    long long* dC;
    CK(cudaMalloc(&dP, plane.size() * 2));
    CK(cudaMemcpy(dP, plane.data(), plane.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dD, 128 * 128 * 4));
    CK(cudaMalloc(&dC, 16));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    for (int n : {16, 32, 64, 128}) {
        std::vector<float> bf(n * 32);
        std::vector<__half> bm(n * 32);
        for (int j = 0; j < n; ++j)
            for (int kk = 0; kk < 32; ++kk) {
                const float v = (float)((rand() % 2001) - 1000) / 1024.0f;
                bf[j * 32 + kk] = __half2float(__float2half(v));
                bm[((j >> 3) * 4 + (kk >> 3)) * 64 + (j & 7) * 8 + (kk & 7)] = __float2half(v);
            }
        __half* dB;
        CK(cudaMalloc(&dB, bm.size() * 2));
        CK(cudaMemcpy(dB, bm.data(), bm.size() * 2, cudaMemcpyHostToDevice));
        k<<<1, 128, kSmem>>>(dP, dB, dD, dC, n, 0, 1, 1);
        CK(cudaDeviceSynchronize());
        std::vector<float> D(128 * n);
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0, scale = 0;
        for (int m = 0; m < 128; ++m)
            for (int j = 0; j < n; ++j) {
                double ref = 0;
                for (int kk = 0; kk < 32; ++kk) ref += (double)pf[(m + (kk >> 3)) * 8 + (kk & 7)] * bf[j * 32 + kk];
                worst = fmax(worst, fabs(ref - D[m * n + j]));
                scale = fmax(scale, fabs(ref));
            }
        printf("N %3d aliased-layout GEMM check: max |err| %.3e on values up to %.3e\n", n, worst, scale);
        for (int md = 2; md <= 5; ++md) {
            const int mode = md >> 1, dep = md & 1;
            const int reps = 16;
            k<<<1, 128, kSmem>>>(dP, dB, dD, dC, n, mode, reps, dep);
            CK(cudaDeviceSynchronize());
            long long c;
            CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
            printf("N %3d %s A, %s accumulators: %lld cycles for %d MMAs (M 128, K 16) = %.1f cycles per MMA (one CTA alone)\n", n,
                   mode == 1 ? "aliased  " : "canonical", dep ? "pairwise dependent" : "independent", c, reps * kPlanes * 2, (double)c / (reps * kPlanes * 2));
            // all SMs busy: same kernel on 148 CTAs, wall time
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0));
            CK(cudaEventCreate(&e1));
            k<<<148, 128, kSmem>>>(dP, dB, dD, nullptr, n, mode, 64, dep);
            CK(cudaEventRecord(e0));
            k<<<148, 128, kSmem>>>(dP, dB, dD, nullptr, n, mode, 64, dep);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("      148 CTAs x %d MMAs: %.1f us -> %.1f ns per MMA per SM\n", 64 * kPlanes * 2, ms * 1e3, ms * 1e6 / (64 * kPlanes * 2));
        }
        CK(cudaFree(dB));
    }
    {
        std::vector<__half> bm(16 * 32);
        __half* dB;
        CK(cudaMalloc(&dB, bm.size() * 2));
        CK(cudaMemset(dB, 0, bm.size() * 2));
        k<<<1, 128, kSmem>>>(dP, dB, dD, dC, 16, 3, 256, 0);
        CK(cudaDeviceSynchronize());
        long long c[2];
        CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
        printf("tcgen05.ld, 4 warps, 32 registers per thread per rep: 16 x (32x32b.x2, stride 32 columns) %.1f cycles, 1 x 32x32b.x32 %.1f cycles\n",
               c[0] / 256.0, c[1] / 256.0);
    }
    return 0;
}
