// Tensor-core side of the "DFT as GEMM" question: how long does the tcgen05 pipe need for the two
// 16-point DFT stages of 150 k frames when they are fed as split-bf16 GEMMs -- with everything else (operand staging,
// twiddles, fp32 -> 3 x bf16 splitting, untangle, power, mel) left out?  A LOWER bound for that formulation.
//
// Per group of 8 frames and per stage: D[128 x 32] (fp32, TMEM) = sum over the six split products (a1 b1, a1 b2, a2 b1,
// a1 b3, a2 b2, a3 b1: what fp32-grade accuracy needs, tools/ubench/dft_tc_accuracy.py) of A_i[128 x 32] . B_j[32 x 32]:
// rows = (frame, n2) resp. (frame, k1), K = (re | im) x 16 points, N = (re | im) x 16 outputs.  12 tcgen05.mma
// (kind::f16, bf16 operands from shared memory, M 128, N 32, K 16) per stage, issued by one thread, committed to an
// mbarrier; then the four warps read the accumulator back (tcgen05.ld 32x32b.x32), as any epilogue must -- double
// buffered, so the MMAs of the next step run under the read-back of the current one.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dft_tc dft_tc.cu && ./dft_tc
//
// Prints the GEMM check of the first tile against the host (so the descriptors are known to be right) and the time for
// 150 368 frames on 148 persistent CTAs.
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

constexpr int kM = 128, kN = 32, kK = 32;          // one stage of 8 frames
constexpr int kABytes = kM * kK * 2;               // 8 KB per split piece, K-major core-matrix layout
constexpr int kBBytes = kN * kK * 2;               // 2 KB
constexpr int kSmem = 3 * kABytes + 3 * kBBytes + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 bytes, contiguous (128 B);
// LBO = distance between core matrices along K, SBO = distance between 8-row groups along M / N (16-byte units)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    return d;                                      // base offset 0, swizzle NONE
}

// instruction descriptor: D fp32, A / B bf16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
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

// element (row, k) of a K-major operand with `kk` columns: core matrices [row / 8][k / 8], 8 rows x 8 bf16 each
__host__ __device__ inline int op_index(int row, int k, int kk) { return ((row >> 3) * (kk >> 3) + (k >> 3)) * 64 + (row & 7) * 8 + (k & 7); }

__global__ void __launch_bounds__(128, 1) dft_tc_kernel(const uint16_t* __restrict__ a_split, const uint16_t* __restrict__ b_split,
                                                         float* __restrict__ d_out, int groups_per_cta, int n_stages) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint16_t* const sA = reinterpret_cast<uint16_t*>(smem);
    uint16_t* const sB = reinterpret_cast<uint16_t*>(smem + 3 * kABytes);
    uint64_t* const bar = reinterpret_cast<uint64_t*>(smem + 3 * kABytes + 3 * kBBytes);
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + 3 * kABytes + 3 * kBBytes + 32);
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < 3 * kM * kK; i += 128) sA[i] = a_split[i];
    for (int i = tid; i < 3 * kN * kK; i += 128) sB[i] = b_split[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // operands written with st.shared -> read by the tensor pipe
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = make_idesc(kM, kN);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    constexpr int kProd[6][2] = {{0, 0}, {0, 1}, {1, 0}, {0, 2}, {1, 1}, {2, 0}};
    float keep = 0.f;
    // two accumulators / two mbarriers: the MMAs of step i + 1 are issued before step i is waited for and read back, so
    // the tensor pipe never idles behind the read-back (what a real kernel's MMA warp / epilogue warps split would do)
    auto issue = [&](int i) {
        const uint32_t acc = tmem + 32 * (i & 1);
#pragma unroll
        for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int ks = 0; ks < kK / 16; ++ks) {
                // K step of 16 bf16 = two core matrices along K: +256 bytes; LBO 128 B, SBO (K / 8) x 128 B
                const uint64_t da = make_desc(a0 + kProd[p][0] * kABytes + ks * 256, 128, (kK / 8) * 128);
                const uint64_t db = make_desc(b0 + kProd[p][1] * kBBytes + ks * 256, 128, (kK / 8) * 128);
                mma_bf16(acc, da, db, idesc, (p | ks) != 0);
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + (i & 1))) : "memory");
    };
    const int total = groups_per_cta * n_stages;
    if (tid == 0) issue(0);
    uint32_t parity[2] = {0, 0};
    for (int i = 0; i < total; ++i) {
        if (tid == 0 && i + 1 < total) issue(i + 1);          // buffer (i + 1) & 1 was read back in step i - 1 (barrier below)
        mbar_wait(smem_u32(bar + (i & 1)), parity[i & 1]);
        parity[i & 1] ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[32];
        const uint32_t taddr = tmem + 32 * (i & 1) + ((uint32_t)(32 * warp) << 16);
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
        if (i == 0 && blockIdx.x == 0 && d_out != nullptr)
            for (int c = 0; c < 32; ++c) d_out[tid * 32 + c] = __uint_as_float(r[c]);
#pragma unroll
        for (int c = 0; c < 32; ++c) keep += __uint_as_float(r[c]);            // the values are used
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                       // every warp has read buffer i & 1: it may be overwritten
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (keep == 123.456f && d_out != nullptr) d_out[0] = keep;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

static uint16_t bf16_rne(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x7FFF + ((u >> 16) & 1);
    return (uint16_t)(u >> 16);
}
static float bf16_to_f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

int main() {
    // A: 128 x 32 fp32 values of audio-like dynamic range, B: the 32 x 32 real matrix of the 16-point complex DFT;
    // both split into three bf16 pieces
    std::vector<float> A(kM * kK), B(kN * kK);
    srand(1);
    for (auto& v : A) v = 3000.f * ((float)rand() / RAND_MAX - 0.5f);
    for (int n = 0; n < 32; ++n)
        for (int k = 0; k < 32; ++k) {                                       // B[n][k]: output column n, input k
            const int ko = k & 15, no = n & 15;
            const double a = -2.0 * M_PI * ko * no / 16.0;
            const double re = cos(a), im = sin(a);
            double v;
            if (n < 16) v = k < 16 ? re : -im;                               // Re out = sum re*xr - im*xi
            else v = k < 16 ? im : re;                                       // Im out = sum im*xr + re*xi
            B[n * kK + k] = (float)v;
        }
    std::vector<uint16_t> As(3 * kM * kK), Bs(3 * kN * kK);
    auto split3 = [](float x, uint16_t* out, size_t stride) {
        float r = x;
        for (int i = 0; i < 3; ++i) {
            out[i * stride] = bf16_rne(r);
            r -= bf16_to_f(out[i * stride]);
        }
    };
    for (int m = 0; m < kM; ++m)
        for (int k = 0; k < kK; ++k) split3(A[m * kK + k], &As[op_index(m, k, kK)], (size_t)kM * kK);
    for (int n = 0; n < kN; ++n)
        for (int k = 0; k < kK; ++k) split3(B[n * kK + k], &Bs[op_index(n, k, kK)], (size_t)kN * kK);
    uint16_t *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, As.size() * 2));
    CK(cudaMalloc(&dB, Bs.size() * 2));
    CK(cudaMalloc(&dD, kM * kN * 4));
    CK(cudaMemcpy(dA, As.data(), As.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bs.data(), Bs.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(dft_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);

    dft_tc_kernel<<<1, 128, kSmem>>>(dA, dB, dD, 1, 1);
    CK(cudaDeviceSynchronize());
    std::vector<float> D(kM * kN);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0.0, scale = 0.0;
    for (int m = 0; m < kM; ++m)
        for (int n = 0; n < kN; ++n) {
            double ref = 0.0;
            for (int k = 0; k < kK; ++k) ref += (double)A[m * kK + k] * (double)B[n * kK + k];
            worst = fmax(worst, fabs(ref - D[m * kN + n]));
            scale = fmax(scale, fabs(ref));
        }
    printf("GEMM check (six split bf16 products vs fp64): max |err| %.3e on values up to %.3e (relative %.2e)\n", worst, scale, worst / scale);

    const int frames = 150368, groups = frames / 8;                          // 18 796 groups of 8 frames
    const int per_cta = (groups + sms - 1) / sms;
    for (int stages = 1; stages <= 2; ++stages) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i) dft_tc_kernel<<<sms, 128, kSmem>>>(dA, dB, nullptr, per_cta, stages);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 20; ++i) dft_tc_kernel<<<sms, 128, kSmem>>>(dA, dB, nullptr, per_cta, stages);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / 20.0;
        const double flop = 2.0 * kM * kN * kK * 6.0 * stages * per_cta * sms;
        printf("%d DFT stage(s): %.1f us per %d frames (%d CTAs x %d groups of 8): %.1f TFLOP/s of bf16 tensor work, "
               "MMA issue + accumulator read-back only\n", stages, us, per_cta * sms * 8, sms, per_cta, flop / us / 1e6);
    }
    return 0;
}
