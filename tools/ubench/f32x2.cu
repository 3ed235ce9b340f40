// Microbenchmark: throughput of scalar FFMA/FADD vs packed fma/add.f32x2 on sm_100a (developer tool).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

template <int MODE>
__global__ void k(float* out, int iters, float s) {
    // 16 independent chains per thread
    float a[16];
    unsigned long long p[8];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
    const unsigned long long ss = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 0.5f);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], ss, ss);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = a[i] + s;
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fadd2(p[i], ss);
        } else if (MODE == 5) {   // packed mul
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(p[i]) : "l"(p[i]), "l"(ss));
        } else if (MODE == 6) {   // 8 packed adds + 8 scalar FMUL
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fadd2(p[i], ss); a[i] = a[i] * s; }
        } else if (MODE == 7) {   // 8 packed adds + 8 integer LOP3/IADD (ALU pipe)
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fadd2(p[i], ss); a[i] = __int_as_float((__float_as_int(a[i]) ^ it) + i); }
        } else if (MODE == 4) {   // mixed: 8 scalar FFMA + 4 packed per iteration (same flops as mode 0)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], s, 0.5f);
#pragma unroll
            for (int i = 4; i < 8; ++i) p[i] = ffma2(p[i], ss, ss);
        }
    }
    float r = 0;
    for (int i = 0; i < 16; ++i) r += a[i];
    for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, float* d, double flops_per_iter_thread, int per_sm = 8) {
    const int iters = 4096, blocks = 148 * per_sm, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, 64, 1.0001f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, iters, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double tf = flops_per_iter_thread * iters * blocks * threads / (ms * 1e-3) / 1e12;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s (fp32)\n", name, ms, tf);
}

int main() {
    float* d;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("scalar FFMA x16", d, 32);
    run<1>("fma.rn.f32x2 x8", d, 32);
    run<2>("scalar FADD x16", d, 16);
    run<3>("add.rn.f32x2 x8", d, 16);
    run<4>("8 FFMA + 4 FFMA2", d, 32);
    run<5>("mul.rn.f32x2 x8", d, 16);
    run<6>("8 FADD2 + 8 FMUL", d, 24);
    run<7>("8 FADD2 + 8x(LOP3+IADD)", d, 16);
    printf("-- 2 CTAs (16 warps) per SM --\n");
    run<0>("scalar FFMA x16", d, 32, 2);
    run<1>("fma.rn.f32x2 x8", d, 32, 2);
    run<2>("scalar FADD x16", d, 16, 2);
    run<3>("add.rn.f32x2 x8", d, 16, 2);
    run<5>("mul.rn.f32x2 x8", d, 16, 2);
    run<6>("8 FADD2 + 8 FMUL", d, 24, 2);
    run<7>("8 FADD2 + 8x(LOP3+IADD)", d, 16, 2);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
