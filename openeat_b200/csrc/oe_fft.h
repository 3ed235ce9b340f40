// In-register FFT building blocks shared by the sm_100a kernels and the host-side
// emulator used by the CPU tests (tests/test_host_emul.py).
//
// One real 512-point frame (kaldi.py:616 rfft of the zero-padded 400-sample frame h) is computed
// as a 256-point complex FFT of z[n] = h[2n] + i h[2n+1] followed by the real-FFT untangle.  The
// 256-point FFT is split Cooley-Tukey style as 256 = 16 x 16.  A group of 16 threads owns TWO
// frames, carried as the two halves of Blackwell's packed f32x2 registers (`V2`): every FADD2 /
// FMUL2 / FFMA2 serves both frames, halving the issue slots of the transform.
//   stage A  thread tau holds z[16*n1 + tau], n1 = 0..15, of both frames and runs one packed
//            16-point DIF FFT in registers (n1 >= 13 are zero padding: 16*13 >= 200),
//   twiddle  Y_tau[k1] *= W256^(tau*k1),
//   exchange through shared memory (half-warp local, XOR-swizzled),
//   stage B  lane k1 runs one packed 16-point FFT over tau for row k1, giving Z[k1 + 16*k2],
//   exchange of the rows (half-warp local): lane k1 fetches Z[256-k] from the conjugate row
//            (16 - k1) mod 16 -- rows 0 and 8 are their own partners, no special case,
//   untangle X[k] = E + W512^k O from Z[k] and Z[256-k]; only |X[k]|^2 is formed, k = 0..255.
// All loops are unrolled at compile time with constant indices so every array stays in
// registers and every twiddle is an immediate.  The same templates run on the host with plain
// floats (T = float) or an emulated pair (T = V2) for the CPU tests.
#pragma once
#include <type_traits>

#if defined(__CUDACC__)
#define OE_HD __host__ __device__ __forceinline__
#define OE_CX constexpr __host__ __device__
#else
#define OE_HD inline
#define OE_CX constexpr
#endif

namespace oe {

template <int I, int N, class F>
OE_HD void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// ---- packed pair of fp32 (two frames side by side) ----
struct V2 {
#if defined(__CUDA_ARCH__)
    unsigned long long v;     // one aligned 64-bit register pair: lo = frame 0, hi = frame 1
#else
    float lo, hi;
#endif
};

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ V2 v2_make(float lo, float hi) {
    V2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float v2_lo(V2 a) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
    return lo;
}
__device__ __forceinline__ float v2_hi(V2 a) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
    return hi;
}
__device__ __forceinline__ V2 vadd(V2 a, V2 b) {
    V2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ V2 vsub(V2 a, V2 b) {
    V2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ V2 vmul(V2 a, V2 b) {
    V2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ V2 vfma(V2 a, V2 b, V2 c) {     // a * b + c
    V2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
#else
inline V2 v2_make(float lo, float hi) { return V2{lo, hi}; }
inline float v2_lo(V2 a) { return a.lo; }
inline float v2_hi(V2 a) { return a.hi; }
inline V2 vadd(V2 a, V2 b) { return V2{a.lo + b.lo, a.hi + b.hi}; }
inline V2 vsub(V2 a, V2 b) { return V2{a.lo - b.lo, a.hi - b.hi}; }
inline V2 vmul(V2 a, V2 b) { return V2{a.lo * b.lo, a.hi * b.hi}; }
inline V2 vfma(V2 a, V2 b, V2 c) { return V2{a.lo * b.lo + c.lo, a.hi * b.hi + c.hi}; }
#endif
OE_HD V2 vbcast(float s) { return v2_make(s, s); }
OE_HD V2 vneg(V2 a) { return vsub(vbcast(0.f), a); }

// scalar overloads so the FFT templates also instantiate with T = float
OE_HD float vadd(float a, float b) { return a + b; }
OE_HD float vsub(float a, float b) { return a - b; }
OE_HD float vmul(float a, float b) { return a * b; }
OE_HD float vfma(float a, float b, float c) { return a * b + c; }
OE_HD float vneg(float a) { return -a; }
template <class T> OE_HD T vconst(float s);
template <> OE_HD float vconst<float>(float s) { return s; }
template <> OE_HD V2 vconst<V2>(float s) { return vbcast(s); }

// ---- compile-time trigonometry (double Taylor series after octant reduction) ----
constexpr double kPi = 3.14159265358979323846264338327950288;

OE_CX double taylor_sin(double x) {   // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 12; ++i) {
        term *= -x2 / ((2 * i) * (2 * i + 1));
        sum += term;
    }
    return sum;
}
OE_CX double taylor_cos(double x) {   // |x| <= pi/4
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 12; ++i) {
        term *= -x2 / ((2 * i - 1) * (2 * i));
        sum += term;
    }
    return sum;
}
// cos / sin of 2*pi*k/n for integer k, exact symmetries first.
OE_CX double cos2pi(int k, int n) {
    k %= n;
    if (k < 0) k += n;
    if (2 * k > n) k = n - k;                 // cos(2pi - x) = cos x
    if (4 * k > n) return -cos2pi(n - 2 * k, 2 * n);   // cos(pi - y), y = 2pi*(n/2-k)/n
    if (8 * k > n) {                          // cos x = sin(pi/2 - x)
        return taylor_sin(2.0 * kPi * (n - 4 * k) / (4.0 * n));
    }
    return taylor_cos(2.0 * kPi * k / n);
}
OE_CX double sin2pi(int k, int n) {
    k %= n;
    if (k < 0) k += n;
    if (2 * k > n) return -sin2pi(n - k, n);
    if (4 * k > n) return sin2pi(n - 2 * k, 2 * n);     // sin(pi - y)
    if (8 * k > n) {                          // sin x = cos(pi/2 - x)
        return taylor_cos(2.0 * kPi * (n - 4 * k) / (4.0 * n));
    }
    return taylor_sin(2.0 * kPi * k / n);
}

template <int N>
OE_CX int bitrev(int i) {
    int r = 0;
    for (int b = 1; b < N; b <<= 1) {
        r = (r << 1) | (i & 1);
        i >>= 1;
    }
    return r;
}

// One radix-2 decimation-in-frequency butterfly with twiddle W_M^J = exp(-2*pi*i*J/M):
//   a' = a + b,  b' = (a - b) * W.   B_ZERO: b is known to be zero (pruned padding).
template <int J, int M, bool B_ZERO, class T>
OE_HD void dif_butterfly(T& ar, T& ai, T& br, T& bi) {
    T dr, di;
    if constexpr (B_ZERO) {
        dr = ar;
        di = ai;
    } else {
        const T ur = ar, ui = ai;
        ar = vadd(ur, br);
        ai = vadd(ui, bi);
        dr = vsub(ur, br);
        di = vsub(ui, bi);
    }
    if constexpr (J == 0) {
        br = dr;
        bi = di;
    } else if constexpr (4 * J == M) {            // W = -i
        br = di;
        bi = vneg(dr);
    } else if constexpr (8 * J == M) {            // W = (1 - i)/sqrt2
        const T r = vconst<T>(0.70710678118654752440f);
        br = vmul(vadd(dr, di), r);
        bi = vmul(vsub(di, dr), r);
    } else if constexpr (8 * J == 3 * M) {        // W = (-1 - i)/sqrt2
        const T r = vconst<T>(0.70710678118654752440f);
        br = vmul(vsub(di, dr), r);
        bi = vmul(vadd(dr, di), vconst<T>(-0.70710678118654752440f));
    } else {
        constexpr float c = static_cast<float>(cos2pi(J, M));
        constexpr float s = static_cast<float>(sin2pi(J, M));
        br = vfma(dr, vconst<T>(c), vmul(di, vconst<T>(s)));     // (dr + i di)(c - i s)
        bi = vfma(di, vconst<T>(c), vmul(dr, vconst<T>(-s)));
    }
}

template <int N, int HALF, int ZERO_FROM, class T>
struct DifStage {
    static OE_HD void run(T (&re)[N], T (&im)[N]) {
        static_for<0, N / (2 * HALF)>([&](auto blk) {
            static_for<0, HALF>([&](auto jj) {
                constexpr int j = decltype(jj)::value;
                constexpr int a = decltype(blk)::value * 2 * HALF + j;
                constexpr int b = a + HALF;
                dif_butterfly<j, 2 * HALF, (b >= ZERO_FROM), T>(re[a], im[a], re[b], im[b]);
            });
        });
        if constexpr (HALF > 1) DifStage<N, HALF / 2, N, T>::run(re, im);   // later stages: no zeros
    }
};

// In-place N-point DIF FFT (forward, e^{-i...}).  Afterwards position i holds X[bitrev<N>(i)].
// Inputs at positions >= ZERO_FROM (only meaningful for ZERO_FROM > N/2) must be zero and are
// never read by the first stage.
template <int N, int ZERO_FROM = N, class T = float>
OE_HD void fft_dif(T (&re)[N], T (&im)[N]) {
    static_assert(ZERO_FROM > N / 2, "pruning only covers the upper half");
    DifStage<N, N / 2, ZERO_FROM, T>::run(re, im);
}

// Conjugate partner of Z[k1 + 16 k2] in the 16 x 16 layout: Z[256 - k] sits in row (16 - k1) mod 16 at
// index 15 - k2 (k1 >= 1) or (16 - k2) mod 16 (k1 == 0: k = 16 k2, 256 - k = 16 (16 - k2)).
OE_HD int partner_row(int k1) { return (16 - k1) & 15; }
OE_HD int partner_k2(int k1, int k2) { return k1 == 0 ? ((16 - k2) & 15) : 15 - k2; }

// Real-FFT untangle, one output: P = Z[k], Q = Z[256-k], (c, s) = (cos, sin)(2*pi*k/512).
// Returns 4*|X[k]|^2 (the 1/4 is folded into the mel weights).
template <class T>
OE_HD T untangle_power(T pr, T pi, T qr, T qi, T c, T s) {
    const T er = vadd(pr, qr), ei = vsub(pi, qi);      // 2E = P + conj(Q)
    const T orr = vadd(pi, qi), oi = vsub(qr, pr);     // 2O = (P - conj(Q)) / i
    const T ar = vfma(s, oi, vfma(c, orr, er));        // 2X = 2E + (c - i s) * 2O
    const T ai = vsub(vfma(c, oi, ei), vmul(s, orr));
    return vfma(ar, ar, vmul(ai, ai));
}

}  // namespace oe
