// In-register FFT building blocks shared by the sm_100a kernels and the host-side
// emulator used by the CPU tests (tests/test_host_emul.py).
//
// One real 512-point frame (kaldi.py:616 rfft of the zero-padded 400-sample frame h) is computed
// as a 256-point complex FFT of z[n] = h[2n] + i h[2n+1] followed by the real-FFT untangle.  The
// 256-point FFT is split Cooley-Tukey style as 256 = 16 x 16.  A group of 16 threads owns TWO
// frames:
//   stage A  thread tau holds z[16*n1 + tau], n1 = 0..15, of both frames and runs a 16-point DIF
//            FFT per frame in registers (n1 >= 13 are zero padding: 16*13 >= 200),
//   twiddle  Y_tau[k1] *= W256^(tau*k1),
//   exchange through shared memory (half-warp local),
//   stage B  thread t = 8*f + u runs, for frame f, two 16-point FFTs over tau for the rows
//            k1 in {u, 16 - u} ({0, 8} for u = 0), giving Z[k1 + 16*k2],
//   untangle X[k] = E + W512^k O and X[256-k] = conj(E - W512^k O) from Z[k] and Z[256-k], which
//            by construction live in the same thread; only |X|^2 is formed.
// All loops are unrolled at compile time with constant indices so every array stays in
// registers and every twiddle is an immediate.
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
template <int J, int M, bool B_ZERO>
OE_HD void dif_butterfly(float& ar, float& ai, float& br, float& bi) {
    float dr, di;
    if constexpr (B_ZERO) {
        dr = ar;
        di = ai;
    } else {
        const float ur = ar, ui = ai;
        ar = ur + br;
        ai = ui + bi;
        dr = ur - br;
        di = ui - bi;
    }
    if constexpr (J == 0) {
        br = dr;
        bi = di;
    } else if constexpr (4 * J == M) {            // W = -i
        br = di;
        bi = -dr;
    } else if constexpr (8 * J == M) {            // W = (1 - i)/sqrt2
        constexpr float r = 0.70710678118654752440f;
        br = (dr + di) * r;
        bi = (di - dr) * r;
    } else if constexpr (8 * J == 3 * M) {        // W = (-1 - i)/sqrt2
        constexpr float r = 0.70710678118654752440f;
        br = (di - dr) * r;
        bi = -(dr + di) * r;
    } else {
        constexpr float c = static_cast<float>(cos2pi(J, M));
        constexpr float s = static_cast<float>(sin2pi(J, M));
        br = dr * c + di * s;                     // (dr + i di)(c - i s)
        bi = di * c - dr * s;
    }
}

template <int N, int HALF, int ZERO_FROM>
struct DifStage {
    static OE_HD void run(float (&re)[N], float (&im)[N]) {
        static_for<0, N / (2 * HALF)>([&](auto blk) {
            static_for<0, HALF>([&](auto jj) {
                constexpr int j = decltype(jj)::value;
                constexpr int a = decltype(blk)::value * 2 * HALF + j;
                constexpr int b = a + HALF;
                dif_butterfly<j, 2 * HALF, (b >= ZERO_FROM)>(re[a], im[a], re[b], im[b]);
            });
        });
        if constexpr (HALF > 1) DifStage<N, HALF / 2, N>::run(re, im);   // later stages: no zeros
    }
};

// In-place N-point DIF FFT (forward, e^{-i...}).  Afterwards position i holds X[bitrev<N>(i)].
// Inputs at positions >= ZERO_FROM (only meaningful for ZERO_FROM > N/2) must be zero and are
// never read by the first stage.
template <int N, int ZERO_FROM = N>
OE_HD void fft_dif(float (&re)[N], float (&im)[N]) {
    static_assert(ZERO_FROM > N / 2, "pruning only covers the upper half");
    DifStage<N, N / 2, ZERO_FROM>::run(re, im);
}

// Rows of the 16 x 16 decomposition owned by stage-B lane u (0..7) of a frame.
OE_HD int stage_b_row_a(int u) { return u; }
OE_HD int stage_b_row_b(int u) { return u == 0 ? 8 : 16 - u; }

// Real-FFT untangle of one conjugate pair.  P = Z[k], Q = Z[256-k], (c, s) = (cos, sin)(2*pi*k/512).
// Returns 4*|X[k]|^2 and 4*|X[256-k]|^2 (the 1/4 is folded into the mel weights).
OE_HD void untangle_power(float pr, float pi, float qr, float qi, float c, float s,
                          float& pk, float& pnk) {
    const float er = pr + qr, ei = pi - qi;      // 2E = P + conj(Q)
    const float orr = pi + qi, oi = qr - pr;     // 2O = (P - conj(Q)) / i
    const float tr = c * orr + s * oi;           // T = (c - i s) * 2O
    const float ti = c * oi - s * orr;
    const float ar = er + tr, ai = ei + ti;      // 2X[k]
    const float br = er - tr, bi = ei - ti;      // 2 conj(X[256-k])
    pk = ar * ar + ai * ai;
    pnk = br * br + bi * bi;
}

}  // namespace oe
