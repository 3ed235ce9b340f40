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

// ---------------------------------------------------------------------------------------------
// Second-generation blocks (oe_fbank2_kernel): radix-4 16-point FFT and the pair untangle.
//
// fft16_r4: two radix-4 DIF passes.  Pass 1 combines positions (j, j+4, j+8, j+12) and leaves output m
// at position j + 4m, times W16^(j m); pass 2 is a 4-point DFT over positions 4m .. 4m+3.  Afterwards
// position p holds X[r4pos(p)] (base-4 digit reversal, an involution).  The W16^4 = -i twiddle is folded
// into the adds of pass 2, so the transform has no negations: 128 add/sub + 32 twiddle operations
// (12 fewer add/sub with PRUNE13: inputs 13..15 are zero and never read).
OE_CX int r4pos(int k) { return 4 * (k & 3) + (k >> 2); }

// 4-point forward DFT in place: a <- out0, b <- out1, c <- out2, d <- out3.
// D_ZERO: d is zero (never read).  C_ROT: the c slot holds c0 and stands for -i * c0.
template <bool D_ZERO, bool C_ROT, class T>
OE_HD void bfly4(T& ar, T& ai, T& br, T& bi, T& cr, T& ci, T& dr, T& di) {
    T t0r, t0i, t1r, t1i, t2r, t2i, t3r, t3i;
    if constexpr (C_ROT) {                     // c = (ci, -cr)
        t0r = vadd(ar, ci);
        t0i = vsub(ai, cr);
        t1r = vsub(ar, ci);
        t1i = vadd(ai, cr);
    } else {
        t0r = vadd(ar, cr);
        t0i = vadd(ai, ci);
        t1r = vsub(ar, cr);
        t1i = vsub(ai, ci);
    }
    if constexpr (D_ZERO) {
        t2r = br;
        t2i = bi;
        t3r = br;
        t3i = bi;
    } else {
        t2r = vadd(br, dr);
        t2i = vadd(bi, di);
        t3r = vsub(br, dr);
        t3i = vsub(bi, di);
    }
    ar = vadd(t0r, t2r);
    ai = vadd(t0i, t2i);
    cr = vsub(t0r, t2r);
    ci = vsub(t0i, t2i);
    br = vadd(t1r, t3i);                       // out1 = t1 - i t3
    bi = vsub(t1i, t3r);
    dr = vsub(t1r, t3i);                       // out3 = t1 + i t3
    di = vadd(t1i, t3r);
}

// (r + i im) *= W16^E = exp(-2 pi i E / 16), E in {1, 2, 3, 6, 9}
template <int E, class T>
OE_HD void mul_w16(T& r, T& i) {
    if constexpr (E == 2) {
        const T h = vconst<T>(0.70710678118654752440f);
        const T a = vadd(r, i), b = vsub(i, r);
        r = vmul(a, h);
        i = vmul(b, h);
    } else if constexpr (E == 6) {
        const T a = vsub(i, r), b = vadd(r, i);
        r = vmul(a, vconst<T>(0.70710678118654752440f));
        i = vmul(b, vconst<T>(-0.70710678118654752440f));
    } else {
        constexpr float c = static_cast<float>(cos2pi(E, 16));
        constexpr float s = static_cast<float>(sin2pi(E, 16));
        const T nr = vfma(r, vconst<T>(c), vmul(i, vconst<T>(s)));     // (r + i im)(c - i s)
        const T ni = vfma(i, vconst<T>(c), vmul(r, vconst<T>(-s)));
        r = nr;
        i = ni;
    }
}

template <bool PRUNE13, class T>
OE_HD void fft16_r4(T (&re)[16], T (&im)[16]) {
    bfly4<false, false, T>(re[0], im[0], re[4], im[4], re[8], im[8], re[12], im[12]);
    bfly4<PRUNE13, false, T>(re[1], im[1], re[5], im[5], re[9], im[9], re[13], im[13]);
    bfly4<PRUNE13, false, T>(re[2], im[2], re[6], im[6], re[10], im[10], re[14], im[14]);
    bfly4<PRUNE13, false, T>(re[3], im[3], re[7], im[7], re[11], im[11], re[15], im[15]);
    mul_w16<1, T>(re[5], im[5]);               // position j + 4m carries W16^(j m); (j, m) = (2, 2) is folded below
    mul_w16<2, T>(re[9], im[9]);
    mul_w16<3, T>(re[13], im[13]);
    mul_w16<2, T>(re[6], im[6]);
    mul_w16<6, T>(re[14], im[14]);
    mul_w16<3, T>(re[7], im[7]);
    mul_w16<6, T>(re[11], im[11]);
    mul_w16<9, T>(re[15], im[15]);
    bfly4<false, false, T>(re[0], im[0], re[1], im[1], re[2], im[2], re[3], im[3]);
    bfly4<false, false, T>(re[4], im[4], re[5], im[5], re[6], im[6], re[7], im[7]);
    bfly4<false, true, T>(re[8], im[8], re[9], im[9], re[10], im[10], re[11], im[11]);
    bfly4<false, false, T>(re[12], im[12], re[13], im[13], re[14], im[14], re[15], im[15]);
}

// (r + i im) *= (c - i s) with a per-lane scalar twiddle shared by both packed frames.  Scalar FMUL/FFMA on
// the register halves: same FMA-pipe time as two packed operations and no broadcast moves.
OE_HD void cmul_lane(V2& r, V2& i, float c, float s) {
    const float rl = v2_lo(r), rh = v2_hi(r), il = v2_lo(i), ih = v2_hi(i);
    r = v2_make(rl * c + il * s, rh * c + ih * s);
    i = v2_make(il * c - rl * s, ih * c - rh * s);
}
OE_HD void cmul_lane(float& r, float& i, float c, float s) {
    const float nr = r * c + i * s, ni = i * c - r * s;
    r = nr;
    i = ni;
}

// Real-FFT untangle of the pair (k, 256 - k): P = Z[k], Q = Z[256 - k], (c, s) = (cos, sin)(2 pi k / 512).
// 2 X[k] = E + W O and 2 conj(X[256 - k]) = E - W O with E = P + conj(Q), O = (P - conj(Q)) / i share E, O and
// W O.  Returns 4 |X[k]|^2 in pk and 4 |X[256 - k]|^2 in pq (the 1/4 is folded into the mel weights).
template <class T>
OE_HD void untangle_pair(T pr, T pi, T qr, T qi, float c, float s, T& pk, T& pq) {
    const T er = vadd(pr, qr), ei = vsub(pi, qi);
    T orr = vadd(pi, qi), oi = vsub(qr, pr);
    cmul_lane(orr, oi, c, s);                                  // W O
    const T ar = vadd(er, orr), ai = vadd(ei, oi);
    const T br = vsub(er, orr), bi = vsub(ei, oi);
    pk = vfma(ar, ar, vmul(ai, ai));
    pq = vfma(br, br, vmul(bi, bi));
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
