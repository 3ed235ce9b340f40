// Host-side emulation of the per-frame math of the fbank kernel (stage A -> twiddle ->
// exchange -> stage B -> untangle), running the very same templates from oe_fft.h that the
// sm_100a kernel instantiates.  Built with g++ into liboe_emul.so for tests/test_host_emul.py;
// it is test tooling and is never loaded by the product path.
#include <cmath>

#include "oe_fft.h"

namespace {
struct Cf {
    float re, im;
};
}  // namespace

extern "C" {

// h: one windowed frame (400 samples; the kernel builds it from shared memory).
// pw: 257 power-spectrum bins |X[k]|^2.
void oe_emul_frame(const float* h, float* pw) {
    Cf E[16][16];                                  // exchange buffer: [k1][tau]
    for (int tau = 0; tau < 16; ++tau) {           // ---- stage A ----
        float zr[16], zi[16];
        for (int n1 = 0; n1 < 16; ++n1) {
            const int j = 2 * (16 * n1 + tau);
            zr[n1] = (j < 400) ? h[j] : 0.0f;
            zi[n1] = (j + 1 < 400) ? h[j + 1] : 0.0f;
        }
        oe::fft_dif<16, 13>(zr, zi);
        for (int pos = 0; pos < 16; ++pos) {
            const int k1 = oe::bitrev<16>(pos);
            const double ang = -2.0 * oe::kPi * (double)(tau * k1) / 256.0;
            const float c = (float)std::cos(ang), s = (float)std::sin(ang);
            E[k1][tau].re = zr[pos] * c - zi[pos] * s;
            E[k1][tau].im = zr[pos] * s + zi[pos] * c;
        }
    }
    for (int u = 0; u < 8; ++u) {                  // ---- stage B ----
        const int ra = oe::stage_b_row_a(u), rb = oe::stage_b_row_b(u);
        float ar[16], ai[16], br[16], bi[16];
        for (int n2 = 0; n2 < 16; ++n2) {
            ar[n2] = E[ra][n2].re;
            ai[n2] = E[ra][n2].im;
            br[n2] = E[rb][n2].re;
            bi[n2] = E[rb][n2].im;
        }
        oe::fft_dif<16>(ar, ai);
        oe::fft_dif<16>(br, bi);
        auto tw = [](int k, float& c, float& s) {
            c = (float)std::cos(2.0 * oe::kPi * k / 512.0);
            s = (float)std::sin(2.0 * oe::kPi * k / 512.0);
        };
        float c, s, pk, pnk;
        if (u != 0) {
            for (int k2 = 0; k2 < 16; ++k2) {      // P = Z[u + 16 k2], Q = Z[(16-u) + 16 (15-k2)]
                const int p = oe::bitrev<16>(k2), q = oe::bitrev<16>(15 - k2);
                const int k = u + 16 * k2;
                tw(k, c, s);
                oe::untangle_power(ar[p], ai[p], br[q], bi[q], c, s, pk, pnk);
                pw[k] = 0.25f * pk;
                pw[256 - k] = 0.25f * pnk;
            }
        } else {
            for (int k2 = 0; k2 <= 8; ++k2) {      // row 0: P = Z[16 k2], Q = Z[16 ((16-k2) mod 16)]
                const int p = oe::bitrev<16>(k2), q = oe::bitrev<16>((16 - k2) & 15);
                const int k = 16 * k2;
                tw(k, c, s);
                oe::untangle_power(ar[p], ai[p], ar[q], ai[q], c, s, pk, pnk);
                pw[k] = 0.25f * pk;
                pw[256 - k] = 0.25f * pnk;
            }
            for (int k2 = 0; k2 < 8; ++k2) {       // row 8: P = Z[8 + 16 k2], Q = Z[8 + 16 (15-k2)]
                const int p = oe::bitrev<16>(k2), q = oe::bitrev<16>(15 - k2);
                const int k = 8 + 16 * k2;
                tw(k, c, s);
                oe::untangle_power(br[p], bi[p], br[q], bi[q], c, s, pk, pnk);
                pw[k] = 0.25f * pk;
                pw[256 - k] = 0.25f * pnk;
            }
        }
    }
}

}  // extern "C"
