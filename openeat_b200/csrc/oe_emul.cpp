// Host-side emulation of the per-frame math of the fbank kernel (stage A -> twiddle -> exchange ->
// stage B -> row exchange -> untangle), running the very same templates from oe_fft.h that the sm_100a
// kernel instantiates -- with the packed two-frame type V2 emulated as a pair of floats.  Built with g++
// into liboe_emul.so for tests/test_host_emul.py; it is test tooling and is never loaded by the product path.
#include <cmath>

#include "oe_fft.h"

extern "C" {

// ha, hb: two windowed frames (400 samples each; the kernel builds them from shared memory).
// pa, pb: 256 power-spectrum bins |X[k]|^2, k = 0..255, per frame.
void oe_emul_frame_pair(const float* ha, const float* hb, float* pa, float* pb) {
    static oe::V2 Er[16][16], Ei[16][16];          // first exchange: [k1][tau]
    static oe::V2 Zr[16][16], Zi[16][16];          // second exchange: [k1][k2]
    for (int tau = 0; tau < 16; ++tau) {           // ---- stage A ----
        oe::V2 zr[16], zi[16];
        for (int n1 = 0; n1 < 16; ++n1) {
            const int j = 2 * (16 * n1 + tau);
            zr[n1] = oe::v2_make(j < 400 ? ha[j] : 0.f, j < 400 ? hb[j] : 0.f);
            zi[n1] = oe::v2_make(j + 1 < 400 ? ha[j + 1] : 0.f, j + 1 < 400 ? hb[j + 1] : 0.f);
        }
        oe::fft_dif<16, 13, oe::V2>(zr, zi);
        for (int pos = 0; pos < 16; ++pos) {
            const int k1 = oe::bitrev<16>(pos);
            const double ang = 2.0 * oe::kPi * (double)((tau * k1) % 256) / 256.0;
            const oe::V2 c = oe::vbcast((float)std::cos(ang)), s = oe::vbcast((float)-std::sin(ang));
            Er[k1][tau] = oe::vsub(oe::vmul(zr[pos], c), oe::vmul(zi[pos], s));
            Ei[k1][tau] = oe::vfma(zr[pos], s, oe::vmul(zi[pos], c));
        }
    }
    for (int k1 = 0; k1 < 16; ++k1) {              // ---- stage B: one row per lane ----
        oe::V2 ar[16], ai[16];
        for (int n2 = 0; n2 < 16; ++n2) {
            ar[n2] = Er[k1][n2];
            ai[n2] = Ei[k1][n2];
        }
        oe::fft_dif<16, 16, oe::V2>(ar, ai);
        for (int pos = 0; pos < 16; ++pos) {
            Zr[k1][oe::bitrev<16>(pos)] = ar[pos];
            Zi[k1][oe::bitrev<16>(pos)] = ai[pos];
        }
    }
    for (int k1 = 0; k1 < 16; ++k1) {              // ---- untangle: partner row, one output per (k1, k2) ----
        const int pr = oe::partner_row(k1);
        for (int k2 = 0; k2 < 16; ++k2) {
            const int k = k1 + 16 * k2, q2 = oe::partner_k2(k1, k2);
            const oe::V2 c = oe::vbcast((float)std::cos(2.0 * oe::kPi * k / 512.0));
            const oe::V2 s = oe::vbcast((float)std::sin(2.0 * oe::kPi * k / 512.0));
            const oe::V2 p = oe::untangle_power(Zr[k1][k2], Zi[k1][k2], Zr[pr][q2], Zi[pr][q2], c, s);
            pa[k] = 0.25f * oe::v2_lo(p);
            pb[k] = 0.25f * oe::v2_hi(p);
        }
    }
}

// Second-generation decomposition (oe_fbank2_kernel.cuh): radix-4 stages, half-size partner exchange and the
// pair untangle.  Lane k1 keeps Z[k1 + 16 k2] for k2 = 0..7, publishes k2 = 8..15 and pairs its k2 with the
// partner row's 15 - k2; row 0 is its own partner with k2 <-> 16 - k2, which the kernel expresses as a one-slot
// offset of lane 0's read address; bin 128 = |Z[128]|^2 is lane 0's extra output; bins 0 and 256 carry no mel
// weight and hold garbage.
void oe_emul_frame_pair_v2(const float* ha, const float* hb, float* pa, float* pb) {
    static oe::V2 Er[16][16], Ei[16][16];          // first exchange: [k1][tau]
    static oe::V2 Pr[16][9], Pi[16][9];            // published half rows: slot k2 - 8 (+ one unwritten pad slot)
    static oe::V2 Kr[16][16], Ki[16][16];          // lane-resident Z[k1][k2]
    for (int tau = 0; tau < 16; ++tau) {           // ---- stage A ----
        oe::V2 zr[16], zi[16];
        for (int n1 = 0; n1 < 16; ++n1) {
            const int j = 2 * (16 * n1 + tau);
            zr[n1] = oe::v2_make(j < 400 ? ha[j] : 0.f, j < 400 ? hb[j] : 0.f);
            zi[n1] = oe::v2_make(j + 1 < 400 ? ha[j + 1] : 0.f, j + 1 < 400 ? hb[j + 1] : 0.f);
        }
        oe::fft16_r4<true, oe::V2>(zr, zi);
        for (int k1 = 0; k1 < 16; ++k1) {
            const int pos = oe::r4pos(k1);
            const double ang = 2.0 * oe::kPi * (double)((tau * k1) % 256) / 256.0;
            oe::V2 r = zr[pos], i = zi[pos];
            if (k1 != 0) oe::cmul_lane(r, i, (float)std::cos(ang), (float)std::sin(ang));
            Er[k1][tau] = r;
            Ei[k1][tau] = i;
        }
    }
    for (int k1 = 0; k1 < 16; ++k1) {              // ---- stage B + publish ----
        oe::V2 ar[16], ai[16];
        for (int n2 = 0; n2 < 16; ++n2) {
            ar[n2] = Er[k1][n2];
            ai[n2] = Ei[k1][n2];
        }
        oe::fft16_r4<false, oe::V2>(ar, ai);
        for (int k2 = 0; k2 < 16; ++k2) {
            Kr[k1][k2] = ar[oe::r4pos(k2)];
            Ki[k1][k2] = ai[oe::r4pos(k2)];
            if (k2 >= 8) {
                Pr[k1][k2 - 8] = Kr[k1][k2];
                Pi[k1][k2 - 8] = Ki[k1][k2];
            }
        }
        Pr[k1][8] = Pi[k1][8] = oe::vbcast(std::nanf(""));     // the pad slot lane 0 reads for its (unused) slot 0
    }
    for (int k = 0; k <= 256; ++k) pa[k] = pb[k] = std::nanf("");
    for (int k1 = 0; k1 < 16; ++k1) {              // ---- pair untangle ----
        const int prow = (16 - k1) & 15;
        const int shift = k1 == 0 ? 1 : 0;         // lane 0: slot 8 - k2 instead of 7 - k2
        for (int k2 = 0; k2 < 8; ++k2) {
            const int k = k1 + 16 * k2;
            const float c = (float)std::cos(2.0 * oe::kPi * k / 512.0), s = (float)std::sin(2.0 * oe::kPi * k / 512.0);
            oe::V2 pk, pq;
            oe::untangle_pair<oe::V2>(Kr[k1][k2], Ki[k1][k2], Pr[prow][7 - k2 + shift], Pi[prow][7 - k2 + shift], c, s, pk, pq);
            pa[k] = 0.25f * oe::v2_lo(pk);
            pb[k] = 0.25f * oe::v2_hi(pk);
            pa[256 - k] = 0.25f * oe::v2_lo(pq);
            pb[256 - k] = 0.25f * oe::v2_hi(pq);
        }
        if (k1 == 0) {                             // bin 128: X[128] = conj(Z[128])
            const oe::V2 p = oe::vfma(Kr[0][8], Kr[0][8], oe::vmul(Ki[0][8], Ki[0][8]));
            pa[128] = oe::v2_lo(p);
            pb[128] = oe::v2_hi(p);
        }
    }
}

}  // extern "C"

// ---- the GPU FLAC decoder's reader / predictor / CRC code run on the host, frame by frame (tests/test_flac.py checks it
// against the host decoder without a GPU; `-m gpu` tests then run the real kernel) ----
#define __device__
#define __forceinline__ inline
#include "oe_flac_gpu.cuh"

extern "C" void oe_emul_flac_decode(const unsigned char* comp, long long comp_bytes, const oe_flac_frame* frames, long long n_frames,
                                    short* pcm, int* errors, int verify_crc) {
    for (long long f = 0; f < n_frames; ++f) oe_flacgpu::emulate_frame(comp, comp_bytes + 16, frames[f], pcm, errors, verify_crc);
}
