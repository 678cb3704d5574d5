// Host-side constant tables of the PSS+SSS search path (see DESIGN.md "Canonical tables").
#pragma once
#include <cstdint>
#include <vector>

namespace ltb {

struct PssTaps { float re[128]; float im[128]; };

// srslte_pss_init_N_id_2 restated: conj(IDFT_128(ZC_u))/sqrt(128)/62, taps 0..64 rounded to
// float, 65..127 mirrored; N_id_2 = 2 is the exact conjugate of N_id_2 = 1.
void make_pss_taps(int n_id_2, PssTaps &out);

// gr-filter rational_resampler.design_filter(1, decim, 0.4) -> firdes.low_pass(Kaiser, beta 7)
std::vector<float> make_decim_taps(int decim);
// the same taps by polyphase branch, zero padded: out[v * 33 + q] = taps[q * decim + v]; empty if
// a branch would need more than 33 taps (never for decim <= 64)
std::vector<float> make_decim_branch_taps(int decim);

// LTB_FRONTEND_TC_INT (csrc/ltb_tc_frontend.cuh): the taps of rate `decim` quantised to T[j] = rint(taps[j] * 2^shift),
// shift = 23 + floor(log2(decim)) (27 at decim 16: |T| < 2^23, three balanced base-256 digits), and the tap table of the
// tensor-core kernel in its shared-memory image ([208 rows][128 B], K-major, 128-byte swizzle).  A k-step is 32 bytes of
// one component (sc16: 16 samples x {lo, hi}; sc8: 32 samples; fc32 as 23-bit fixed point: 8 samples x {b0, b1, b2, -});
// row 4 d + v of the table of phase ph (bytes 32 i .. 32 i + 31 of the row, ph = i gcd(samples per k-step, decim)) holds
// at each byte the digit of T[decim d - ph - p] that byte of sample p multiplies into weight 256^v (fc32: 256^(v+1)).
// sum_t receives sum_j T[j].
int tc_tap_shift_for(int decim);
std::vector<int32_t> make_tc_taps(int decim, long long *sum_t);
std::vector<int8_t> make_tc_btab(int fmt, int decim);

struct SssTables {
  int32_t c0[31], c1[31], s_tilde[31], z_tilde[31];
  int32_t n_id_1[900];
};
void make_sss_tables(int n_id_2, SssTables &out);

// overlap-save FFT correlator: W_1024^i (exact at multiples of 256) and the filter spectra
// 2^-10 * DFT_1024(h zero padded), natural order
void make_fft1024_twiddles(float *re, float *im);  // 1024 entries
void make_os_filter(int n_id_2, float *H_re, float *H_im);   // 1024 entries

void make_cexp_table(float *re, float *im);        // 4097 entries
void make_fft128_twiddles(float *re, float *im);   // 64 entries

}  // namespace ltb
