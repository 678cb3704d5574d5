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

// LTB_FRONTEND_TC_INT (csrc/ltb_tc_frontend.cuh): the D = 16 taps quantised to T[j] = rint(taps[j] * 2^27), and
// the tap table of the tensor-core kernel in its shared-memory image ([208 rows][128 B], K-major, 128-byte
// swizzle, the first 32 bytes of a row used).  sc16 (fmt 1): row 4 d + v holds digit v (lo byte) / digit v - 1
// (hi byte) of T[16 d - p'], p' = 0..15; sc8 (fmt 2): digit v of T[16 d - p'], p' = 0..31; fc32 as 23-bit fixed point
// (fmt 0): two tables, bytes 0..31 (samples 0..7 of an output's 16) and 32..63 (samples 8..15), byte 4 p8 + bi of
// row 4 d + v' holds digit v' + 1 - bi of T[16 d - 8 h - p8].  sum_t receives sum_j T[j].
std::vector<int32_t> make_tc_taps(long long *sum_t);
std::vector<int8_t> make_tc_btab(int fmt);

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
