/*
 * ltetrigger_oracle.c -- CPU oracle (test infrastructure, never shipped or linked by
 * the product).  See ltetrigger_oracle.h for scope, provenance and parity status.
 *
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=off -pthread -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off matters: every fused multiply-add below is an explicit fmaf()/fma(),
 * every other product/sum is individually rounded; the CUDA kernels do the same with
 * __fmaf_rn/__fmul_rn/__fadd_rn, so results can be compared bit for bit.
 *
 * [ref]  = /root/reference (read-only).  [A.x] = SURVEY.md Appendix A (srsLTE
 * release_18_06_1 / GNU Radio 3.7 semantics restated from memory; source absent).
 */
#include "ltetrigger_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <unistd.h>

#define NLAG ORC_CONV_LEN          /* 9726 */
#define AVG_LEN 9729               /* fft_size + frame_size + 1 [A.2] */
#define FFTN 9728                  /* frame_size + fft_size: the reference's conv FFT size [A.2] */
#define PSS_EMA_ALPHA 0.2f         /* q->ema_alpha default [A.2 step 3] */

/* fc32 input through the same front end: the product first puts each float on a 23-bit fixed-point grid over
 * +-full_scale (ltb_tc_frontend.cuh tc_split<LTB_FMT_FC32>; two fused multiply-adds, the first saturating to [0, 1]):
 *   u = sat(fma(x, 0.5 / full_scale, 0.5))     m = bits(fma(u, 2^23 - 2, 2^23 + 1)) & 0x7fffff     1 <= m <= 2^23 - 1
 * and then evaluates, exactly,
 *   A[k] = sum_j ( T[j] (m[16 k - j] - 2^22) - lo8(m[16 k - j]) * d0(T[j]) )
 * i.e. the sum of all byte x tap-digit products except the one of weight 1 (lowest sample byte times lowest balanced
 * tap digit: four accumulator columns per output hold the weights 256 .. 256^4, nothing is kept for weight 1; the
 * omitted term is some 2^-30 of full scale).  A is a multiple of 256;
 *   y[k] = (float)(A[k] / 256) * out_scale     out_scale = (float)(full_scale / (2^22 - 1) / 2^19)
 * Samples before the stream are zeros (m = 2^22, low byte 0: no contribution). */
static uint32_t tcq_fc32(float x, float q_inv)
{
  float u = fmaf(x, q_inv, 0.5f);
  if (!(u > 0.0f)) u = 0.0f;                 /* also NaN -> 0, as fma.rn.sat does */
  if (u > 1.0f) u = 1.0f;
  float t = fmaf(u, 8388606.0f, 8388609.0f);
  uint32_t b; memcpy(&b, &t, 4);
  return b & 0x7fffffu;
}

static int tcint_shift(int decim) { int l = 0; while (decim > 1) { decim >>= 1; l++; } return 23 + l; }   /* 27 at decim 16 */

int64_t orc_decimate_tcint_fc32(const orc_cf *x, int64_t n_in, float full_scale, orc_cf *y)
{
  return orc_decimate_tcint_fc32_d(x, n_in, 16, full_scale, y);
}

int64_t orc_decimate_tcint_fc32_d(const orc_cf *x, int64_t n_in, int decim, float full_scale, orc_cf *y)
{
  float taps[4096];
  int ntaps = orc_decim_taps(decim, taps, 4096);
  if (ntaps <= 0 || ntaps > 2112 || !(full_scale > 0.0f)) return -1;
  int32_t T[2112], D0[2112];
  const int shift = tcint_shift(decim);
  for (int j = 0; j < ntaps; j++) {
    T[j] = (int32_t)llrint(ldexp((double)taps[j], shift));
    D0[j] = ((T[j] + 128) & 255) - 128;      /* lowest balanced base-256 digit */
  }
  const float q_inv = (float)(0.5 / (double)full_scale);
  const float out_scale = (float)((double)full_scale / 4194303.0 * ldexp(1.0, 8 - shift));
  int32_t *m = malloc(sizeof(int32_t) * 2 * (size_t)n_in);
  for (int64_t i = 0; i < n_in; i++) { m[2 * i] = (int32_t)tcq_fc32(x[i].re, q_inv); m[2 * i + 1] = (int32_t)tcq_fc32(x[i].im, q_inv); }
  int64_t n_out = (n_in + decim - 1) / decim;
  for (int64_t k = 0; k < n_out; k++) {
    int64_t are = 0, aim = 0;
    const int64_t top = k * decim;
    const int jmax = top < ntaps - 1 ? (int)top : ntaps - 1;
    const int32_t *mp = m + 2 * top;
    for (int j = 0; j <= jmax; j++) {
      const int32_t mr = mp[-2 * j], mi = mp[-2 * j + 1];
      are += (int64_t)T[j] * (mr - 4194304) - (int64_t)(mr & 255) * D0[j];
      aim += (int64_t)T[j] * (mi - 4194304) - (int64_t)(mi & 255) * D0[j];
    }
    if ((are & 255) || (aim & 255)) { free(m); return -1; }      /* cannot happen: every kept product has weight >= 256 */
    y[k].re = (float)(are / 256) * out_scale;
    y[k].im = (float)(aim / 256) * out_scale;
  }
  free(m);
  return n_out;
}

/* ------------------------------------------------------------------------- */
/* Tables                                                                     */
/* ------------------------------------------------------------------------- */

static double cos128(int j) { j &= 127; return cos(2.0 * M_PI * (double)j / 128.0); }
static double sin128(int j) { j &= 127; return sin(2.0 * M_PI * (double)j / 128.0); }

/* Zadoff-Chu PSS in frequency, d_u[i], i=0..61 [A.1]; double precision, exact argument
 * reduction (u*k mod 126) so that d[i]==d[61-i] and d_34 == conj(d_29) hold to rounding. */
static void pss_freq(int root, double dre[62], double dim[62])
{
  for (int i = 0; i < 62; i++) {
    long k = (i < 31) ? (long)i * (i + 1) : (long)(i + 1) * (i + 2);
    long r = ((long)root * k) % 126;
    double ang = -M_PI * (double)r / 63.0;
    dre[i] = cos(ang);
    dim[i] = sin(ang);
  }
}

/* h = conj(t)/62, t = IDFT_128(bins)/sqrt(128) with d[0..30] on bins -31..-1 and
 * d[31..61] on bins +1..+31 [A.1].  Canonical form: taps 0..64 rounded to float, taps
 * 65..127 mirrored (h[128-m] = h[m]), N_id_2=2 defined as conj of N_id_2=1. */
void orc_pss_taps(int n_id_2, float h_re[128], float h_im[128])
{
  double dre[62], dim[62];
  pss_freq(n_id_2 == 0 ? 25 : 29, dre, dim);
  const double scale = 1.0 / sqrt(128.0) / 62.0;
  for (int n = 0; n <= 64; n++) {
    double tr = 0.0, ti = 0.0;
    for (int i = 0; i < 62; i++) {
      int b = (i < 31) ? i - 31 : i - 30;
      int j = ((b * n) % 128 + 128) % 128;
      double c = cos128(j), s = sin128(j);
      tr += dre[i] * c - dim[i] * s;
      ti += dre[i] * s + dim[i] * c;
    }
    h_re[n] = (float)(tr * scale);
    h_im[n] = (float)(-ti * scale);          /* conj */
    if (n_id_2 == 2) h_im[n] = -h_im[n];     /* root 34 = conj(root 29) */
  }
  for (int n = 65; n < 128; n++) { h_re[n] = h_re[128 - n]; h_im[n] = h_im[128 - n]; }
}

/* gr::fft::window::kaiser + firdes::low_pass + rational_resampler.design_filter [A.7] */
static double izero(double x)
{
  double sum = 1, u = 1, halfx = x / 2.0; int n = 1;
  do { double t = halfx / (double)n; n += 1; t *= t; u *= t; sum += u; } while (u >= 1e-21 * sum);
  return sum;
}

int orc_decim_taps(int decim, float *taps, int max_taps)
{
  if (decim <= 1) return 0;
  const double beta = 7.0, fractional_bw = 0.4, halfband = 0.5;
  double rate = 1.0 / (double)decim;
  double trans_width = rate * (halfband - fractional_bw);
  double mid = rate * halfband - trans_width / 2.0;
  double atten = beta / 0.1102 + 8.7;
  int ntaps = (int)(atten * 1.0 / (22.0 * trans_width));
  if ((ntaps & 1) == 0) ntaps++;
  if (ntaps > max_taps) return -1;
  float *w = (float *)malloc(sizeof(float) * ntaps);
  double ibeta = 1.0 / izero(beta), inm1 = 1.0 / (double)(ntaps - 1);
  for (int i = 0; i < ntaps; i++) {
    double t = 2 * i * inm1 - 1;
    w[i] = (float)(izero(beta * sqrt(1.0 - t * t)) * ibeta);
  }
  int M = (ntaps - 1) / 2;
  double fwT0 = 2 * M_PI * mid / 1.0;
  for (int n = -M; n <= M; n++) {
    if (n == 0) taps[n + M] = (float)(fwT0 / M_PI * w[n + M]);
    else        taps[n + M] = (float)(sin(n * fwT0) / (n * M_PI) * w[n + M]);
  }
  double fmax = taps[M];
  for (int n = 1; n <= M; n++) fmax += 2 * taps[n + M];
  double gain = 1.0 / fmax;
  for (int i = 0; i < ntaps; i++) taps[i] = (float)(taps[i] * gain);
  free(w);
  return ntaps;
}

/* 36.211 6.11.2.1 m-sequences and the srsLTE table layout [A.5][A.6] */
void orc_sss_tables(int n_id_2, int c0[31], int c1[31], int s_tilde[31], int z_tilde[31], int n_id_1_table[900])
{
  int x[31], c_tilde[31];
  memset(x, 0, sizeof x); x[4] = 1;
  for (int i = 0; i < 26; i++) x[i + 5] = (x[i + 2] + x[i]) % 2;
  for (int i = 0; i < 31; i++) s_tilde[i] = 1 - 2 * x[i];
  memset(x, 0, sizeof x); x[4] = 1;
  for (int i = 0; i < 26; i++) x[i + 5] = (x[i + 3] + x[i]) % 2;
  for (int i = 0; i < 31; i++) c_tilde[i] = 1 - 2 * x[i];
  memset(x, 0, sizeof x); x[4] = 1;
  for (int i = 0; i < 26; i++) x[i + 5] = (x[i + 4] + x[i + 2] + x[i + 1] + x[i]) % 2;
  for (int i = 0; i < 31; i++) z_tilde[i] = 1 - 2 * x[i];
  for (int i = 0; i < 31; i++) {
    c0[i] = c_tilde[(i + n_id_2) % 31];
    c1[i] = c_tilde[(i + n_id_2 + 3) % 31];
  }
  memset(n_id_1_table, 0, sizeof(int) * 900);       /* bzero'd struct: unassigned cells read 0 */
  for (int nid = 0; nid < 168; nid++) {
    int qp = nid / 30;
    int q = (nid + qp * (qp + 1) / 2) / 30;
    int mp = nid + q * (q + 1) / 2;
    int m0 = mp % 31;
    int m1 = (m0 + mp / 31 + 1) % 31;
    n_id_1_table[m0 * 30 + (m1 - 1)] = nid;
  }
}

/* srslte_cexptab_init(4096): tab[i] = cexpf(j 2 pi i / size); one spare entry (the
 * reference mallocs size+1 and leaves it unset; defined here as tab[0]) [A.3] */
void orc_cexptab(float tab_re[4097], float tab_im[4097])
{
  for (int i = 0; i < 4096; i++) {
    double a = 2.0 * M_PI * (double)i / 4096.0;
    tab_re[i] = (float)cos(a);
    tab_im[i] = (float)sin(a);
  }
  tab_re[4096] = tab_re[0];
  tab_im[4096] = tab_im[0];
}

void orc_fft128_twiddles(float w_re[64], float w_im[64])
{
  for (int k = 0; k < 64; k++) {
    w_re[k] = (float)cos128(k);
    w_im[k] = (float)(-sin128(k));
  }
}

/* ------------------------------------------------------------------------- */
/* Front end                                                                  */
/* ------------------------------------------------------------------------- */

void orc_sc16_to_fc32(const int16_t *iq, int64_t n, float scale, orc_cf *out)
{
  for (int64_t i = 0; i < n; i++) {
    out[i].re = (float)iq[2 * i] * scale;
    out[i].im = (float)iq[2 * i + 1] * scale;
  }
}

void orc_sc8_to_fc32(const int8_t *iq, int64_t n, float scale, orc_cf *out)
{
  for (int64_t i = 0; i < n; i++) {
    out[i].re = (float)iq[2 * i] * scale;
    out[i].im = (float)iq[2 * i + 1] * scale;
  }
}

/* rational_resampler_ccc(1, D): y[k] = sum_j taps[j] x[kD - j], zero history [A.7].
 * Canonical order: the polyphase branches v = j mod D are addressed by their position
 * p = (D - v) % D inside an aligned block of D input samples.  Each position accumulates
 *   P[p] = fma(taps[qD+v], x[kD-qD-v], P[p])   over q = 0..32 ascending
 * (taps beyond ntaps are zeros) in one chain per component, and the D partials are summed by
 * the balanced pairwise tree  ((P0+P1)+(P2+P3)) + ((P4+P5)+(P6+P7)) ...; when D is not a power
 * of two the tree is padded with zero partials up to the next one.  D = 1..64: the reference
 * resamples by any integer ratio (examples/cell_search_file.py:50-57). */
int64_t orc_decimate(const orc_cf *x, int64_t n_in, int decim, orc_cf *y)
{
  if (decim <= 1) { memcpy(y, x, sizeof(orc_cf) * n_in); return n_in; }
  if (decim > 64) return -1;
  float taps[4096];
  int ntaps = orc_decim_taps(decim, taps, 4096);
  const int Q = 33;                       /* ceil(ntaps / D) <= 33 for every D */
  if (ntaps <= 0 || ntaps > Q * decim) return -1;
  int P2 = 1;
  while (P2 < decim) P2 <<= 1;
  int64_t n_out = (n_in + decim - 1) / decim;
  for (int64_t k = 0; k < n_out; k++) {
    float pr[64], pi[64];
    for (int p = decim; p < P2; p++) { pr[p] = 0.f; pi[p] = 0.f; }
    for (int p = 0; p < decim; p++) {
      const int v = (decim - p) % decim;
      float ar = 0.f, ai = 0.f;
      for (int q = 0; q < Q; q++) {
        int j = q * decim + v;
        float t = j < ntaps ? taps[j] : 0.f;
        int64_t idx = k * decim - j;
        float xr = 0.f, xi = 0.f;
        if (idx >= 0) { xr = x[idx].re; xi = x[idx].im; }
        ar = fmaf(t, xr, ar);
        ai = fmaf(t, xi, ai);
      }
      pr[p] = ar; pi[p] = ai;
    }
    for (int w = 1; w < P2; w <<= 1)
      for (int p = 0; p < P2; p += 2 * w) { pr[p] = pr[p] + pr[p + w]; pi[p] = pi[p] + pi[p + w]; }
    y[k].re = pr[0]; y[k].im = pi[0];
  }
  return n_out;
}

/* The same filter the way a CPU implementation evaluates it (gr::filter::fir_filter_ccf /
 * volk dot product: one output at a time over contiguous taps, SIMD lanes as independent
 * accumulators).  Used for the timed CPU baseline only (ORC_CONV_FFT, the reference-class mode):
 * the canonical order above is scalar by construction and would charge the reference path for
 * something its own resampler does not do.  Same taps, rounding differs in the last bits. */
int64_t orc_decimate_fast(const orc_cf *x, int64_t n_in, int decim, orc_cf *y)
{
  if (decim <= 1) { memcpy(y, x, sizeof(orc_cf) * n_in); return n_in; }
  if (decim > 64) return -1;
  float taps[4096];
  int ntaps = orc_decim_taps(decim, taps, 4096);
  if (ntaps <= 0) return -1;
  /* reversed taps, each twice (re, im lanes), zero padded to a multiple of 32 floats */
  const int nt2 = ((2 * ntaps + 31) / 32) * 32;
  float *t2 = calloc(nt2 + 32, sizeof(float));
  for (int j = 0; j < ntaps; j++) { t2[2 * j] = taps[ntaps - 1 - j]; t2[2 * j + 1] = taps[ntaps - 1 - j]; }
  /* input with ntaps - 1 zeros in front (zero initial state) and padding behind, as floats */
  const int64_t lead = ntaps - 1;
  float *xf = calloc(2 * (size_t)(lead + n_in) + nt2 + 32, sizeof(float));
  memcpy(xf + 2 * lead, x, sizeof(orc_cf) * n_in);
  int64_t n_out = (n_in + decim - 1) / decim;
  for (int64_t k = 0; k < n_out; k++) {
    const float *w = xf + 2 * (k * decim);                /* x[kD - (ntaps-1)] .. x[kD] */
    float acc[32];                                        /* four 8-float vectors in flight */
    for (int l = 0; l < 32; l++) acc[l] = 0.f;
    for (int i = 0; i < nt2; i += 32)
      for (int l = 0; l < 32; l++) acc[l] = fmaf(t2[i + l], w[i + l], acc[l]);
    float re = 0.f, im = 0.f;
    for (int l = 0; l < 32; l += 2) { re += acc[l]; im += acc[l + 1]; }
    y[k].re = re; y[k].im = im;
  }
  free(t2); free(xf);
  return n_out;
}

/* ORC_FRONT_TCINT: the arithmetic of the product's integer tensor-core front end (LTB_FRONTEND_TC_INT,
 * gr-ltetrigger_b200/csrc/ltb_tc_frontend.cuh), restated with 64-bit integers.  The default taps of
 * rational_resampler_ccc(1, 16) are quantised once, T[j] = rint(taps[j] * 2^27) (|T| < 2^23: three balanced
 * base-256 digits on the tensor core); the int16 samples are used as they are; per component
 *   A[k] = sum_j T[j] x[16 k - j]          exact, |A| < 2^44
 *   y[k] = (float)A[k] * 2^-42             one rounding (2^-27 for the taps, 2^-15 for the sc16 scale)
 * Exact arithmetic does not depend on evaluation order, so any correct integer evaluation -- this loop, or
 * int8 digit products accumulated in int32 and recombined -- gives the same bits. */
static int64_t tcint_run(const int16_t *iq16, const int8_t *iq8, int64_t n_in, int decim, orc_cf *y)
{
  float taps[4096];
  int ntaps = orc_decim_taps(decim, taps, 4096);
  if (ntaps <= 0 || ntaps > 2112) return -1;
  int32_t T[2112];
  const int shift = tcint_shift(decim);
  for (int j = 0; j < ntaps; j++) T[j] = (int32_t)llrint(ldexp((double)taps[j], shift));
  const float scale = (float)ldexp(1.0, -(shift + (iq16 ? 15 : 7)));
  int64_t n_out = (n_in + decim - 1) / decim;
  for (int64_t k = 0; k < n_out; k++) {
    int64_t are = 0, aim = 0;
    const int64_t top = k * decim;
    const int jmax = top < ntaps - 1 ? (int)top : ntaps - 1;      /* zero history before the stream */
    if (iq16) {
      const int16_t *xp = iq16 + 2 * top;
      for (int j = 0; j <= jmax; j++) { are += (int64_t)T[j] * xp[-2 * j]; aim += (int64_t)T[j] * xp[-2 * j + 1]; }
    } else {
      const int8_t *xp = iq8 + 2 * top;
      for (int j = 0; j <= jmax; j++) { are += (int64_t)T[j] * xp[-2 * j]; aim += (int64_t)T[j] * xp[-2 * j + 1]; }
    }
    y[k].re = (float)are * scale;
    y[k].im = (float)aim * scale;
  }
  return n_out;
}

int64_t orc_decimate_tcint_sc16(const int16_t *iq, int64_t n_in, orc_cf *y) { return tcint_run(iq, NULL, n_in, 16, y); }
/* other rates: taps scaled by 2^(23 + floor(log2 decim)) (|T| < 2^23 at every rate), y = (float)A * 2^-(shift + 15) */
int64_t orc_decimate_tcint_sc16_d(const int16_t *iq, int64_t n_in, int decim, orc_cf *y) { return tcint_run(iq, NULL, n_in, decim, y); }

/* the same for int8 I/Q: y[k] = (float)A[k] * 2^-34 (2^-27 for the taps, 2^-7 for the sc8 scale) */
int64_t orc_decimate_tcint_sc8(const int8_t *iq, int64_t n_in, orc_cf *y) { return tcint_run(NULL, iq, n_in, 16, y); }
int64_t orc_decimate_tcint_sc8_d(const int8_t *iq, int64_t n_in, int decim, orc_cf *y) { return tcint_run(NULL, iq, n_in, decim, y); }

/* ------------------------------------------------------------------------- */
/* PSS matched filter                                                         */
/* ------------------------------------------------------------------------- */

typedef struct {
  float hr[3][128], hi[3][128];       /* canonical float taps per N_id_2 */
  float tab_re[4097], tab_im[4097];   /* cexptab */
  float w_re[64], w_im[64];           /* FFT128 twiddles */
  /* FFT-mode tables */
  float *tw512_re, *tw512_im;         /* 256 twiddles of the 512-point stage */
  float *twN_re, *twN_im;             /* W_9728^(n1*k2), [k2][n1] */
  float w19_re[19], w19_im[19];
  float *H_re[3], *H_im[3];           /* FFT_9728(h_pad) per root, natural order, with 1/N folded in */
  /* overlap-save mode tables */
  float tw1024_re[1024], tw1024_im[1024];   /* W_1024^i */
  float os_H_re[3][1024], os_H_im[3][1024]; /* DFT_1024(h zero padded) * 2^-10, natural order */
  int ready;
} tables_t;

static tables_t T;

static void fft9728(const float *xr, const float *xi, float *Xr, float *Xi, int inverse, float *wk);

static pthread_once_t tables_once = PTHREAD_ONCE_INIT;

static void tables_build(void)
{
  {
    {
      for (int r = 0; r < 3; r++) orc_pss_taps(r, T.hr[r], T.hi[r]);
      orc_cexptab(T.tab_re, T.tab_im);
      orc_fft128_twiddles(T.w_re, T.w_im);
      T.tw512_re = malloc(sizeof(float) * 256); T.tw512_im = malloc(sizeof(float) * 256);
      for (int k = 0; k < 256; k++) {
        double a = -2.0 * M_PI * k / 512.0;
        T.tw512_re[k] = (float)cos(a); T.tw512_im[k] = (float)sin(a);
      }
      T.twN_re = malloc(sizeof(float) * 512 * 19); T.twN_im = malloc(sizeof(float) * 512 * 19);
      for (int k2 = 0; k2 < 512; k2++)
        for (int n1 = 0; n1 < 19; n1++) {
          double a = -2.0 * M_PI * (double)((long)n1 * k2) / (double)FFTN;
          T.twN_re[k2 * 19 + n1] = (float)cos(a); T.twN_im[k2 * 19 + n1] = (float)sin(a);
        }
      for (int k = 0; k < 19; k++) {
        double a = -2.0 * M_PI * k / 19.0;
        T.w19_re[k] = (float)cos(a); T.w19_im[k] = (float)sin(a);
      }
      float *pr = calloc(FFTN, sizeof(float)), *pi = calloc(FFTN, sizeof(float));
      float *wk = malloc(sizeof(float) * FFTN * 4);
      for (int r = 0; r < 3; r++) {
        memset(pr, 0, sizeof(float) * FFTN); memset(pi, 0, sizeof(float) * FFTN);
        memcpy(pr, T.hr[r], sizeof(float) * 128); memcpy(pi, T.hi[r], sizeof(float) * 128);
        T.H_re[r] = malloc(sizeof(float) * FFTN); T.H_im[r] = malloc(sizeof(float) * FFTN);
        fft9728(pr, pi, T.H_re[r], T.H_im[r], 0, wk);
        for (int k = 0; k < FFTN; k++) { T.H_re[r][k] /= (float)FFTN; T.H_im[r][k] /= (float)FFTN; }
      }
      free(pr); free(pi); free(wk);
      orc_fft1024_twiddles(T.tw1024_re, T.tw1024_im);
      for (int r = 0; r < 3; r++) orc_os_filter(r, T.os_H_re[r], T.os_H_im[r]);
      T.ready = 1;
    }
  }
}

static void tables_init(void) { pthread_once(&tables_once, tables_build); }

/* minimal pthread parallel-for: jobs are claimed with an atomic counter */
typedef struct { void (*fn)(int, void *); void *ctx; int n; int next; } pfor_t;
static void *pfor_worker(void *arg)
{
  pfor_t *p = (pfor_t *)arg;
  for (;;) {
    int j = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
    if (j >= p->n) break;
    p->fn(j, p->ctx);
  }
  return NULL;
}
static void parallel_for(int n, int nthreads, void (*fn)(int, void *), void *ctx)
{
  if (nthreads <= 0) { long c = sysconf(_SC_NPROCESSORS_ONLN); nthreads = c > 0 ? (int)c : 1; }
  if (nthreads > n) nthreads = n;
  pfor_t p = { fn, ctx, n, 0 };
  if (nthreads <= 1) { pfor_worker(&p); return; }
  pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
  for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, pfor_worker, &p);
  for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  free(th);
}

/* Canonical direct form for `cnt` consecutive lags starting at sample index n0 of the
 * planar signal (xr, xi); xr[n0-128 .. n0+cnt) must be readable (zeros where the
 * window is truncated).  Folded taps m = 0..64 [DESIGN.md "Canonical arithmetic"]:
 *   s_0 = x[n], s_m = x[n-m] + x[n-128+m] (1<=m<=63), s_64 = x[n-64]
 *   A += hr[m]*s.re   B += hi[m]*s.im   C += hr[m]*s.im   D += hi[m]*s.re   (fmaf)
 *   N_id_2 0,1: y = (A-B, C+D);  N_id_2 2 (conjugate taps of 1): y = (A+B, C-D)
 *   power = fmaf(y.re, y.re, y.im*y.im)                                           */
#define TILE 256
static void corr_direct(const float *xr, const float *xi, int64_t n0, int64_t cnt, int n_id_2, float *power)
{
  const int g = (n_id_2 == 0) ? 0 : 1;
  const float *hr = T.hr[g], *hi = T.hi[g];
  for (int64_t t0 = 0; t0 < cnt; t0 += TILE) {
    int len = (int)((cnt - t0 < TILE) ? cnt - t0 : TILE);
    float A[TILE], B[TILE], C[TILE], D[TILE];
    for (int k = 0; k < len; k++) A[k] = B[k] = C[k] = D[k] = 0.f;
    const float *pr = xr + n0 + t0, *pi = xi + n0 + t0;
    for (int m = 0; m <= 64; m++) {
      const float cr = hr[m], ci = hi[m];
      if (m == 0 || m == 64) {
        for (int k = 0; k < len; k++) {
          float sr = pr[k - m], si = pi[k - m];
          A[k] = fmaf(cr, sr, A[k]); B[k] = fmaf(ci, si, B[k]);
          C[k] = fmaf(cr, si, C[k]); D[k] = fmaf(ci, sr, D[k]);
        }
      } else {
        for (int k = 0; k < len; k++) {
          float sr = pr[k - m] + pr[k - 128 + m], si = pi[k - m] + pi[k - 128 + m];
          A[k] = fmaf(cr, sr, A[k]); B[k] = fmaf(ci, si, B[k]);
          C[k] = fmaf(cr, si, C[k]); D[k] = fmaf(ci, sr, D[k]);
        }
      }
    }
    if (n_id_2 == 2) {
      for (int k = 0; k < len; k++) { float re = A[k] + B[k], im = C[k] - D[k]; power[t0 + k] = fmaf(re, re, im * im); }
    } else {
      for (int k = 0; k < len; k++) { float re = A[k] - B[k], im = C[k] + D[k]; power[t0 + k] = fmaf(re, re, im * im); }
    }
  }
}

/* ---- 9728-point FFT (19 x 512), float32, for the reference-class convolution ---- */
/* data layout: element n = 19*n2 + n1 viewed as [n2][n1]; radix-2 DIF over n2 on rows of
 * 19 contiguous values, per-row twiddle W_N^(n1*k2), then 19-point DFTs. */
static unsigned bitrev9(unsigned v) { unsigned r = 0; for (int i = 0; i < 9; i++) { r = (r << 1) | (v & 1); v >>= 1; } return r; }

static void fft9728(const float *xr, const float *xi, float *Xr, float *Xi, int inverse, float *wk)
{
  float *ar = wk, *ai = wk + FFTN;           /* [512][19] */
  float *zr = wk + 2 * FFTN, *zi = wk + 3 * FFTN;
  memcpy(ar, xr, sizeof(float) * FFTN);
  if (inverse) for (int i = 0; i < FFTN; i++) ai[i] = -xi[i]; else memcpy(ai, xi, sizeof(float) * FFTN);
  /* radix-2 decimation in frequency over the 512 axis; output rows in bit-reversed order */
  for (int half = 256; half >= 1; half >>= 1) {
    int step = 256 / half;
    for (int base = 0; base < 512; base += 2 * half) {
      for (int j = 0; j < half; j++) {
        float wr = T.tw512_re[j * step], wi = T.tw512_im[j * step];
        float *ur = ar + (base + j) * 19, *ui = ai + (base + j) * 19;
        float *vr = ar + (base + j + half) * 19, *vi = ai + (base + j + half) * 19;
        for (int c = 0; c < 19; c++) {
          float sr = ur[c] + vr[c], si = ui[c] + vi[c];
          float dr = ur[c] - vr[c], di = ui[c] - vi[c];
          ur[c] = sr; ui[c] = si;
          vr[c] = dr * wr - di * wi; vi[c] = dr * wi + di * wr;
        }
      }
    }
  }
  /* twiddle by W_N^(n1*k2), k2 = bitrev(row) */
  for (int row = 0; row < 512; row++) {
    int k2 = (int)bitrev9((unsigned)row);
    const float *tr = T.twN_re + k2 * 19, *ti = T.twN_im + k2 * 19;
    float *pr = ar + row * 19, *pi = ai + row * 19;
    float *qr = zr + k2 * 19, *qi = zi + k2 * 19;
    for (int c = 0; c < 19; c++) {
      qr[c] = pr[c] * tr[c] - pi[c] * ti[c];
      qi[c] = pr[c] * ti[c] + pi[c] * tr[c];
    }
  }
  /* 19-point DFTs: X[k2 + 512*k1] = sum_n1 W19^(n1*k1) z[k2][n1] */
  for (int k1 = 0; k1 < 19; k1++) {
    float cr[19], ci[19];
    for (int n1 = 0; n1 < 19; n1++) { int e = (n1 * k1) % 19; cr[n1] = T.w19_re[e]; ci[n1] = T.w19_im[e]; }
    float *outr = Xr + 512 * k1, *outi = Xi + 512 * k1;
    for (int k2 = 0; k2 < 512; k2++) {
      const float *qr = zr + k2 * 19, *qi = zi + k2 * 19;
      float sr = 0.f, si = 0.f;
      for (int n1 = 0; n1 < 19; n1++) {
        sr += qr[n1] * cr[n1] - qi[n1] * ci[n1];
        si += qr[n1] * ci[n1] + qi[n1] * cr[n1];
      }
      outr[k2] = sr; outi[k2] = inverse ? -si : si;
    }
  }
}

typedef struct { float *buf; } fftwork_t;

/* Reference-class evaluation [A.2 step 1]: zero-padded 9728-point FFT convolution. */
static void corr_fft(const orc_cf *win, int n_id_2, float *power, float *wk /* 10*FFTN floats */)
{
  float *xr = wk, *xi = wk + FFTN, *Xr = wk + 2 * FFTN, *Xi = wk + 3 * FFTN;
  float *yr = wk + 4 * FFTN, *yi = wk + 5 * FFTN, *w2 = wk + 6 * FFTN;
  for (int i = 0; i < ORC_HALF; i++) { xr[i] = win[i].re; xi[i] = win[i].im; }
  memset(xr + ORC_HALF, 0, sizeof(float) * (FFTN - ORC_HALF));
  memset(xi + ORC_HALF, 0, sizeof(float) * (FFTN - ORC_HALF));
  fft9728(xr, xi, Xr, Xi, 0, w2);
  const float *Hr = T.H_re[n_id_2], *Hi = T.H_im[n_id_2];
  for (int k = 0; k < FFTN; k++) {
    float pr = Xr[k] * Hr[k] - Xi[k] * Hi[k], pi = Xr[k] * Hi[k] + Xi[k] * Hr[k];
    Xr[k] = pr; Xi[k] = pi;
  }
  fft9728(Xr, Xi, yr, yi, 1, w2);
  for (int k = 0; k < NLAG; k++) power[k] = fmaf(yr[k], yr[k], yi[k] * yi[k]);
}

void orc_pss_corr_window(const orc_cf *win, int n_id_2, int conv_mode, float *power)
{
  tables_init();
  if (conv_mode == ORC_CONV_FFT) {
    float *wk = malloc(sizeof(float) * FFTN * 10);
    corr_fft(win, n_id_2, power, wk);
    free(wk);
    return;
  }
  const int PADN = 128 + ORC_HALF + 128;
  float *xr = calloc(PADN, sizeof(float)), *xi = calloc(PADN, sizeof(float));
  for (int i = 0; i < ORC_HALF; i++) { xr[128 + i] = win[i].re; xi[128 + i] = win[i].im; }
  corr_direct(xr + 128, xi + 128, 0, NLAG, n_id_2, power);
  free(xr); free(xi);
}

void orc_pss_corr_stream(const orc_cf *x, int64_t n, int n_id_2, float *power)
{
  tables_init();
  float *xr = calloc(n + 128, sizeof(float)), *xi = calloc(n + 128, sizeof(float));
  for (int64_t i = 0; i < n; i++) { xr[128 + i] = x[i].re; xi[128 + i] = x[i].im; }
  corr_direct(xr + 128, xi + 128, 0, n, n_id_2, power);
  free(xr); free(xi);
}

/* ------------------------------------------------------------------------- */
/* Overlap-save mode (ORC_CONV_OS): the matched filter as 1024-point FFT blocks  */
/* aligned to absolute sample indices.  This is the canonical arithmetic of the  */
/* GPU's pss_corr_fft_kernel, restated here operation for operation.             */
/*                                                                               */
/* Block b covers outputs n in [896 b, 896 b + 896) from inputs                   */
/* x[896 b - 128 .. 896 b + 896): X = FFT(x_blk), Y_g = H_g . X, y_g = IFFT(Y_g),  */
/* P_g[896 b - 128 + n] = |y_g[n]|^2 for n in [128, 1024).                         */
/* FFT_1024 is a four-step transform, n = 32 n1 + n2, f = k1 + 32 k2:              */
/*   A[k1][n2] = DFT32_{n1} x[32 n1 + n2]        radix-2 DIF, W_32 literals         */
/*   B[k1][n2] = W_1024^(n2 k1) A[k1][n2]        table, canonical complex product   */
/*   X[k1 + 32 k2] = DFT32_{n2} B[k1][n2]        radix-2 DIF                        */
/* and the inverse runs the mirrored graph (DIT over k2, conjugate table twiddle,   */
/* DIF over k1) with conjugated twiddles; 2^-10 is folded into H.  Butterflies:     */
/*   DIF: a' = a + b, b' = w (a - b)      DIT: t = w b, a' = a + t, b' = a - t        */
/* with w b = (fma(wr, br, -(wi*bi)), fma(wr, bi, wi*br)); w = 1 and w = -+j are      */
/* applied exactly (copy / swap with a sign).                                         */
/* ------------------------------------------------------------------------- */
static const float W32[16][2] = {
  {1.000000000e+00f, 0.000000000e+00f},   {9.807852507e-01f, -1.950903237e-01f},
  {9.238795042e-01f, -3.826834261e-01f},  {8.314695954e-01f, -5.555702448e-01f},
  {7.071067691e-01f, -7.071067691e-01f},  {5.555702448e-01f, -8.314695954e-01f},
  {3.826834261e-01f, -9.238795042e-01f},  {1.950903237e-01f, -9.807852507e-01f},
  {0.000000000e+00f, -1.000000000e+00f},  {-1.950903237e-01f, -9.807852507e-01f},
  {-3.826834261e-01f, -9.238795042e-01f}, {-5.555702448e-01f, -8.314695954e-01f},
  {-7.071067691e-01f, -7.071067691e-01f}, {-8.314695954e-01f, -5.555702448e-01f},
  {-9.238795042e-01f, -3.826834261e-01f}, {-9.807852507e-01f, -1.950903237e-01f}};

void orc_fft1024_twiddles(float w_re[1024], float w_im[1024])
{
  for (int i = 0; i < 1024; i++) {
    double a = 2.0 * M_PI * (double)i / 1024.0;
    w_re[i] = (float)cos(a); w_im[i] = (float)(-sin(a));
  }
  w_re[0] = 1.f;    w_im[0] = 0.f;   w_re[256] = 0.f; w_im[256] = -1.f;
  w_re[512] = -1.f; w_im[512] = 0.f; w_re[768] = 0.f; w_im[768] = 1.f;
}

/* H_g[f] = 2^-10 sum_m h_g[m] W_1024^(f m), double accumulation in ascending m */
void orc_os_filter(int n_id_2, float H_re[1024], float H_im[1024])
{
  float hr[128], hi[128];
  orc_pss_taps(n_id_2, hr, hi);
  for (int f = 0; f < 1024; f++) {
    double ar = 0.0, ai = 0.0;
    for (int m = 0; m < 128; m++) {
      double a = 2.0 * M_PI * (double)((f * m) & 1023) / 1024.0;
      double c = cos(a), sn = -sin(a);
      ar += (double)hr[m] * c - (double)hi[m] * sn;
      ai += (double)hr[m] * sn + (double)hi[m] * c;
    }
    H_re[f] = (float)(ar / 1024.0); H_im[f] = (float)(ai / 1024.0);
  }
}

static inline void cmul_w(float wr, float wi, float br, float bi, float *re, float *im)
{
  *re = fmaf(wr, br, -(wi * bi));
  *im = fmaf(wr, bi, wi * br);
}

/* twiddle W_32^idx (conj if inv) times (dr, di), exact for idx 0 and 8 */
static inline void tw32(int idx, int inv, float dr, float di, float *re, float *im)
{
  if (idx == 0) { *re = dr; *im = di; }
  else if (idx == 8) { if (inv) { *re = -di; *im = dr; } else { *re = di; *im = -dr; } }
  else cmul_w(W32[idx][0], inv ? -W32[idx][1] : W32[idx][1], dr, di, re, im);
}

/* radix-2 DIF, natural order in, bit-reversed order out */
static void fft32_dif(float *vr, float *vi, int inv)
{
  for (int s = 0; s < 5; s++) {
    int half = 16 >> s;
    for (int base = 0; base < 32; base += 2 * half)
      for (int k = 0; k < half; k++) {
        int i = base + k, j = i + half;
        float ar = vr[i], ai = vi[i], br = vr[j], bi = vi[j];
        vr[i] = ar + br; vi[i] = ai + bi;
        tw32(k << s, inv, ar - br, ai - bi, &vr[j], &vi[j]);
      }
  }
}

/* radix-2 DIT, bit-reversed order in, natural order out */
static void fft32_dit(float *vr, float *vi, int inv)
{
  for (int s = 0; s < 5; s++) {
    int half = 1 << s;
    for (int base = 0; base < 32; base += 2 * half)
      for (int k = 0; k < half; k++) {
        int i = base + k, j = i + half;
        float tr, ti;
        tw32(k * (16 >> s), inv, vr[j], vi[j], &tr, &ti);
        float ar = vr[i], ai = vi[i];
        vr[i] = ar + tr; vi[i] = ai + ti;
        vr[j] = ar - tr; vi[j] = ai - ti;
      }
  }
}

static int bitrev5(int v) { return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4); }

/* forward: x natural (n = 32 n1 + n2) -> X stored at [r][k1] with f = k1 + 32 bitrev5(r) */
static void fft1024_fwd(const float *xr, const float *xi, float *Xr, float *Xi)
{
  static __thread float Br[32][32], Bi[32][32];     /* [k1][n2] */
  float vr[32], vi[32];
  for (int n2 = 0; n2 < 32; n2++) {
    for (int n1 = 0; n1 < 32; n1++) { vr[n1] = xr[32 * n1 + n2]; vi[n1] = xi[32 * n1 + n2]; }
    fft32_dif(vr, vi, 0);
    for (int r = 0; r < 32; r++) {
      int k1 = bitrev5(r), t = (n2 * k1) & 1023;
      cmul_w(T.tw1024_re[t], T.tw1024_im[t], vr[r], vi[r], &Br[k1][n2], &Bi[k1][n2]);
    }
  }
  for (int k1 = 0; k1 < 32; k1++) {
    for (int n2 = 0; n2 < 32; n2++) { vr[n2] = Br[k1][n2]; vi[n2] = Bi[k1][n2]; }
    fft32_dif(vr, vi, 0);
    for (int r = 0; r < 32; r++) { Xr[r * 32 + k1] = vr[r]; Xi[r * 32 + k1] = vi[r]; }
  }
}

/* inverse (unscaled): Y at [r][k1] as above -> y natural */
static void fft1024_inv(const float *Yr, const float *Yi, float *yr, float *yi)
{
  static __thread float Dr[32][32], Di[32][32];     /* [n2][k1] */
  float vr[32], vi[32];
  for (int k1 = 0; k1 < 32; k1++) {
    for (int r = 0; r < 32; r++) { vr[r] = Yr[r * 32 + k1]; vi[r] = Yi[r * 32 + k1]; }
    fft32_dit(vr, vi, 1);                               /* over k2 -> n2 natural */
    for (int n2 = 0; n2 < 32; n2++) {
      int t = (k1 * n2) & 1023;
      cmul_w(T.tw1024_re[t], -T.tw1024_im[t], vr[n2], vi[n2], &Dr[n2][k1], &Di[n2][k1]);
    }
  }
  for (int n2 = 0; n2 < 32; n2++) {
    for (int k1 = 0; k1 < 32; k1++) { vr[k1] = Dr[n2][k1]; vi[k1] = Di[n2][k1]; }
    fft32_dif(vr, vi, 1);                               /* over k1 -> n1 bit-reversed */
    for (int r = 0; r < 32; r++) { yr[32 * bitrev5(r) + n2] = vr[r]; yi[32 * bitrev5(r) + n2] = vi[r]; }
  }
}

void orc_fft1024(const orc_cf *in, orc_cf *out, int inverse)
{
  tables_init();
  float xr[1024], xi[1024], Xr[1024], Xi[1024];
  if (!inverse) {
    for (int i = 0; i < 1024; i++) { xr[i] = in[i].re; xi[i] = in[i].im; }
    fft1024_fwd(xr, xi, Xr, Xi);
    for (int r = 0; r < 32; r++)
      for (int k1 = 0; k1 < 32; k1++) { out[k1 + 32 * bitrev5(r)].re = Xr[r * 32 + k1]; out[k1 + 32 * bitrev5(r)].im = Xi[r * 32 + k1]; }
  } else {
    for (int r = 0; r < 32; r++)
      for (int k1 = 0; k1 < 32; k1++) { Xr[r * 32 + k1] = in[k1 + 32 * bitrev5(r)].re; Xi[r * 32 + k1] = in[k1 + 32 * bitrev5(r)].im; }
    fft1024_inv(Xr, Xi, xr, xi);
    for (int i = 0; i < 1024; i++) { out[i].re = xr[i]; out[i].im = xi[i]; }
  }
}

/* power of all three roots for the whole blocks of x[0..n): out[g][i], i < 896 * (n / 896);
 * x[<0] = 0.  Returns the number of outputs per root. */
int64_t orc_pss_corr_os(const orc_cf *x, int64_t n, float *p0, float *p1, float *p2)
{
  tables_init();
  float *pw[3] = {p0, p1, p2};
  float xr[1024], xi[1024], Xr[1024], Xi[1024], Yr[1024], Yi[1024], yr[1024], yi[1024];
  int64_t nblk = n / ORC_OS_STEP;
  for (int64_t b = 0; b < nblk; b++) {
    int64_t base = b * ORC_OS_STEP - 128;
    for (int i = 0; i < 1024; i++) {
      int64_t idx = base + i;
      xr[i] = idx >= 0 ? x[idx].re : 0.f; xi[i] = idx >= 0 ? x[idx].im : 0.f;
    }
    fft1024_fwd(xr, xi, Xr, Xi);
    for (int g = 0; g < 3; g++) {
      if (!pw[g]) continue;
      for (int r = 0; r < 32; r++)
        for (int k1 = 0; k1 < 32; k1++) {
          int f = k1 + 32 * bitrev5(r);
          cmul_w(T.os_H_re[g][f], T.os_H_im[g][f], Xr[r * 32 + k1], Xi[r * 32 + k1], &Yr[r * 32 + k1], &Yi[r * 32 + k1]);
        }
      fft1024_inv(Yr, Yi, yr, yi);
      for (int i = 128; i < 1024; i++) pw[g][base + i] = fmaf(yr[i], yr[i], yi[i] * yi[i]);
    }
  }
  return nblk * ORC_OS_STEP;
}

/* ------------------------------------------------------------------------- */
/* srslte_pss_find_pss [A.2]                                                  */
/* ------------------------------------------------------------------------- */

typedef struct {
  float avg[AVG_LEN];          /* conv_output_avg; entries >= 9726 stay 0 */
  float peak_value;
  float *xr, *xi;              /* padded planes for the direct form */
  float *power;
  float *fftwk;
  int conv_mode, n_id_2;
  const float *os_power;       /* ORC_CONV_OS: whole-stream block powers, os_power[k] = lag k of this window */
} pss_core_t;

static int vec_max_fi(const float *x, int len)
{
  float m = -3.402823466e+38f; int p = 0;
  for (int i = 0; i < len; i++) if (x[i] > m) { m = x[i]; p = i; }
  return p;
}

static int find_pss(pss_core_t *q, const orc_cf *in, float *psr_out)
{
  if (q->conv_mode == ORC_CONV_FFT) {
    corr_fft(in, q->n_id_2, q->power, q->fftwk);
  } else if (q->conv_mode == ORC_CONV_OS) {
    /* lags that see the window's zero padding (k < 127, k >= 9600): direct form on the window;
     * interior lags: the stream's overlap-save block values */
    for (int i = 0; i < ORC_HALF; i++) { q->xr[128 + i] = in[i].re; q->xi[128 + i] = in[i].im; }
    corr_direct(q->xr + 128, q->xi + 128, 0, 127, q->n_id_2, q->power);
    corr_direct(q->xr + 128, q->xi + 128, ORC_HALF, NLAG - ORC_HALF, q->n_id_2, q->power + ORC_HALF);
    memcpy(q->power + 127, q->os_power + 127, sizeof(float) * (ORC_HALF - 127));
  } else {
    for (int i = 0; i < ORC_HALF; i++) { q->xr[128 + i] = in[i].re; q->xi[128 + i] = in[i].im; }
    corr_direct(q->xr + 128, q->xi + 128, 0, NLAG, q->n_id_2, q->power);
  }
  const float alpha = PSS_EMA_ALPHA, beta = 1 - PSS_EMA_ALPHA;
  float *avg = q->avg;
  for (int k = 0; k < NLAG; k++) {
    float a = q->power[k] * alpha;
    float b = avg[k] * beta;
    avg[k] = a + b;
  }
  int p = vec_max_fi(avg, NLAG);
  q->peak_value = avg[p];
  const int conv_output_len = NLAG + 1;   /* 9727 */
  int ub = p + 1;
  while (avg[ub + 1] <= avg[ub] && ub < conv_output_len) ub++;
  int lb;
  if (p > 2) { lb = p - 1; while (avg[lb - 1] <= avg[lb] && lb > 1) lb--; }
  else lb = 0;
  int dist_r = conv_output_len - 1 - ub; if (dist_r < 0) dist_r = 0;
  int sl_right = ub + vec_max_fi(&avg[ub], dist_r);
  int sl_left = vec_max_fi(avg, lb);
  float side = avg[sl_left] > avg[sl_right] ? avg[sl_left] : avg[sl_right];
  *psr_out = avg[p] / side;
  return p;
}

/* ------------------------------------------------------------------------- */
/* Canonical atan2 (double, fixed polynomial; identical sequence on the GPU)   */
/* ------------------------------------------------------------------------- */
static double canon_atan2(double y, double x)
{
  double ax = fabs(x), ay = fabs(y);
  double mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
  if (mx == 0.0) return 0.0;
  double q = mn / mx;
  double t = q, off = 0.0;
  if (q > 0.41421356237309503) { t = (q - 1.0) / (q + 1.0); off = 0.78539816339744828; }
  double t2 = t * t;
  double p = 1.0 / 35.0;                       /* terms k = 17 .. 0 of sum (-1)^k t^(2k)/(2k+1) */
  for (int k = 16; k >= 0; k--) {
    double c = 1.0 / (double)(2 * k + 1);
    p = -p;                                    /* alternate sign: p_k = c_k - t2 * p_{k+1} */
    p = fma(p, t2, c);
  }
  double r = fma(t, p, off);
  if (ay > ax) r = 1.5707963267948966 - r;
  if (x < 0.0) r = 3.1415926535897931 - r;
  if (y < 0.0) r = -r;
  return r;
}

/* ------------------------------------------------------------------------- */
/* pss block  [ref lib/pss_impl.cc, lib/pss_impl.h]                           */
/* ------------------------------------------------------------------------- */
struct orc_pss {
  pss_core_t core;
  /* tracking_t (lib/pss_impl.h:41-50) */
  int score, timer, is_tracking;
  int tracking_lost;
  float psr_data[ORC_MAVG]; size_t psr_i;
  float psr, psr_max; int peak_pos;
  float cfo_data[ORC_MAVG]; size_t cfo_i;
  float cfo_last_freq;          /* d_cfo.last_freq */
  float cfo_table_freq;         /* frequency the cur_cexp table was generated for */
  int n_id_2; float thr; int track_after, track_every;
  int win_index;
};

orc_pss *orc_pss_new(int n_id_2, float thr, int track_after, int track_every, int conv_mode)
{
  if (n_id_2 < 0 || n_id_2 > 2) return NULL;
  tables_init();
  orc_pss *b = calloc(1, sizeof *b);
  b->n_id_2 = n_id_2; b->thr = thr; b->track_after = track_after; b->track_every = track_every;
  b->core.conv_mode = conv_mode; b->core.n_id_2 = n_id_2;
  b->core.xr = calloc(128 + ORC_HALF + 128, sizeof(float));
  b->core.xi = calloc(128 + ORC_HALF + 128, sizeof(float));
  b->core.power = calloc(AVG_LEN, sizeof(float));
  b->core.fftwk = (conv_mode == ORC_CONV_FFT) ? malloc(sizeof(float) * FFTN * 10) : NULL;
  return b;
}

void orc_pss_free(orc_pss *b)
{
  if (!b) return;
  free(b->core.xr); free(b->core.xi); free(b->core.power); free(b->core.fftwk); free(b);
}

/* lib/pss_impl.cc:94-109 */
static float moving_avg(const float *data, size_t npts)
{
  if (!npts) return 0.0f;
  double acc = 0.0;
  if (npts > ORC_MAVG) npts = ORC_MAVG;
  for (size_t i = 0; i < npts; i++) acc += data[i];
  return (float)(acc / (double)npts);
}

static void pss_reset_avg(orc_pss *b) { memset(b->core.avg, 0, sizeof b->core.avg); }   /* srslte_pss_reset */

/* lib/pss_impl.cc:111-127 */
static void incr_score(orc_pss *b)
{
  int max_score = b->track_after;
  if (b->is_tracking && b->score == max_score) return;
  b->score++;
  if (!b->is_tracking && b->score == max_score) { b->is_tracking = 1; pss_reset_avg(b); }
}

/* lib/pss_impl.cc:129-152 (the partial memsets are equivalent to full clears) */
static void reset_score(orc_pss *b)
{
  if (b->score == 0) return;
  b->score = 0; b->timer = 0; b->is_tracking = 0;
  pss_reset_avg(b);
  memset(b->psr_data, 0, sizeof b->psr_data); b->psr_i = 0;
  memset(b->cfo_data, 0, sizeof b->cfo_data); b->cfo_last_freq = 0; b->cfo_i = 0;
  b->tracking_lost = 1;
}

/* srslte_pss_cfo_compute [A.3] */
static float pss_cfo_compute(const orc_pss *b, const orc_cf *r)
{
  const float *hr = T.hr[b->n_id_2], *hi = T.hi[b->n_id_2];
  float y0r = 0, y0i = 0, y1r = 0, y1i = 0;
  for (int n = 0; n < 64; n++) {
    y0r = fmaf(hr[n], r[n].re, y0r); y0r = fmaf(-hi[n], r[n].im, y0r);
    y0i = fmaf(hr[n], r[n].im, y0i); y0i = fmaf(hi[n], r[n].re, y0i);
  }
  for (int n = 64; n < 128; n++) {
    y1r = fmaf(hr[n], r[n].re, y1r); y1r = fmaf(-hi[n], r[n].im, y1r);
    y1i = fmaf(hr[n], r[n].im, y1i); y1i = fmaf(hi[n], r[n].re, y1i);
  }
  /* conj(y0) * y1 */
  float pr = fmaf(y0r, y1r, y0i * y1i);
  float pi = fmaf(y0r, y1i, -(y0i * y1r));
  return (float)(canon_atan2((double)pi, (double)pr) / M_PI);
}

/* srslte_cfo_correct + srslte_cexptab_gen [A.3] over n samples (in place allowed) */
static void cfo_correct(orc_pss *b, const orc_cf *in, orc_cf *out, float freq, int n)
{
  if (fabsf(b->cfo_last_freq - freq) > 0.0f) { b->cfo_last_freq = freq; b->cfo_table_freq = freq; }
  float phase_inc = b->cfo_table_freq * 4096.0f;
  float phase = 0.f;
  for (int i = 0; i < n; i++) {
    while (phase >= 4096.0f) phase -= 4096.0f;
    while (phase < 0.f) phase += 4096.0f;
    unsigned idx = (unsigned)phase;
    float cr = T.tab_re[idx], ci = T.tab_im[idx];
    float xr = in[i].re, xi = in[i].im;
    out[i].re = fmaf(cr, xr, -(ci * xi));
    out[i].im = fmaf(cr, xi, ci * xr);
    phase += phase_inc;
  }
}

/* lib/pss_impl.cc:154-223 */
int orc_pss_work(orc_pss *b, const orc_cf *in, orc_cf *out, int *nconsume, orc_rec *rec)
{
  uint32_t flags = 0;
  if (!b->is_tracking || b->timer == 0) {
    b->timer = b->track_every;
    b->peak_pos = find_pss(&b->core, in, &b->psr);
    b->psr_data[b->psr_i++ % ORC_MAVG] = b->psr;
    flags |= ORC_F_SEARCHED;
  } else {
    b->timer--;
  }
  int over = b->psr > b->thr;
  if (over) { incr_score(b); flags |= ORC_F_OVER; } else reset_score(b);
  if (b->psr > b->psr_max) b->psr_max = b->psr;

  int noutput = 0;
  float cfo = 0.f, mcfo = 0.f;
  int peak_used = b->peak_pos;
  int frame_start = 0;
  if (over || b->tracking_lost) {
    frame_start = b->peak_pos - ORC_SLOT;
    b->peak_pos = ORC_SLOT;
    noutput = ORC_HALF;
    *nconsume = frame_start + noutput;
    memcpy(out, in + frame_start, sizeof(orc_cf) * ORC_HALF);
    flags |= ORC_F_EMIT;
    if (b->is_tracking) {
      flags |= ORC_F_TRACKING;
      cfo = pss_cfo_compute(b, &out[ORC_SLOT - ORC_SYM]);
      b->cfo_data[b->cfo_i++ % ORC_MAVG] = cfo;
      mcfo = moving_avg(b->cfo_data, b->cfo_i);
      cfo_correct(b, out, out, -mcfo / (float)ORC_SYM, ORC_HALF);
      /* srslte_pss_chest output is never consumed -- not evaluated */
    } else {
      flags |= ORC_F_TAG_LOST;
      b->tracking_lost = 0;
    }
  } else {
    *nconsume = ORC_HALF;
  }
  if (rec) {
    rec->n_id_2 = b->n_id_2; rec->win_index = b->win_index; rec->flags = flags;
    rec->peak_pos = peak_used; rec->score = b->score; rec->psr = b->psr;
    rec->peak_value = b->core.peak_value; rec->cfo = cfo; rec->mean_cfo = mcfo;
    rec->m0 = rec->m1 = -1; rec->m0_val = rec->m1_val = 0.f; rec->n_id_1 = -1; rec->cell_id = -1;
    rec->cp_norm_avg = rec->cp_ext_avg = 0.f;
    rec->emit_start = frame_start;   /* caller adds win_start */
  }
  b->win_index++;
  return noutput;
}

void orc_pss_set_os_power(orc_pss *b, const float *os_power) { b->core.os_power = os_power; }
float orc_pss_max_psr(const orc_pss *b) { return b->psr_max; }
float orc_pss_mean_psr(const orc_pss *b) { return moving_avg(b->psr_data, b->psr_i); }
float orc_pss_mean_cfo(const orc_pss *b) { return moving_avg(b->cfo_data, b->cfo_i); }
float orc_pss_psr_threshold(const orc_pss *b) { return b->thr; }
void  orc_pss_set_psr_threshold(orc_pss *b, float t) { b->thr = t; }
float orc_pss_tracking_score(const orc_pss *b) { return (float)b->score; }

/* ------------------------------------------------------------------------- */
/* sss block  [ref lib/sss_impl.cc]                                           */
/* ------------------------------------------------------------------------- */
struct orc_sss {
  int n_id_2;
  int frame_type;                  /* 0 FDD (the reference), 1 TDD */
  float m_norm_avg, m_ext_avg;     /* srslte_sync_t CP EMA state [A.4] */
  int c0[31], c1[31], s_tilde[31], z_tilde[31], n_id_1_table[900];
};

orc_sss *orc_sss_new(int n_id_2)
{
  if (n_id_2 < 0 || n_id_2 > 2) return NULL;
  tables_init();
  orc_sss *s = calloc(1, sizeof *s);
  s->n_id_2 = n_id_2;
  orc_sss_tables(n_id_2, s->c0, s->c1, s->s_tilde, s->z_tilde, s->n_id_1_table);
  return s;
}
void orc_sss_free(orc_sss *s) { free(s); }
void orc_sss_set_frame_type(orc_sss *s, int frame_type) { s->frame_type = frame_type; }

/* canonical radix-2 DIT, natural-order output, twiddle W[k] = exp(-j 2 pi k/128) */
void orc_fft128(const orc_cf *in, orc_cf *out)
{
  tables_init();
  float ar[128], ai[128];
  for (int i = 0; i < 128; i++) {
    unsigned r = 0, v = (unsigned)i;
    for (int b = 0; b < 7; b++) { r = (r << 1) | (v & 1); v >>= 1; }
    ar[r] = in[i].re; ai[r] = in[i].im;
  }
  for (int half = 1; half < 128; half <<= 1) {
    int step = 64 / half;
    for (int base = 0; base < 128; base += 2 * half) {
      for (int j = 0; j < half; j++) {
        float wr = T.w_re[j * step], wi = T.w_im[j * step];
        float br = ar[base + j + half], bi = ai[base + j + half];
        float tr = fmaf(wr, br, -(wi * bi));
        float ti = fmaf(wr, bi, wi * br);
        float ur = ar[base + j], ui = ai[base + j];
        ar[base + j] = ur + tr; ai[base + j] = ui + ti;
        ar[base + j + half] = ur - tr; ai[base + j + half] = ui - ti;
      }
    }
  }
  for (int i = 0; i < 128; i++) { out[i].re = ar[i]; out[i].im = ai[i]; }
}

/* srslte_sync_detect_cp(q, in, peak_pos=960) [A.4] */
static int detect_cp(orc_sss *s, const orc_cf *in)
{
  const int fft = ORC_SYM, peak_pos = ORC_SLOT;
  const int cp_len[2] = {9, 32};
  int nsym = peak_pos / (fft + cp_len[1]); if (nsym > 3) nsym = 3;
  float R[2], Mv[2];
  for (int h = 0; h < 2; h++) {
    const int cp = cp_len[h];
    const orc_cf *p = &in[peak_pos - nsym * (fft + cp)];
    float Rs = 0.f, Cs = 0.f;
    for (int sy = 0; sy < nsym; sy++) {
      float dot = 0.f, pw = 0.f;
      for (int i = 0; i < cp; i++) {
        dot = fmaf(p[fft + i].re, p[i].re, dot);
        dot = fmaf(p[fft + i].im, p[i].im, dot);
      }
      for (int i = 0; i < cp; i++) {
        pw = fmaf(p[i].re, p[i].re, pw);
        pw = fmaf(p[i].im, p[i].im, pw);
      }
      Rs += dot;
      Cs += (float)cp * (pw / (float)cp);
      p += fft + cp;
    }
    R[h] = Rs;
    Mv[h] = (Cs > 0.f) ? Rs / Cs : 0.f;
  }
  float mn = Mv[0] / (float)nsym, me = Mv[1] / (float)nsym;
  s->m_norm_avg = (float)(0.1 * (double)mn + (1 - 0.1) * (double)s->m_norm_avg);
  s->m_ext_avg  = (float)(0.1 * (double)me + (1 - 0.1) * (double)s->m_ext_avg);
  if (s->m_norm_avg > s->m_ext_avg) return 1;
  if (s->m_norm_avg < s->m_ext_avg) return 0;
  return R[0] > R[1] ? 1 : 0;
}

/* srslte_sss_m0m1_partial(M=1, ce=NULL) [A.5] */
static void sss_m0m1(const orc_sss *s, const orc_cf *sym, int *m0, float *m0v, int *m1, float *m1v)
{
  orc_cf F[128];
  orc_fft128(sym, F);
  float y0r[31], y0i[31], y1r[31], y1i[31];
  for (int i = 0; i < 31; i++) {
    int e = 2 * i, o = 2 * i + 1;                       /* positions in the 62-carrier vector */
    int be = (e < 31) ? 97 + e : e - 30;                /* bins -31..-1 then +1..+31 */
    int bo = (o < 31) ? 97 + o : o - 30;
    y0r[i] = F[be].re * (float)s->c0[i]; y0i[i] = F[be].im * (float)s->c0[i];
    y1r[i] = F[bo].re * (float)s->c1[i]; y1i[i] = F[bo].im * (float)s->c1[i];
  }
  float corr[31];
  for (int m = 0; m < 31; m++) {
    float ar = 0.f, ai = 0.f;
    for (int i = 0; i < 31; i++) {
      float sv = (float)s->s_tilde[(i + m) % 31];
      ar = fmaf(y0r[i], sv, ar); ai = fmaf(y0i[i], sv, ai);
    }
    corr[m] = fmaf(ar, ar, ai * ai);
  }
  *m0 = vec_max_fi(corr, 31); *m0v = corr[*m0];
  for (int i = 0; i < 31; i++) {
    float z = (float)s->z_tilde[(i + (*m0 % 8)) % 31];
    y1r[i] *= z; y1i[i] *= z;
  }
  for (int m = 0; m < 31; m++) {
    float ar = 0.f, ai = 0.f;
    for (int i = 0; i < 31; i++) {
      float sv = (float)s->s_tilde[(i + m) % 31];
      ar = fmaf(y1r[i], sv, ar); ai = fmaf(y1i[i], sv, ai);
    }
    corr[m] = fmaf(ar, ar, ai * ai);
  }
  *m1 = vec_max_fi(corr, 31); *m1v = corr[*m1];
}

/* srslte_sss_N_id_1 [A.6] */
static int sss_n_id_1(const orc_sss *s, uint32_t m0, uint32_t m1)
{
  int nid = -1;
  if (m1 > m0) { if (m0 < 30 && m1 - 1 < 30) nid = s->n_id_1_table[m0 * 30 + (m1 - 1)]; }
  else         { if (m1 < 30 && m0 - 1 < 30) nid = s->n_id_1_table[m1 * 30 + (m0 - 1)]; }
  return nid;
}

/* lib/sss_impl.cc:83-156 */
int orc_sss_work(orc_sss *s, const orc_cf *in, int tag_lost, orc_cf *out, orc_rec *rec)
{
  if (tag_lost) {                                       /* :93-100 */
    s->m_norm_avg = s->m_ext_avg = 0.f;                 /* srslte_sync_reset */
    if (out) memcpy(out, in, sizeof(orc_cf) * ORC_HALF);
    return ORC_HALF;
  }
  int cp_norm = detect_cp(s, in);                       /* :104-108 */
  int cp_len = cp_norm ? 9 : 32;
  int sss_idx = ORC_SLOT - 2 * ORC_SYM - cp_len;        /* :110 */
  /* TDD (not in the reference, 36.211 6.11.2.2): the SSS is the last symbol of slots 1 and 11, three
   * symbols before the PSS (symbol 2 of slots 2 and 12); with the PSS body at [832, 960) slot 2 starts
   * at 832 - (2 (128 + cp) + cp) - (normal CP: 1 extra sample in symbol 0) and the SSS body ends there */
  if (s->frame_type == 1) sss_idx = ORC_SLOT - ORC_SYM - (3 * cp_len + 2 * ORC_SYM + (cp_norm ? 1 : 0)) - ORC_SYM;
  int m0, m1; float m0v, m1v;
  sss_m0m1(s, &in[sss_idx], &m0, &m0v, &m1, &m1v);      /* :112-116 */
  int nid = sss_n_id_1(s, (uint32_t)m0, (uint32_t)m1);  /* :118 */
  if (rec) {
    rec->flags |= ORC_F_SSS | (cp_norm ? ORC_F_CP_NORM : 0);
    rec->m0 = m0; rec->m1 = m1; rec->m0_val = m0v; rec->m1_val = m1v;
    rec->n_id_1 = nid;
    rec->cp_norm_avg = s->m_norm_avg; rec->cp_ext_avg = s->m_ext_avg;
  }
  if (nid < 0) return ORC_HALF;                         /* :119-120: no tags, out untouched */
  if (rec) { rec->cell_id = 3 * nid + s->n_id_2; rec->flags |= ORC_F_CELL; }   /* :124,:141-150 */
  if (out) memcpy(out, in, sizeof(orc_cf) * ORC_HALF);
  return ORC_HALF;
}

/* ------------------------------------------------------------------------- */
/* chains                                                                     */
/* ------------------------------------------------------------------------- */
int orc_chain_run(const orc_cf *y, int64_t n, int stream, int n_id_2, float thr,
                  int track_after, int track_every, int conv_mode, orc_rec *recs, int max_recs)
{
  const int frame_type = (conv_mode >> 8) & 1;              /* ORC_FRAME_TDD rides on conv_mode */
  conv_mode &= 0xff;
  orc_pss *p = orc_pss_new(n_id_2, thr, track_after, track_every, conv_mode);
  orc_sss *s = orc_sss_new(n_id_2);
  if (!p || !s) return -2;
  orc_sss_set_frame_type(s, frame_type);
  orc_cf *buf = calloc(n + ORC_SLOT, sizeof(orc_cf));       /* GR zero history in front */
  memcpy(buf + ORC_SLOT, y, sizeof(orc_cf) * n);
  orc_cf *out = malloc(sizeof(orc_cf) * ORC_HALF);
  float *os = NULL;
  if (conv_mode == ORC_CONV_OS) {
    /* whole blocks of the stream; a window's interior lags end 8765 samples before the data the
     * scheduler rule requires, so the incomplete last block is never read */
    os = calloc(n + 1024, sizeof(float));
    orc_pss_corr_os(y, n, n_id_2 == 0 ? os : NULL, n_id_2 == 1 ? os : NULL, n_id_2 == 2 ? os : NULL);
  }
  int nrec = 0; int64_t R = 0;
  while (R + ORC_LOOKAHEAD <= n) {
    if (nrec >= max_recs) { nrec = -1; break; }
    orc_rec *rec = &recs[nrec];
    int nconsume = 0;
    orc_pss_set_os_power(p, os ? os + R : NULL);
    int nout = orc_pss_work(p, buf + ORC_SLOT + R, out, &nconsume, rec);
    rec->stream = stream; rec->win_start = R;
    rec->emit_start = nout ? R + rec->emit_start : -1;
    if (nout) orc_sss_work(s, out, (rec->flags & ORC_F_TAG_LOST) != 0, NULL, rec);
    R += nconsume;
    nrec++;
  }
  free(buf); free(out); free(os); orc_pss_free(p); orc_sss_free(s);
  return nrec;
}

typedef struct {
  const void *iq; int fmt; int64_t n_in, n_out; int decim;
  float thr; int track_after, track_every, conv_mode;
  orc_cf **ys; orc_rec *tmp; int *cnt; size_t per_stream; int per_chain; int fail;
  float full_scale;
} trig_t;

static void trig_frontend(int s, void *arg)
{
  trig_t *t = (trig_t *)arg;
  if ((t->conv_mode & ORC_FRONT_TCINT) && t->decim > 1 && (t->fmt != 0 || t->full_scale > 0.0f)) {   /* integer front end */
    t->ys[s] = malloc(sizeof(orc_cf) * t->n_out);
    int64_t rc = t->fmt == 0 ? orc_decimate_tcint_fc32_d((const orc_cf *)t->iq + (size_t)s * t->n_in, t->n_in, t->decim, t->full_scale, t->ys[s])
               : t->fmt == 1 ? orc_decimate_tcint_sc16_d((const int16_t *)t->iq + (size_t)s * t->n_in * 2, t->n_in, t->decim, t->ys[s])
                             : orc_decimate_tcint_sc8_d((const int8_t *)t->iq + (size_t)s * t->n_in * 2, t->n_in, t->decim, t->ys[s]);
    if (rc < 0) t->fail = 1;
    return;
  }
  orc_cf *x = malloc(sizeof(orc_cf) * t->n_in);
  if (t->fmt == 1) orc_sc16_to_fc32((const int16_t *)t->iq + (size_t)s * t->n_in * 2, t->n_in, 1.0f / 32768.0f, x);
  else if (t->fmt == 2) orc_sc8_to_fc32((const int8_t *)t->iq + (size_t)s * t->n_in * 2, t->n_in, 1.0f / 128.0f, x);
  else memcpy(x, (const orc_cf *)t->iq + (size_t)s * t->n_in, sizeof(orc_cf) * t->n_in);
  if (t->decim > 1) {
    t->ys[s] = malloc(sizeof(orc_cf) * t->n_out);
    if ((t->conv_mode & 0xff) == ORC_CONV_FFT) orc_decimate_fast(x, t->n_in, t->decim, t->ys[s]);   /* timed baseline */
    else orc_decimate(x, t->n_in, t->decim, t->ys[s]);
    free(x);
  }
  else t->ys[s] = x;
}

static void trig_chain(int job, void *arg)
{
  trig_t *t = (trig_t *)arg;
  int s = job / 3, r = job % 3;
  int c = orc_chain_run(t->ys[s], t->n_out, s, r, t->thr, t->track_after, t->track_every, t->conv_mode,
                        t->tmp + (size_t)s * t->per_stream + (size_t)r * t->per_chain, t->per_chain);
  if (c < 0) { t->fail = 1; c = 0; }
  t->cnt[job] = c;
}

int orc_trigger_run(const void *iq, int fmt, int64_t n_in, int n_streams, int decim,
                    float thr, int track_after, int track_every, int conv_mode,
                    int nthreads, orc_rec *recs, int max_recs)
{
  return orc_trigger_run2(iq, fmt, n_in, n_streams, decim, thr, track_after, track_every, conv_mode, 0.0f, nthreads, recs, max_recs);
}

int orc_trigger_run2(const void *iq, int fmt, int64_t n_in, int n_streams, int decim,
                     float thr, int track_after, int track_every, int conv_mode, float fc32_full_scale,
                     int nthreads, orc_rec *recs, int max_recs)
{
  tables_init();
  if (thr <= 1.5f) thr = 1.5f;                              /* downlink_trigger_c.py:71-73 */
  trig_t t = { iq, fmt, n_in, 0, decim, thr, track_after, track_every, conv_mode, NULL, NULL, NULL, 0, 0, 0, fc32_full_scale };
  t.n_out = (decim <= 1) ? n_in : (n_in + decim - 1) / decim;
  t.per_chain = (int)(t.n_out / (ORC_HALF - ORC_SLOT)) + 2;
  t.per_stream = 3 * (size_t)t.per_chain;
  t.tmp = malloc(sizeof(orc_rec) * t.per_stream * n_streams);
  t.cnt = calloc((size_t)n_streams * 3, sizeof(int));
  t.ys = calloc(n_streams, sizeof(orc_cf *));
  /* front end once per stream (the reference's three chains share one resampler) */
  parallel_for(n_streams, nthreads, trig_frontend, &t);
  /* pss->sss chains, one job per (stream, N_id_2) */
  parallel_for(n_streams * 3, nthreads, trig_chain, &t);
  int total = 0, fail = t.fail;
  for (int job = 0; job < n_streams * 3 && !fail; job++) {
    if (total + t.cnt[job] > max_recs) { fail = 1; break; }
    memcpy(recs + total, t.tmp + (size_t)(job / 3) * t.per_stream + (size_t)(job % 3) * t.per_chain, sizeof(orc_rec) * t.cnt[job]);
    total += t.cnt[job];
  }
  for (int s = 0; s < n_streams; s++) free(t.ys[s]);
  free(t.ys); free(t.tmp); free(t.cnt);
  return fail ? -1 : total;
}
