/*
 * ltetrigger_oracle.h -- CPU oracle for the gr-ltetrigger PSS+SSS search path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it.  The product (gr-ltetrigger_b200/) never does.
 *
 * What it restates (all file:line are under the read-only reference tree):
 *   lib/pss_impl.cc:94-223   pss block: general_work, incr/reset_score, moving avg
 *   lib/sss_impl.cc:83-156   sss block: work
 *   python/downlink_trigger_c.py:13-73  three chains, threshold clamp
 * and, because the arithmetic itself lives in srsLTE release_18_06_1 and GNU Radio
 * 3.7 gr-filter (both absent from /root/reference and from this machine), the
 * published algorithms of
 *   srslte_pss_find_pss / srslte_pss_reset / srslte_pss_cfo_compute,
 *   srslte_cfo_correct + cexptab, srslte_sync_detect_cp / set_cp / reset,
 *   srslte_sss_m0m1_partial (M=1, ce=NULL), srslte_sss_N_id_1,
 *   gr::filter::firdes::low_pass (Kaiser) + rational_resampler_ccc(1, D)
 * as written down in SURVEY.md Appendix A.
 *
 * PARITY STATUS: the reference's own tests pin only (cell_id, cp_len) for the four
 * bundled test_frames (python/qa_downlink_trigger_c.py:67-203); this oracle is
 * checked against those.  Every intermediate quantity (peak index, PSR, CFO, m0/m1,
 * CP metric) is "parity unpinned" against real srsLTE: it could not be run here.
 *
 * Canonical arithmetic: float32 with explicit fused multiply-adds in a fixed order
 * (documented in DESIGN.md "Canonical arithmetic").  The CUDA kernels evaluate the
 * same expression trees, which is what makes bit-exact comparison possible.  A
 * second convolution mode (ORC_CONV_FFT) evaluates the PSS matched filter the way
 * the reference does -- one zero-padded 9728-point FFT convolution per window --
 * and is used as the CPU timing baseline and as a tolerance cross-check.
 */
#ifndef LTETRIGGER_ORACLE_H
#define LTETRIGGER_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } orc_cf;

enum { ORC_CONV_DIRECT = 0, ORC_CONV_FFT = 1, ORC_CONV_OS = 2 };
#define ORC_FRAME_TDD 0x100    /* or-ed into conv_mode of orc_chain_run / orc_trigger_run: TDD SSS position */
#define ORC_FRONT_TCINT 0x200  /* or-ed into conv_mode of orc_trigger_run: sc16 / sc8 input at decim 16 goes through the exact
                                  integer front end (orc_decimate_tcint_sc16) instead of the float32 decimator */
#define ORC_OS_STEP 896        /* outputs per 1024-point overlap-save block (ORC_CONV_OS) */

#define ORC_SLOT      960
#define ORC_HALF      9600
#define ORC_SYM       128
#define ORC_CONV_LEN  9726      /* lags examined by find_pss: conv_output_len-1 */
#define ORC_LOOKAHEAD 18365     /* largest nconsume: (9725-960)+9600 */
#define ORC_MAVG      200

/* flag bits of orc_rec.flags -- same meaning as ltb_window_rec.flags in the product */
#define ORC_F_SEARCHED   0x01   /* find_pss ran this call (lib/pss_impl.cc:163-169) */
#define ORC_F_OVER       0x02   /* d_psr > threshold (:174) */
#define ORC_F_EMIT       0x04   /* a half-frame was produced (:184-195) */
#define ORC_F_TRACKING   0x08   /* block was tracking when it emitted (:197) */
#define ORC_F_TAG_LOST   0x10   /* "tracking_lost" tag attached (:210-213) */
#define ORC_F_SSS        0x20   /* sss ran on this half-frame (lib/sss_impl.cc:104-116) */
#define ORC_F_CELL       0x40   /* cell_id + cp_type tags attached (:141-150) */
#define ORC_F_CP_NORM    0x80   /* cp_type == PMT_T */

typedef struct {
  int64_t win_start;    /* absolute search-rate index of the first new sample of this call */
  int64_t emit_start;   /* absolute index of the emitted half-frame's first sample (valid if EMIT) */
  int32_t stream;
  int32_t n_id_2;
  int32_t win_index;    /* ordinal of this general_work call */
  uint32_t flags;
  int32_t peak_pos;     /* d_peak_pos before the :186 overwrite (stale 960 on skipped calls) */
  int32_t score;        /* tracking score after incr/reset */
  float   psr;          /* d_psr (stale on skipped calls) */
  float   peak_value;   /* averaged correlation power at the peak (last search) */
  float   cfo;          /* srslte_pss_cfo_compute result (tracking emits only) */
  float   mean_cfo;     /* mean_cfo() used for the correction */
  int32_t m0, m1;       /* SSS indices (valid if SSS) */
  float   m0_val, m1_val;
  int32_t n_id_1;       /* -1 if srslte_sss_N_id_1 reported an error */
  int32_t cell_id;      /* 3*N_id_1 + N_id_2, -1 if none */
  float   cp_norm_avg, cp_ext_avg;   /* CP EMA state after this call (valid if SSS) */
} orc_rec;

/* ---- tables ------------------------------------------------------------ */
/* 128 time-domain matched-filter taps h = conj(t)/62 for N_id_2.      */
void orc_pss_taps(int n_id_2, float h_re[128], float h_im[128]);
/* GR rational_resampler_ccc(1,D) default taps. Returns ntaps (0 for D==1). */
int  orc_decim_taps(int decim, float *taps, int max_taps);
/* SSS tables for one N_id_2: c0/c1 (31), s_tilde/z_tilde (31), N_id_1 table (30x30).   */
void orc_sss_tables(int n_id_2, int c0[31], int c1[31], int s_tilde[31], int z_tilde[31], int n_id_1_table[900]);
void orc_cexptab(float tab_re[4097], float tab_im[4097]);   /* entry 4096 == entry 0 */
void orc_fft128_twiddles(float w_re[64], float w_im[64]);

/* ---- front end ----------------------------------------------------------- */
/* y[k] = sum_j taps[j] x[kD-j], zero initial state; n_out = ceil(n_in/D); -1 if D > 64. */
int64_t orc_decimate(const orc_cf *x, int64_t n_in, int decim, orc_cf *y);
/* the same filter evaluated output by output with SIMD-lane accumulators (what a CPU resampler does);
 * used by the timed reference-class mode (ORC_CONV_FFT) only */
int64_t orc_decimate_fast(const orc_cf *x, int64_t n_in, int decim, orc_cf *y);
void    orc_sc16_to_fc32(const int16_t *iq, int64_t n, float scale, orc_cf *out);
void    orc_sc8_to_fc32(const int8_t *iq, int64_t n, float scale, orc_cf *out);

/* ---- srsLTE pieces, exposed for unit tests -------------------------------- */
/* Raw correlation power |x (*) h|^2 at the 9726 lags of one zero-padded 9600-sample window. */
void orc_pss_corr_window(const orc_cf *win, int n_id_2, int conv_mode, float *power /*9726*/);
/* Sliding (untruncated) correlation power for a whole stream: P[n] for n in [0,n). x[<0]=0. */
void orc_pss_corr_stream(const orc_cf *x, int64_t n, int n_id_2, float *power);
void orc_fft128(const orc_cf *in, orc_cf *out);   /* forward, unnormalised, natural order */
/* ORC_CONV_OS: the canonical arithmetic of the GPU's overlap-save FFT correlator.  1024-point
 * four-step FFT (natural order in and out; inverse unscaled), its twiddle table, the filter
 * spectra (2^-10 folded in), and the block correlator over a whole stream (powers of the whole
 * 896-output blocks of x[0..n), x[<0] = 0; NULL skips a root; returns outputs per root). */
void orc_fft1024(const orc_cf *in, orc_cf *out, int inverse);
void orc_fft1024_twiddles(float w_re[1024], float w_im[1024]);
void orc_os_filter(int n_id_2, float H_re[1024], float H_im[1024]);
int64_t orc_pss_corr_os(const orc_cf *x, int64_t n, float *p0, float *p1, float *p2);

/* ---- blocks ---------------------------------------------------------------- */
typedef struct orc_pss orc_pss;
typedef struct orc_sss orc_sss;

orc_pss *orc_pss_new(int n_id_2, float psr_threshold, int track_after, int track_every, int conv_mode);
void     orc_pss_free(orc_pss *);
/* One general_work call.  `in` points at the first new sample; in[-960 .. 18365) must be
 * readable.  Writes 0 or 9600 samples to out, returns noutput; *nconsume as consume_each. */
int      orc_pss_work(orc_pss *, const orc_cf *in, orc_cf *out, int *nconsume, orc_rec *rec);
/* ORC_CONV_OS only: os_power[k] = overlap-save power at lag k of the window the next work call
 * searches (orc_chain_run sets it; the block cannot know the stream's absolute alignment) */
void     orc_pss_set_os_power(orc_pss *, const float *os_power);
float    orc_pss_max_psr(const orc_pss *);
float    orc_pss_mean_psr(const orc_pss *);
float    orc_pss_mean_cfo(const orc_pss *);
float    orc_pss_psr_threshold(const orc_pss *);
void     orc_pss_set_psr_threshold(orc_pss *, float);
float    orc_pss_tracking_score(const orc_pss *);

orc_sss *orc_sss_new(int n_id_2);
void     orc_sss_free(orc_sss *);
/* 0 FDD (the reference): SSS one symbol before the PSS; 1 TDD: three symbols before it */
void     orc_sss_set_frame_type(orc_sss *, int frame_type);
/* One work call on an aligned half-frame; tag_lost = "tracking_lost" tag present on item 0.
 * Returns 9600. rec gets the SSS fields and flag bits merged in. */
int      orc_sss_work(orc_sss *, const orc_cf *in, int tag_lost, orc_cf *out, orc_rec *rec);

/* LTB_FRONTEND_TC_INT restated: exact integer decimate-by-16 of interleaved int16 I/Q (zero history), one
 * rounding per output.  Returns the number of outputs, -1 on error. */
int64_t orc_decimate_tcint_sc16(const int16_t *iq, int64_t n_in, orc_cf *y);
int64_t orc_decimate_tcint_sc8(const int8_t *iq, int64_t n_in, orc_cf *y);
/* fc32 input taken as 23-bit fixed point over +-full_scale, then the same exact integers (see the .c file) */
int64_t orc_decimate_tcint_fc32(const orc_cf *x, int64_t n_in, float full_scale, orc_cf *y);
/* the same at the other rates the product's kernel has (decim 2 .. 16): taps scaled by 2^(23 + floor(log2 decim)) */
int64_t orc_decimate_tcint_sc16_d(const int16_t *iq, int64_t n_in, int decim, orc_cf *y);
int64_t orc_decimate_tcint_sc8_d(const int8_t *iq, int64_t n_in, int decim, orc_cf *y);
int64_t orc_decimate_tcint_fc32_d(const orc_cf *x, int64_t n_in, int decim, float full_scale, orc_cf *y);

/* ---- whole chains ------------------------------------------------------------ */
/* pss(N_id_2) -> sss(N_id_2) over one search-rate stream y[0..n) (GR zero history before it),
 * driven by the scheduler rule "call general_work only while win_start + 18365 <= n".
 * One record per call.  Returns the number of records (<= max_recs). */
int orc_chain_run(const orc_cf *y, int64_t n, int stream, int n_id_2, float psr_threshold,
                  int track_after, int track_every, int conv_mode,
                  orc_rec *recs, int max_recs);

/* downlink_trigger_c topology over a batch of equal-length streams of raw input
 * (fc32 if fmt==0, sc16 if fmt==1, sc8 if fmt==2; `decim` 1..64); three chains per stream.
 * Records are ordered (stream, n_id_2, win_index).  nthreads<=0 -> all cores.
 * Returns the record count, or -1 if max_recs is too small. */
int orc_trigger_run(const void *iq, int fmt, int64_t n_in_per_stream, int n_streams, int decim,
                    float psr_threshold, int track_after, int track_every, int conv_mode,
                    int nthreads, orc_rec *recs, int max_recs);
/* the same with the range of the fixed-point grid that ORC_FRONT_TCINT puts fc32 input on (0: fc32 input keeps the
 * float32 decimator) */
int orc_trigger_run2(const void *iq, int fmt, int64_t n_in_per_stream, int n_streams, int decim,
                     float psr_threshold, int track_after, int track_every, int conv_mode, float fc32_full_scale,
                     int nthreads, orc_rec *recs, int max_recs);

#ifdef __cplusplus
}
#endif
#endif
