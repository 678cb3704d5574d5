/*
 * ltetrigger_b200.h -- C ABI of libltetrigger_b200.so
 *
 * B200-native (sm_100a) implementation of the one hot path of NTIA/gr-ltetrigger:
 * the LTE PSS+SSS synchronisation-signal search.  Plain C, caller-owned buffers, no
 * exceptions, no torch/GNU Radio types.  Every entry point names the reference
 * interface it replaces (paths are relative to the reference tree; "srslte_*" symbols
 * are the srsLTE release_18_06_1 calls made from those lines -- the reference's inner
 * FFI boundary).
 *
 * Return convention (same as srsLTE, tested at lib/sss_impl.cc:119):
 *    0  LTB_SUCCESS
 *   -1  LTB_ERROR                 (CUDA failure, no device, ...; see ltb_last_error)
 *   -2  LTB_ERROR_INVALID_INPUTS
 * One host thread per object at a time (as the reference's blocks: one scheduler
 * thread calls work); accessors may be polled from other threads and are, like the
 * reference's (lib/pss_impl.h:95-100), unsynchronised snapshots.
 */
#ifndef LTETRIGGER_B200_H
#define LTETRIGGER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTB_API __attribute__((visibility("default")))

#define LTB_SUCCESS               0
#define LTB_ERROR                (-1)
#define LTB_ERROR_INVALID_INPUTS (-2)

/* geometry at the 1.92 Msps search rate (lib/pss_impl.h:52-55, lib/sss_impl.h:43-47) */
#define LTB_SLOT_LEN      960
#define LTB_HALF_FRAME    9600
#define LTB_SYMBOL_SZ     128
#define LTB_CONV_LEN      9726     /* lags examined per window by srslte_pss_find_pss */
#define LTB_LOOKAHEAD     18365    /* largest consume of one general_work call: (9725-960)+9600 */
#define LTB_MOVING_AVG_SZ 200      /* lib/pss_impl.h:31 */
#define LTB_MIN_PSR_THRESHOLD 1.5f /* python/downlink_trigger_c.py:10 */

/* input sample formats (little-endian interleaved I/Q) */
#define LTB_FMT_FC32 0             /* gr_complex, 8 B/sample: what the reference consumes */
#define LTB_FMT_SC16 1             /* int16 I/Q, 4 B/sample, scaled by 1/32768 on device */
#define LTB_FMT_SC8  2             /* int8 I/Q, 2 B/sample, scaled by 1/128 on device */
#define LTB_MAX_DECIM 64

/* evaluation of the three-root matched filter (both bit-exact against the oracle's restatement) */
#define LTB_CORR_DIRECT 0          /* folded direct form, FFMA2 */
#define LTB_CORR_FFT    1          /* 1024-point overlap-save FFT blocks aligned to absolute sample indices */
#define LTB_OS_STEP     896        /* outputs per overlap-save block */

/* frame structure: where the SSS sits relative to the PSS (36.211 6.11.2.2) */
#define LTB_FRAME_FDD   0          /* the symbol before the PSS: the reference (lib/sss_impl.cc:110) */
#define LTB_FRAME_TDD   1          /* three symbols before the PSS; cell id and CP type only -- the emitted
                                      half-frame keeps the reference's PSS alignment, which in TDD does not
                                      start at a subframe boundary, so ltb_mib_decode does not apply */

/* arithmetic of the decimating front end */
#define LTB_FRONTEND_FP32   0      /* canonical float32 expression trees (FFMA2), every format and rate */
#define LTB_FRONTEND_TC_INT 1      /* exact integer arithmetic on the tensor cores (tcgen05.mma kind::i8): taps
                                      quantised to three balanced base-256 digits, the int16 / int8 samples are
                                      their own digits, int32 accumulation in TMEM, one rounding to float32 per
                                      output.  decim 2 / 4 / 8 / 12 / 16 / 24 / 32 for fc32, from 4 for sc16, from 8 for
                                      sc8 (the LTE sampling rates and 46.08 / 61.44 Msps; a 16-output row must be whole
                                      256-byte pieces);
                                      needs 16-byte aligned rows.  fc32 input is first
                                      put on a 23-bit fixed-point grid over +-fc32_full_scale (what a float sample
                                      of an ADC-fed source carries anyway); from there on the same exact integers */

/* how the stages of consecutive calls are scheduled (results are identical) */
#define LTB_PIPE_OVERLAP 0         /* with two calls in flight (submit/collect), the per-chain track + SSS kernels of
                                      call i run on a second, low-priority stream under the front end and correlator
                                      of call i+1 (rings sized for two chunks); forced off by keep_halfframes */
#define LTB_PIPE_SERIAL  1         /* every kernel of a call on one stream, calls back to back */

typedef struct { float re, im; } ltb_cf;

/* ---- per-window record ---------------------------------------------------------
 * One record per pss::general_work call of one chain (stream, N_id_2), carrying what
 * the reference expresses as return values, consume counts, stream tags and the
 * downstream sss::work result for the half-frame that call emitted.                */
#define LTB_F_SEARCHED 0x01u  /* srslte_pss_find_pss ran (lib/pss_impl.cc:163-169) */
#define LTB_F_OVER     0x02u  /* d_psr > d_psr_threshold (:174) */
#define LTB_F_EMIT     0x04u  /* one aligned half-frame produced (:184-195) */
#define LTB_F_TRACKING 0x08u  /* emitted in tracking state: CFO-corrected, no pss tag (:197-209) */
#define LTB_F_TAG_LOST 0x10u  /* stream tag "tracking_lost" on item 0 (:210-213) */
#define LTB_F_SSS      0x20u  /* sss::work decoded this half-frame (lib/sss_impl.cc:104-118) */
#define LTB_F_CELL     0x40u  /* stream tags "cell_id" and "cp_type" attached (:141-150) */
#define LTB_F_CP_NORM  0x80u  /* cp_type == PMT_T (normal CP) */

typedef struct {
  int64_t  win_start;    /* absolute search-rate index of the call's first new sample (nitems_read) */
  int64_t  emit_start;   /* absolute index of the emitted half-frame's first sample; -1 if none */
  int32_t  stream;
  int32_t  n_id_2;
  int32_t  win_index;    /* ordinal of the general_work call on this chain */
  uint32_t flags;        /* LTB_F_* */
  int32_t  peak_pos;     /* d_peak_pos used by this call (stale 960 on skipped searches) */
  int32_t  score;        /* tracking_score() after the call */
  float    psr;          /* d_psr (stale on skipped searches) */
  float    peak_value;   /* averaged correlation power at the peak of the last search */
  float    cfo;          /* srslte_pss_cfo_compute (tracking emits only) */
  float    mean_cfo;     /* mean_cfo() used for the in-place correction */
  int32_t  m0, m1;       /* srslte_sss_m0m1_partial results (-1 if SSS not run) */
  float    m0_val, m1_val;
  int32_t  n_id_1;       /* srslte_sss_N_id_1; -1 on SRSLTE_ERROR / not run */
  int32_t  cell_id;      /* srslte_sync_get_cell_id = 3*N_id_1 + N_id_2; -1 if none */
  float    cp_norm_avg, cp_ext_avg;  /* srslte_sync_detect_cp EMA state after the call */
} ltb_window_rec;        /* 88 bytes */

/* accessors of ltetrigger::pss (include/ltetrigger/pss.h:72-87, lib/pss_impl.h:95-100) */
typedef struct {
  float   max_psr;
  float   mean_psr;
  float   mean_cfo;
  float   psr_threshold;
  float   tracking_score;
  int32_t tracking;      /* bool(d_tracking) */
  int64_t next_window;   /* absolute index the chain's next general_work call starts at */
} ltb_pss_stats;

/* ---- batched trigger engine -----------------------------------------------------
 * Replaces, for n_streams independent IQ streams at once, the stream path of
 *   rational_resampler_ccc(1, decim)            examples/cell_search_file.py:56-57
 *   -> downlink_trigger_c(psr_threshold)        python/downlink_trigger_c.py:18-45
 *        3 x ( pss(N_id_2=k) -> sss(N_id_2=k) )   lib/pss_impl.cc, lib/sss_impl.cc
 * up to (not including) the host-side mib block.  The GNU Radio scheduler's role is
 * fixed to: a chain's general_work is called only while win_start + 18365 <= samples
 * received, so results do not depend on how the stream is cut into chunks.          */
typedef struct ltb_trigger ltb_trigger;

typedef struct {
  uint32_t struct_size;       /* sizeof(ltb_trigger_config), for ABI evolution */
  int32_t  device;            /* CUDA device ordinal */
  int32_t  n_streams;
  int32_t  input_format;      /* LTB_FMT_* */
  int32_t  decim;             /* input rate / 1.92 Msps, any integer 1..LTB_MAX_DECIM as the reference's
                                 CLI accepts (examples/cell_search_file.py:50-57); streaming kernels for
                                 4, 8, 12, 16, the tiled kernel for the other rates up to 15, a general one above */
  int32_t  root_mask;         /* bit k set: run the N_id_2 = k chain; 0 -> 7 (all three) */
  int64_t  max_chunk;         /* largest n_samples (input rate, per stream) of one process call */
  float    psr_threshold;     /* clamped to > 1.5 like downlink_trigger_c.py:71-73 */
  int32_t  track_after;       /* 0 -> 16  (include/ltetrigger/pss.h:68) */
  int32_t  track_every;       /* 0 -> 8 */
  int32_t  record_all;        /* 1: record every general_work call; 0: emitted half-frames only */
  int32_t  keep_halfframes;   /* 1: keep each emitted (CFO-corrected) half-frame for ltb_trigger_fetch_halfframes */
  void    *cuda_stream;       /* cudaStream_t the caller produces device input on (NULL: none): the front end of a
                                 call waits for the work queued there at submit time.  The kernels themselves run
                                 on the library's own two streams; results are complete when collect returns */
  int32_t  corr_mode;         /* LTB_CORR_*; a struct_size that ends before this field selects LTB_CORR_DIRECT */
  int32_t  frame_type;        /* LTB_FRAME_*; default (0, or a shorter struct_size) is FDD like the reference */
  int32_t  frontend_mode;     /* LTB_FRONTEND_*; default 0 = canonical FP32 */
  int32_t  pipeline;          /* LTB_PIPE_*; default 0 = overlapped */
  float    fc32_full_scale;   /* LTB_FRONTEND_TC_INT on fc32 input only: the range the samples are quantised over
                                 (23-bit fixed point, |re|, |im| beyond it saturate); 1.0 for a UHD-style source */
} ltb_trigger_config;

/* pss::make + sss::make + hier-block construction (lib/pss_impl.cc:42-83,
 * lib/sss_impl.cc:45-73, python/downlink_trigger_c.py:18-45). */
LTB_API int ltb_trigger_create(const ltb_trigger_config *cfg, ltb_trigger **out);
/* ~pss_impl / ~sss_impl (lib/pss_impl.cc:88-92, lib/sss_impl.cc:78-81) */
LTB_API int ltb_trigger_destroy(ltb_trigger *t);
/* back to the just-constructed state (new flowgraph run) */
LTB_API int ltb_trigger_reset(ltb_trigger *t);
/* downlink_trigger_c.set_psr_threshold (python/downlink_trigger_c.py:63-69) /
 * pss::set_psr_threshold (lib/pss_impl.h:98).  stream / n_id_2 = -1 selects all (other negative
 * values are rejected); clamp != 0 applies the hier block's > 1.5 floor. */
LTB_API int ltb_trigger_set_psr_threshold(ltb_trigger *t, int stream, int n_id_2, float thr, int clamp);

/* Feed n_samples new input-rate samples per stream (a multiple of 8*decim) and run every
 * chain as far as the lookahead rule allows.  Stream s starts at
 * (char*)iq + s*stream_stride_bytes; pointer and stride must be multiples of the sample size, and
 * multiples of 16 bytes for the fastest decimators (otherwise a slower kernel gives the same bits).
 * *_host takes host memory (pinned for full PCIe
 * rate) and copies it in; *_device takes device memory on cfg.device.  Records are
 * written to `recs` ordered by (stream, n_id_2, win_index); *n_recs is the count
 * (if it exceeds max_recs, collect returns LTB_ERROR_INVALID_INPUTS with the needed count in *n_recs and
 * leaves the call pending: call ltb_trigger_collect again with a larger buffer).  Replaces one scheduler pass of general_work/work calls over the chunk
 * (lib/pss_impl.cc:154-223, lib/sss_impl.cc:83-156). */
LTB_API int ltb_trigger_process_host(ltb_trigger *t, const void *iq, int64_t stream_stride_bytes,
                                     int64_t n_samples, ltb_window_rec *recs, int max_recs, int *n_recs);
LTB_API int ltb_trigger_process_device(ltb_trigger *t, const void *d_iq, int64_t stream_stride_bytes,
                                       int64_t n_samples, ltb_window_rec *recs, int max_recs, int *n_recs);
/* Asynchronous halves of process_device: submit enqueues all kernels on the stream and
 * returns; collect waits for the oldest submitted call and copies its records out.  Up to two
 * calls may be submitted before the first is collected (one with cfg.keep_halfframes), which
 * keeps the stream busy across calls; the input buffer of a call must stay valid until it is
 * collected. */
LTB_API int ltb_trigger_submit_device(ltb_trigger *t, const void *d_iq, int64_t stream_stride_bytes,
                                      int64_t n_samples);
LTB_API int ltb_trigger_collect(ltb_trigger *t, ltb_window_rec *recs, int max_recs, int *n_recs);
/* The same for host input: the host->device copy of call i+1 runs on its own stream while the
 * kernels of call i execute.  The host buffer must stay valid (and, for full PCIe rate, pinned)
 * until the call is collected. */
LTB_API int ltb_trigger_submit_host(ltb_trigger *t, const void *iq, int64_t stream_stride_bytes, int64_t n_samples);

/* max_psr / mean_psr / mean_cfo / psr_threshold / tracking_score of one chain */
LTB_API int ltb_trigger_get_stats(ltb_trigger *t, int stream, int n_id_2, ltb_pss_stats *out);
/* The half-frames emitted by the last process call (cfg.keep_halfframes): 9600 samples
 * each, in record order over records with LTB_F_EMIT -- what pss writes to its output
 * port (lib/pss_impl.cc:193,204) and sss passes through (lib/sss_impl.cc:152). */
LTB_API int ltb_trigger_fetch_halfframes(ltb_trigger *t, ltb_cf *out, int max_halfframes, int *n_halfframes);
/* device time of the last process/submit call's kernels, measured with CUDA events on
 * the launch stream (ms); and the number of kernel launches it made */
LTB_API int ltb_trigger_last_timing(ltb_trigger *t, float *ms_total, int *n_launches);
/* the same split by stage, CUDA events between the launches: [0] convert+decimate front end,
 * [1] three-root PSS correlator, [2] per-chain track kernel, [3] batched SSS */
LTB_API int ltb_trigger_last_kernel_times(ltb_trigger *t, float ms[4]);
LTB_API const char *ltb_last_error(void);
LTB_API const char *ltb_version(void);
LTB_API int ltb_device_count(void);

/* ---- standalone sss block ----------------------------------------------------------
 * ltetrigger::sss for one N_id_2 on caller-supplied aligned half-frames
 * (lib/sss_impl.cc:45-156): srslte_sync_detect_cp / set_cp, srslte_sss_m0m1_partial,
 * srslte_sss_N_id_1, srslte_sync_get_cell_id, srslte_sync_reset on "tracking_lost". */
typedef struct ltb_sss ltb_sss;
LTB_API int ltb_sss_create(int device, int n_id_2, ltb_sss **out);
LTB_API int ltb_sss_destroy(ltb_sss *s);
/* LTB_FRAME_FDD (default) or LTB_FRAME_TDD */
LTB_API int ltb_sss_set_frame_type(ltb_sss *s, int frame_type);
/* n_halfframes consecutive work() calls: in = n*9600 host samples, tag_lost[i] != 0 if
 * the i-th half-frame carries "tracking_lost".  Fills the SSS fields and flag bits of
 * recs[i] (other fields untouched).  Returns LTB_SUCCESS. */
LTB_API int ltb_sss_work(ltb_sss *s, const ltb_cf *in, const int32_t *tag_lost, int n_halfframes,
                         ltb_window_rec *recs);

/* ---- host-side MIB decode (consumer of the path's output; never touches the GPU) ------------
 * What ltetrigger::mib does with a tagged half-frame (lib/mib_impl.cc:148-170:
 * srslte_ue_mib_decode + srslte_pbch_mib_unpack): PBCH of slot 1, one antenna port or two with
 * transmit diversity (hypotheses tried in that order, CRC mask must match the hypothesis).  halfframe = 9600 aligned, CFO-corrected samples as emitted by pss / passed by
 * sss; cell_id and cp_normal from the sss tags.  Returns 1 (SRSLTE_UE_MIB_FOUND) and fills *out,
 * 0 if no MIB was found (e.g. a subframe-5 half-frame), < 0 on invalid arguments. */
typedef struct {
  int32_t nof_prb;          /* 6, 15, 25, 50, 75, 100 */
  int32_t nof_ports;        /* 1, 2, 4 (from the CRC mask) */
  int32_t phich_length;     /* 0 Normal, 1 Extended */
  int32_t phich_resources;  /* 0 "1/6", 1 "1/2", 2 "1", 3 "2" */
  int32_t sfn;              /* system frame number of this frame */
  int32_t sfn_offset;       /* frame position inside the 40 ms PBCH period (scrambling phase) */
} ltb_mib;
LTB_API int ltb_mib_decode(const ltb_cf *halfframe, int cell_id, int cp_normal, ltb_mib *out);

/* ---- kernel-level entry points (parity tests, profiling) ----------------------------- */
/* Sliding matched-filter power |x (*) h_k|^2, k = 0,1,2, for n (multiple of 8) samples of
 * each of n_streams host streams; x[<0] = 0.  power: [n_streams][3][n].
 * = srslte_pss_find_pss's convolution + srslte_vec_abs_square_cf without window truncation. */
LTB_API int ltb_kernel_pss_corr_host(int device, const ltb_cf *x, int n_streams, int64_t n, float *power);
/* The same powers from the overlap-save FFT evaluation (LTB_CORR_FFT) for the whole 896-output blocks
 * of each stream: power: [n_streams][3][896 * (n / 896)]. */
LTB_API int ltb_kernel_pss_corr_fft_host(int device, const ltb_cf *x, int n_streams, int64_t n, float *power);
/* rational_resampler_ccc(1, decim) with default taps on host streams of n_in samples
 * (multiple of decim); fmt as above; y: [n_streams][n_in/decim]. */
LTB_API int ltb_kernel_decimate_host(int device, const void *x, int fmt, int n_streams, int64_t n_in,
                                     int decim, ltb_cf *y);

/* LTB_FRONTEND_TC_INT at kernel level: decimate-by-`decim` of n_streams host streams of n_in interleaved int16
 * (fmt LTB_FMT_SC16: decim 4, 8, 12, 16, 24, 32), int8 (LTB_FMT_SC8: 8 ... 32) or float (LTB_FMT_FC32: 2 ... 32; taken as
 * 23-bit fixed point over +-full_scale) I/Q samples, fed to the tensor-core kernel in calls of `chunk` samples (both
 * multiples of 8 decim; the raw history is carried between the calls as the engine does); y: [n_streams][n_in / decim]. */
LTB_API int ltb_kernel_decimate_tc_host(int device, const void *x, int fmt, int decim, float full_scale, int n_streams,
                                        int64_t n_in, int64_t chunk, ltb_cf *y);

#ifdef LTB_DEBUG
/* Only in the debug build (make -C gr-ltetrigger_b200 debug -> lib/libltetrigger_b200_debug.so, -DLTB_DEBUG);
 * the release library does not export it.  Profiling / cross-check aid, never needed for results.
 * flag 0: decimator dissection (bit 0: skip the staging copies, bit 1: skip the FMA body; outputs are
 * garbage while set); flag 1: route the decimator through its general (bit 0) or tiled (bit 1) kernel. */
LTB_API int ltb_debug_set_flag(int flag, int value);
#endif

/* ---- tables (host only; no GPU needed) -------------------------------------------------- */
/* srslte_pss_init + srslte_pss_set_N_id_2: 128 conj time-domain taps (lib/pss_impl.cc:72-75) */
LTB_API int ltb_table_pss_taps(int n_id_2, float h_re[128], float h_im[128]);
/* gr::filter::rational_resampler_ccc(1, decim) default taps; returns ntaps or <0 */
LTB_API int ltb_table_decim_taps(int decim, float *taps, int max_taps);
/* srslte_sss_init + srslte_sss_set_N_id_2 tables (lib/sss_impl.cc:63-70) */
LTB_API int ltb_table_sss(int n_id_2, int32_t c0[31], int32_t c1[31], int32_t s_tilde[31],
                          int32_t z_tilde[31], int32_t n_id_1_table[900]);
/* srslte_cfo_init's cexptab, 4096 (+1 spare) entries (lib/pss_impl.cc:78) */
LTB_API int ltb_table_cexp(float tab_re[4097], float tab_im[4097]);
LTB_API int ltb_table_fft128_twiddles(float w_re[64], float w_im[64]);
/* LTB_CORR_FFT: W_1024^i and the filter spectrum 2^-10 DFT_1024(h) of one N_id_2, natural order */
LTB_API int ltb_table_fft1024_twiddles(float w_re[1024], float w_im[1024]);
LTB_API int ltb_table_os_filter(int n_id_2, float H_re[1024], float H_im[1024]);
/* LTB_FRONTEND_TC_INT: the tap table of the tensor-core kernel for input format fmt at rate decim, in its shared-memory image
 * ([208 rows][128 bytes], K-major, 128-byte swizzle: 16-byte chunk c of row r sits at chunk c ^ (r & 7)), and the sum
 * of the integer taps T[j] = rint(taps[j] * 2^(23 + floor(log2 decim))).  tests/test_tc_formulation.py replays the kernel's
 * k-steps from it. */
LTB_API int ltb_table_tc_btab(int fmt, int decim, int8_t tab[208 * 128], int64_t *sum_t);

#ifdef __cplusplus
}
#endif
#endif /* LTETRIGGER_B200_H */
