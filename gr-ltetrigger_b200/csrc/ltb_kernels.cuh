// sm_100a kernels of the PSS+SSS search path.  Included once by ltb_api.cu.
//
// Arithmetic contract ("canonical arithmetic"): float32, every fused
// multiply-add is an explicit fma (FFMA / packed FFMA2), every other product or sum is an
// individually rounded __fmul_rn/__fadd_rn, accumulation orders are fixed.  The CPU oracle
// (oracle/ltetrigger_oracle.c, test-only) evaluates the same expression trees, so the
// parity tests compare bits, not tolerances.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/ltetrigger_b200.h"

namespace ltb {

// ------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------
constexpr int kSlot = LTB_SLOT_LEN;        // 960
constexpr int kHalf = LTB_HALF_FRAME;      // 9600
constexpr int kSym = LTB_SYMBOL_SZ;        // 128
constexpr int kNLag = LTB_CONV_LEN;        // 9726
constexpr int kAvgLen = 9732;              // 9729 used (fft+frame+1), padded to 16 B
constexpr int kLookahead = LTB_LOOKAHEAD;  // 18365
constexpr int kMavg = LTB_MOVING_AVG_SZ;   // 200
constexpr int kRotBase = 352;              // first sample of the emitted half-frame kept CFO-corrected in shared memory
constexpr int kRotLen = kSlot - kRotBase;  // 608: reaches back to the TDD extended-CP SSS symbol
constexpr int kMaxDecim = 64;              // rational_resampler_ccc(1, D) for D = 1..64
constexpr int kTailCap = 33 * kMaxDecim;   // decimator history kept per stream (input samples)

// input formats: bytes per complex sample and the scale applied on conversion to float
__host__ __device__ constexpr int fmt_bytes(int fmt) { return fmt == LTB_FMT_FC32 ? 8 : fmt == LTB_FMT_SC16 ? 4 : 2; }
__host__ __device__ constexpr float fmt_scale(int fmt) {
  return fmt == LTB_FMT_SC16 ? 1.0f / 32768.0f : fmt == LTB_FMT_SC8 ? 1.0f / 128.0f : 1.0f;
}
template <int FMT> struct fmt_elem { typedef float2 type; };
template <> struct fmt_elem<LTB_FMT_SC16> { typedef short2 type; };
template <> struct fmt_elem<LTB_FMT_SC8> { typedef char2 type; };

// folded matched-filter coefficients, group 0 = root 25, group 1 = root 29 (root 34 = conj):
//   [g][m][0] = (hr, hi)   [g][m][1] = (hi, hr)      m = 0..64
__constant__ float2 c_pss_coef[2][65][2];
// full 128-tap filters per N_id_2 (CFO estimate): (re, im)
__constant__ float2 c_pss_taps[3][128];
// decimator taps for D = 4, 8, 16 at offsets 72, 208, 472 and D = 12..15 from 1000 (padded with zeros):
// the streaming kernels
__constant__ float c_decim_taps[1000 + 33 * (12 + 13 + 14 + 15)];
__constant__ float2 c_fft128_tw[64];
// SSS tables per N_id_2: c0, c1 (31 each); shared s_tilde, z_tilde; N_id_1 table
__constant__ float c_sss_c0[3][32];
__constant__ float c_sss_c1[3][32];
__constant__ float c_sss_s[32];
__constant__ float c_sss_z[32];
__constant__ short c_sss_nid1[900];

__host__ __device__ constexpr int decim_tap_offset(int d) {
  return d == 2 ? 0 : d == 4 ? 72 : d == 8 ? 208 : d == 16 ? 472 : 1000 + 33 * ((d - 12) * (d + 11) / 2);   // 12..15 back to back
}
__host__ __device__ constexpr int decim_ntaps(int d) {
  return d == 2 ? 65 : d == 4 ? 131 : d == 8 ? 263 : d == 12 ? 393 : d == 13 ? 427 : d == 14 ? 459 : d == 15 ? 493 : d == 16 ? 525 : 0;
}

// ------------------------------------------------------------------------------------
// per-chain state (one pss block + one sss block of the reference)
// ------------------------------------------------------------------------------------
struct ChainState {
  long long next_win;        // absolute search-rate index of the next general_work call (nitems_read)
  int score, timer, tracking, lost;   // tracking_t + d_tracking_lost (lib/pss_impl.h:41-63)
  int peak_pos;              // d_peak_pos
  int win_index;
  float psr, psr_max, peak_value;
  unsigned psr_i, cfo_i;
  float cfo_last_freq;       // d_cfo.last_freq
  float cfo_table_freq;      // frequency the phasor table currently holds
  float cp_norm_avg, cp_ext_avg;   // srslte_sync_t M_norm_avg / M_ext_avg
  float psr_data[kMavg];
  float cfo_data[kMavg];
};

struct TrackParams {
  const float2 *y_ring;      // [n_streams][cap]
  const float *p_ring;       // [n_streams][3][cap]
  ChainState *state;         // [n_streams*3]
  float *avg;                // [n_streams*3][kAvgLen]
  const float *thr;          // [n_streams*3]
  ltb_window_rec *recs;      // [n_streams*3][w_max]
  int *rec_count;            // [n_streams*3]
  float2 *sss_sym;           // [sss_cap][128]
  int *sss_rec;              // [sss_cap] -> index into recs
  int *sss_count;            // global counter
  int *chain_counter;        // work queue head (zeroed before every launch)
  const int *chain_order;    // queue position -> chain: tracking chains (the long jobs) first
  int n_chains;
  int sss_cap;
  float2 *hf_out;            // [n_streams*3][w_max][9600] or null
  const float2 *cexp;        // 4097 entries
  long long n_total;         // search-rate samples received so far
  unsigned cap_mask;
  int cap;
  int w_max;
  int track_after, track_every;
  int record_all;
  int root_mask;
  int tdd;                   // LTB_FRAME_TDD: SSS three symbols before the PSS
};

// ------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }

__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;\n" ::: "memory");
}

// first sample of the SSS symbol in an aligned half-frame (PSS body at [832, 960)).  FDD: the symbol
// before the PSS (lib/sss_impl.cc:110).  TDD (frame structure type 2, 36.211 6.11.2.2; not in the
// reference): the last symbol of the previous slot, three symbols before the PSS, and the first
// symbol of a normal-CP slot carries one extra prefix sample.
__host__ __device__ constexpr int sss_symbol_start(int cp_len, int tdd) {
  return tdd ? kSlot - kSym - (3 * cp_len + 2 * kSym + (cp_len == 9 ? 1 : 0)) - kSym : kSlot - 2 * kSym - cp_len;
}

// canonical complex product a*b:  re = fma(ar, br, -(ai*bi)),  im = fma(ar, bi, ai*br)
__device__ __forceinline__ float2 cmul_canon(float2 a, float2 b) {
  float2 r;
  r.x = __fmaf_rn(a.x, b.x, -__fmul_rn(a.y, b.y));
  r.y = __fmaf_rn(a.x, b.y, __fmul_rn(a.y, b.x));
  return r;
}

// canonical atan2 (double): fixed range reduction + 18-term odd series, explicit fma only
__device__ double canon_atan2(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const double mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
  if (mx == 0.0) return 0.0;
  const double q = __ddiv_rn(mn, mx);
  double t = q, off = 0.0;
  if (q > 0.41421356237309503) {
    t = __ddiv_rn(__dadd_rn(q, -1.0), __dadd_rn(q, 1.0));
    off = 0.78539816339744828;
  }
  const double t2 = __dmul_rn(t, t);
  // 1/(2k+1), k = 0..17: compile-time IEEE quotients, the same doubles the oracle's divisions give
  constexpr double kInvOdd[18] = {1.0 / 1.0,  1.0 / 3.0,  1.0 / 5.0,  1.0 / 7.0,  1.0 / 9.0,  1.0 / 11.0,
                                  1.0 / 13.0, 1.0 / 15.0, 1.0 / 17.0, 1.0 / 19.0, 1.0 / 21.0, 1.0 / 23.0,
                                  1.0 / 25.0, 1.0 / 27.0, 1.0 / 29.0, 1.0 / 31.0, 1.0 / 33.0, 1.0 / 35.0};
  double p = kInvOdd[17];
#pragma unroll
  for (int k = 16; k >= 0; --k) {
    p = -p;
    p = __fma_rn(p, t2, kInvOdd[k]);
  }
  double r = __fma_rn(t, p, off);
  if (ay > ax) r = __dadd_rn(1.5707963267948966, -r);
  if (x < 0.0) r = __dadd_rn(3.1415926535897931, -r);
  if (y < 0.0) r = -r;
  return r;
}

// one input sample as float2, whatever the wire format
template <int FMT>
__device__ __forceinline__ float2 load_in_sample(const char *src, long long idx) {
  if (FMT == LTB_FMT_FC32) return *reinterpret_cast<const float2 *>(src + idx * 8);
  typedef typename fmt_elem<FMT>::type raw_t;
  const raw_t s = *reinterpret_cast<const raw_t *>(src + idx * fmt_bytes(FMT));
  const float k = fmt_scale(FMT);
  return make_float2(__fmul_rn((float)s.x, k), __fmul_rn((float)s.y, k));
}

// ------------------------------------------------------------------------------------
// K0: ingest at D = 1 -- convert (sc16) / copy (fc32) the new chunk into the sample ring
// ------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256) ingest_kernel(const void *__restrict__ in, long long stride_bytes,
                                                     int n_new, float2 *__restrict__ y_ring,
                                                     long long n_base, unsigned cap_mask, int cap, int pairs) {
  const int stream = blockIdx.y;
  const char *src = (const char *)in + (long long)stream * stride_bytes;
  float2 *dst = y_ring + (size_t)stream * cap;
  if (!pairs) {
    // rows that are only sample aligned: one sample per load
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_new; i += gridDim.x * blockDim.x)
      dst[(unsigned)((n_base + i) & cap_mask)] = load_in_sample<FMT>(src, i);
    return;
  }
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n_new; i += gridDim.x * blockDim.x * 2) {
    float4 v;
    if (FMT == LTB_FMT_FC32) {
      v = *reinterpret_cast<const float4 *>(src + (size_t)i * 8);
    } else if (FMT == LTB_FMT_SC16) {
      const short4 s = *reinterpret_cast<const short4 *>(src + (size_t)i * 4);
      const float k = fmt_scale(FMT);
      v = make_float4(__fmul_rn((float)s.x, k), __fmul_rn((float)s.y, k), __fmul_rn((float)s.z, k),
                      __fmul_rn((float)s.w, k));
    } else {
      const char4 s = *reinterpret_cast<const char4 *>(src + (size_t)i * 2);
      const float k = fmt_scale(FMT);
      v = make_float4(__fmul_rn((float)s.x, k), __fmul_rn((float)s.y, k), __fmul_rn((float)s.z, k),
                      __fmul_rn((float)s.w, k));
    }
    *reinterpret_cast<float4 *>(&dst[(unsigned)((n_base + i) & cap_mask)]) = v;
  }
}

// ------------------------------------------------------------------------------------
// K1: polyphase decimator  y[k] = sum_j taps[j] x[kD - j]   (rational_resampler_ccc(1, D))
// canonical order: branch v = j mod D outer, q = j div D inner, one fma chain per component.
// ------------------------------------------------------------------------------------

// Canonical order.  Write the input as aligned blocks X[b][p] = x[b*D + p]
// (p = "position" 0..D-1).  Output k needs, for polyphase branch v = (D - p) % D and tap q,
//   x[kD - qD - v] = X[k - q - (p > 0)][p].
// Every position accumulates  P[p] = fma(taps[qD+v], x[kD-qD-v], P[p])  over q = 0..32 ascending
// (taps beyond ntaps are zeros) in one chain per component; the D partials are then summed by the
// balanced pairwise tree  ((P0+P1)+(P2+P3)) + ((P4+P5)+(P6+P7)) ...
//
// decimate_kernel (D = 2 .. 15): one CTA = 512 outputs of one stream, one warp per position, 16
// consecutive outputs per lane.  The (512+32) blocks x D positions the tile needs are staged once
// in shared memory, row = position, 16-way de-interleaved in the block index so that the element
// a warp needs at one step (block 32 + 16*lane + e) is 32 consecutive float2: a conflict-free
// LDS.64.  The staging copies are 8-byte cp.async (LDGSTS) that write straight into that layout.
// Each lane slides a 16-sample register window down one block per tap: 1 LDS.64 + 1 coefficient
// load per 16 FFMA2, both issued kDecPF steps ahead of their use.  D = 16 has its own kernel
// (decimate_stream_kernel below); the remaining rates run decimate_any_kernel.
constexpr int kDecT = 16;                        // outputs per lane
constexpr int kDecOut = 32 * kDecT;              // 512 outputs per tile
constexpr int kDecQ = 33;                        // taps per polyphase branch (zero padded)
constexpr int kDecGroups = kDecOut + kDecQ - 1;  // 544 blocks per position row
constexpr int kDecSub = kDecGroups / 16;         // 34 columns per sub-row
constexpr int kDecRow = kDecGroups + 1;          // 545 float2: odd stride -> conflict-free fill
constexpr int kDecPF = 3;                        // software prefetch distance (taps)
// (c, c) coefficient pairs [v][33] of the tiled kernel's rates D = 2 .. 15, back to back
__constant__ float2 c_decim_pairs[33 * (15 * 16 / 2 - 1)];
__host__ __device__ constexpr bool decim_is_tiled(int d) { return d >= 2 && d <= 15; }
__host__ __device__ constexpr int decim_pair_offset(int d) { return kDecQ * (d * (d - 1) / 2 - 1); }   // 33 * (2 + .. + d-1)
__host__ __device__ constexpr int decim_pow2(int d) { int p = 1; while (p < d) p <<= 1; return p; }
__host__ __device__ constexpr int decim_min_ctas(int d) { return d <= 4 ? 4 : d <= 12 ? 2 : 1; }
__host__ __device__ constexpr size_t decim_smem_bytes(int d) { return sizeof(float2) * (size_t)d * kDecRow; }

template <int FMT, int D>
__global__ void __launch_bounds__(32 * D, decim_min_ctas(D))
decimate_kernel(const void *__restrict__ in, long long stride_bytes, int n_out, const float2 *__restrict__ tail_in,
                float2 *__restrict__ y_ring, long long n_base, unsigned cap_mask, int cap, int n_streams, int dbg) {
  constexpr int G = D;                            // one warp per position
  constexpr int PW = 1;
  constexpr int NTHR = 32 * G;
  constexpr int POFF = decim_pair_offset(D);
  constexpr int ITER = kDecGroups * D / NTHR;     // 17
  constexpr int BSTEP = NTHR / D;                 // blocks advanced per fill iteration: 32
  static_assert((kDecGroups * D) % NTHR == 0 && NTHR % D == 0, "tile geometry");
  static_assert(decim_is_tiled(D), "D = 16 uses decimate_stream_kernel, other rates decimate_any_kernel");
  extern __shared__ __align__(16) float2 s_x[];   // [D][kDecRow]
  const int stream = blockIdx.y;
  const int k0 = blockIdx.x * kDecOut;
  const char *src = (const char *)in + (long long)stream * stride_bytes;
  const float2 *tail = tail_in + (size_t)stream * kTailCap;
  const long long n_in = (long long)n_out * D;

  // ---- stage blocks k0-33 .. k0+510 (position 0: k0-32 .. k0+511) --------------------------------
  {
    const int p = threadIdx.x % D, c = threadIdx.x / D;
    // block index np = c + BSTEP*it -> offset (np & 15)*kDecSub + (np >> 4), affine in `it`
    float2 *d = s_x + p * kDecRow + (c & 15) * kDecSub + (c >> 4);
    const long long i0 = (long long)D * (k0 - 33 + c) + p + (p == 0 ? D : 0);
    const long long i_first = (long long)D * (k0 - 33), i_last = (long long)D * (k0 + 512);
    if (dbg & 1) {
      // profiling aid: skip the staging copies
    } else if (i_first >= 0 && i_last < n_in) {
      if (FMT == LTB_FMT_FC32) {
        const float2 *gp = reinterpret_cast<const float2 *>(src) + i0;
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
          const int off = (BSTEP == 8) ? (it & 1) * 8 * kDecSub + (it >> 1) : (BSTEP / 16) * it;
          cp_async_8(d + off, gp + (long long)NTHR * it);
        }
        cp_async_wait_all();
      } else {
        typedef typename fmt_elem<FMT>::type raw_t;
        const raw_t *gp = reinterpret_cast<const raw_t *>(src) + i0;
        const float k = fmt_scale(FMT);
#pragma unroll
        for (int b0 = 0; b0 < ITER; b0 += 17) {
          raw_t raw[17];
#pragma unroll
          for (int u = 0; u < 17; ++u) raw[u] = __ldg(gp + (long long)NTHR * (b0 + u));
#pragma unroll
          for (int u = 0; u < 17; ++u) {
            const int it = b0 + u;
            const int off = (BSTEP == 8) ? (it & 1) * 8 * kDecSub + (it >> 1) : (BSTEP / 16) * it;
            d[off] = make_float2(__fmul_rn((float)raw[u].x, k), __fmul_rn((float)raw[u].y, k));
          }
        }
      }
    } else {
      // first / last tile of the chunk: samples before it come from the carried tail, samples
      // past its end do not exist yet (the outputs that would need them are not stored)
#pragma unroll 1
      for (int it = 0; it < ITER; ++it) {
        const long long idx = i0 + (long long)NTHR * it;
        const int off = (BSTEP == 8) ? (it & 1) * 8 * kDecSub + (it >> 1) : (BSTEP / 16) * it;
        float2 val = make_float2(0.f, 0.f);
        if (idx >= 0) { if (idx < n_in) val = load_in_sample<FMT>(src, idx); }
        else if (idx >= -kTailCap) val = tail[kTailCap + idx];
        d[off] = val;
      }
    }
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  float2 acc[kDecT], w[kDecT];
#pragma unroll
  for (int o = 0; o < kDecT; ++o) acc[o] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int h = 0; h < ((dbg & 2) ? 0 : PW); ++h) {
    const int p = g + G * h;
    const int v = (D - p) % D;
    // block 32 + 16*lane + e  ->  sub-row e & 15, column 2 + lane + (e >> 4)
    const float2 *row = s_x + p * kDecRow + 2 + lane;
    const float2 *cf = c_decim_pairs + POFF + v * kDecQ;
#pragma unroll
    for (int o = 0; o < kDecT; ++o) w[o] = row[o * kDecSub];
    float2 pre_x[kDecPF], pre_c[kDecPF];         // elements / coefficients of taps q+1 .. q+PF
#pragma unroll
    for (int j = 0; j < kDecPF; ++j) {
      const int e = -(j + 1);
      pre_x[j] = row[(e & 15) * kDecSub + (e >> 4)];
      pre_c[j] = cf[j];
    }
#pragma unroll
    for (int q = 0; q < kDecQ; ++q) {
      if (q > 0) w[(-q) & 15] = pre_x[(q - 1) % kDecPF];
      const float2 c = pre_c[q % kDecPF];
      if (q > 0 && q - 1 + kDecPF + 1 < kDecQ) {
        const int e = -(q + kDecPF);
        pre_x[(q - 1) % kDecPF] = row[(e & 15) * kDecSub + (e >> 4)];
      }
      if (q + kDecPF < kDecQ) pre_c[q % kDecPF] = cf[q + kDecPF];
#pragma unroll
      for (int o = 0; o < kDecT; ++o) acc[o] = ffma2(c, w[(o - q) & 15], acc[o]);
    }
  }
  if (G == 1) {
#pragma unroll
    for (int o = 0; o < kDecT; ++o) {
      const int k = k0 + lane * kDecT + o;
      if (k < n_out) y_ring[(size_t)stream * cap + (unsigned)((n_base + k) & cap_mask)] = acc[o];
    }
    return;
  }
  // partial sums -> shared memory [g][o][lane] (row stride 33), then the balanced tree
  __syncthreads();
  float2 *part = s_x;
#pragma unroll
  for (int o = 0; o < kDecT; ++o) part[(g * kDecT + o) * 33 + lane] = acc[o];
  __syncthreads();
  for (int i = threadIdx.x; i < kDecOut; i += NTHR) {
    const int o = i & 15, ln = i >> 4;
    constexpr int P2 = decim_pow2(G);                   // D not a power of two: zero partials pad the tree
    float2 pp[P2];
#pragma unroll
    for (int gg = 0; gg < P2; ++gg) pp[gg] = gg < G ? part[(gg * kDecT + o) * 33 + ln] : make_float2(0.f, 0.f);
#pragma unroll
    for (int w2 = 1; w2 < P2; w2 <<= 1) {               // pairwise tree: (P0+P1)+(P2+P3), ...
#pragma unroll
      for (int gg = 0; gg < P2; gg += 2 * w2) pp[gg] = fadd2(pp[gg], pp[gg + w2]);
    }
    const int k = k0 + i;
    if (k < n_out) y_ring[(size_t)stream * cap + (unsigned)((n_base + k) & cap_mask)] = pp[0];
  }
}

// ------------------------------------------------------------------------------------
// K1s: streaming decimator for D = 16 (the rate where the front end dominates the step).
//
// decimate_kernel above stages a transposed tile with 8-byte cp.async; ncu and a dissection
// (copies only: 3-4 TB/s) showed that path cannot feed the FMA pipe at D = 16.  Here the input is
// copied *as it lies in memory* -- one cp.async.bulk (TMA, UBLKCP) of 37 kB per 256 outputs,
// completion on an mbarrier, two buffers per CTA, two persistent 8-warp CTAs per SM -- and the
// layout problem is solved in the thread mapping instead: lane = (position p, half s), so the 16
// lanes of a half-warp read the 16 positions of one input block, 128 contiguous bytes, for every
// tap (conflict-free LDS.64 in natural layout).  Each lane runs the canonical chain of its own
// position for 16 consecutive outputs (sliding 16-sample register window, 1 LDS per 16 FFMA2, its
// 33 taps in registers; FFMA2 takes the tap as a scalar .F32 operand), and the 16 position
// partials of every output are summed in the canonical pairwise tree after a transposition
// through a per-warp shared-memory scratch, which leaves lane l with output 32*warp + l: one
// coalesced store.  The FFMA2 stream saturates register-file read bandwidth (measured,
// tools/ubench_issue.cu: other instructions do not hide in its shadow, they add), so the design
// goal is the fewest non-FFMA2 instructions per 528 FFMA2: 48 LDS + a ~55-instruction reduction.
// Warps per CTA must be a multiple of 4: warp w runs on scheduler w % 4, and every warp does the
// same work, so 6- or 10-warp CTAs overload two of the four schedulers (measured: 4.5 / 4.3 ms
// against 3.6 ms for 8 warps x 2 CTAs or 20 warps x 1 CTA).
// A CTA walks a contiguous run of 256-output segments; the last warp to finish segment i requests
// segment i+2 into the buffer it frees (a shared-memory counter, no CTA barrier), so warps drift
// freely and one segment per CTA is always in flight.
// ------------------------------------------------------------------------------------
constexpr int kStrWarps = 8;
constexpr int kStrThreads = 32 * kStrWarps;
constexpr int kStrSeg = 32 * kStrWarps;               // outputs per segment
constexpr int kStrBlocks = kStrSeg + kDecQ;           // 289 input blocks per segment
constexpr int kStrBufs = 2;
constexpr int kStrScratchRow = 9;                     // float2 per (half, position) row: 8 outputs + 1 pad
constexpr int kStrScratch = 2 * 16 * kStrScratchRow;  // float2 per warp

template <int FMT> __host__ __device__ constexpr int str_buf_bytes() { return kStrBlocks * 16 * fmt_bytes(FMT); }
template <int FMT> __host__ __device__ constexpr size_t decim_stream_smem_bytes() {
  return (size_t)kStrBufs * str_buf_bytes<FMT>() + sizeof(float2) * kStrScratch * kStrWarps;
}

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LTB_WAIT_%=:\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LTB_DONE_%=;\n\t"
      "bra LTB_WAIT_%=;\n\t"
      "LTB_DONE_%=:\n\t"
      "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
               "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes),
               "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// LTB_STREAM_MAXNREG (build option): cap the kernel at that many registers instead of the launch-bounds
// allocation (108 used).  At 96 two CTAs leave room for one track-kernel CTA per SM under the overlapped
// pipeline; measured, the cap costs this kernel 2 % and the co-resident track kernel takes the issue slots
// it needs from this one anyway (5.19 ms per step against 5.08 ms uncapped), so it is off by default.
template <int FMT>
#ifdef LTB_STREAM_MAXNREG
__global__ void __maxnreg__(LTB_STREAM_MAXNREG)
#else
__global__ void __launch_bounds__(kStrThreads, 2)
#endif
decimate_stream_kernel(const void *__restrict__ in, long long stride_bytes, int n_out, const float2 *__restrict__ tail_in,
                       float2 *__restrict__ y_ring, long long n_base, unsigned cap_mask, int cap, int segs_per_stream,
                       int total_segs, int dbg) {
  constexpr int D = 16;
  constexpr int BPS = fmt_bytes(FMT);                              // bytes per input sample
  constexpr int BUF = str_buf_bytes<FMT>();
  static_assert(BUF % 16 == 0, "cp.async.bulk size");
  typedef typename fmt_elem<FMT>::type elem_t;
  extern __shared__ __align__(128) unsigned char s_raw[];          // [kStrBufs][289 blocks][16 positions], scratch
  __shared__ __align__(8) unsigned long long s_full[kStrBufs];
  __shared__ unsigned s_done[kStrBufs];                            // warps finished with a buffer (monotonic)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = lane & 15, half = lane >> 4;
  const long long n_in = (long long)n_out * D;

  const int s_begin = (int)((long long)total_segs * blockIdx.x / gridDim.x);
  const int s_end = (int)((long long)total_segs * (blockIdx.x + 1) / gridDim.x);
  if (s_begin >= s_end) return;

  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < kStrBufs; ++b) { mbar_init(&s_full[b], 1); s_done[b] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  // this lane's taps: position p <-> polyphase branch v = (16 - p) % 16, c[q] = taps[16 q + v]
  // (sc16 / sc8: the 2^-15 / 2^-7 input scale is folded into the taps; both products are exact,
  //  so fma(c * 2^-15, s, acc) == fma(c, s * 2^-15, acc) bit for bit)
  float c[kDecQ];
  {
    const int v = (D - p) % D;
#pragma unroll
    for (int q = 0; q < kDecQ; ++q) {
      const float t = c_decim_taps[decim_tap_offset(D) + q * D + v];   // zero padded beyond ntaps
      c[q] = FMT == LTB_FMT_FC32 ? t : __fmul_rn(t, fmt_scale(FMT));
      // keep the 33 taps in registers: without this the compiler re-reads them from the constant
      // bank inside the FMA loop, and a lane-indexed LDC replays once per distinct address
      asm volatile("" : "+f"(c[q]));
    }
  }

  // segment cursors (stream, k0), advanced incrementally: `cur` is computed on, `req` is the next
  // one to request (kStrBufs ahead)
  struct Seg { int stream, k0; };
  const int k_end = segs_per_stream * kStrSeg;
  auto advance = [&](Seg &S) { S.k0 += kStrSeg; if (S.k0 >= k_end) { S.k0 = 0; S.stream++; } };
  auto is_fast = [&](const Seg &S) { return S.k0 >= kDecQ && (long long)D * (S.k0 + kStrSeg) <= n_in && !(dbg & 1); };
  auto request = [&](const Seg &S, int b) {                        // one thread: TMA for segment S into buffer b
    const char *src = (const char *)in + (long long)S.stream * stride_bytes;
    mbar_expect_tx(&s_full[b], BUF);
    bulk_copy_g2s(s_raw + b * BUF, src + (long long)D * (S.k0 - kDecQ) * BPS, BUF, &s_full[b]);
  };
  Seg cur, req;
  cur.stream = s_begin / segs_per_stream;
  cur.k0 = (s_begin - cur.stream * segs_per_stream) * kStrSeg;
  req = cur;
  for (int j = 0; j < kStrBufs && s_begin + j < s_end; ++j) {
    if (tid == 0 && is_fast(req)) request(req, j);
    advance(req);
  }

  unsigned phase_bits = 0;                                         // parity of each buffer's barrier
  // window element e (= o - q) of this lane sits at block  32*warp + 16*half + 32 + (p == 0) + e
  const int lane_elem = (32 * warp + 16 * half + 32 + (p == 0 ? 1 : 0)) * 16 + p;
  float2 *scratch = reinterpret_cast<float2 *>(s_raw + kStrBufs * BUF) + warp * kStrScratch + half * 16 * kStrScratchRow;

  for (int i = s_begin; i < s_end; ++i) {
    const int b = (i - s_begin) % kStrBufs;
    const Seg S = cur;
    elem_t *buf = reinterpret_cast<elem_t *>(s_raw + b * BUF);
    if (is_fast(S)) {
      mbar_wait(&s_full[b], (phase_bits >> b) & 1u);
      phase_bits ^= 1u << b;
    } else if (!(dbg & 1)) {
      // boundary segment: element-wise, with the carried tail before the chunk and zeros after it
      // (rare: two per stream and call; the only place where the warps of a CTA meet)
      __syncthreads();                                             // every warp has left buffer b
      const char *src = (const char *)in + (long long)S.stream * stride_bytes;
      const long long i_first = (long long)D * (S.k0 - kDecQ);
      const float2 *tail = tail_in + (size_t)S.stream * kTailCap;
      for (int j = tid; j < kStrBlocks * 16; j += kStrThreads) {
        const long long idx = i_first + j;
        float2 val = make_float2(0.f, 0.f);
        if (idx >= 0) { if (idx < n_in) val = load_in_sample<FMT>(src, idx); }
        else if (idx >= -kTailCap) val = tail[kTailCap + idx];
        if (FMT == LTB_FMT_FC32) reinterpret_cast<float2 *>(buf)[j] = val;
        else if (FMT == LTB_FMT_SC16) reinterpret_cast<short2 *>(buf)[j] = make_short2((short)__fmul_rn(val.x, 32768.0f), (short)__fmul_rn(val.y, 32768.0f));
        else reinterpret_cast<char2 *>(buf)[j] = make_char2((signed char)__fmul_rn(val.x, 128.0f), (signed char)__fmul_rn(val.y, 128.0f));
      }
      __syncthreads();
    }

    // Element-major order: window element e (block offset from the lane's base) feeds output o
    // with tap q = o - e.  Walking e downwards keeps every accumulator's chain in ascending q
    // (the canonical order) while the up-to-16 consecutive FFMA2 of one element share their data
    // operand through the register reuse cache: 3 register reads per FFMA2 instead of 4, which is
    // what lets the pipe run at 2 cycles per FFMA2 (tools/ubench_issue.cu).
    float2 acc[kDecT];
    {
      const elem_t *base = buf + lane_elem;
      auto ld = [&](int e) -> float2 {                             // element e of the window
        if (FMT == LTB_FMT_FC32) return reinterpret_cast<const float2 *>(base)[e * 16];
        const elem_t r = base[e * 16];
        return make_float2((float)r.x, (float)r.y);
      };
      constexpr int PF = 4;                                        // elements loaded ahead of their use
      float2 x[PF];
#pragma unroll
      for (int j = 0; j < PF; ++j) x[j] = ld(kDecT - 1 - j);
#pragma unroll
      for (int e = kDecT - 1; e >= -(kDecQ - 1); --e) {
        const int slot = (kDecT - 1 - e) % PF;
        const float2 xe = x[slot];
        if (e - PF >= -(kDecQ - 1)) x[slot] = ld(e - PF);
#pragma unroll
        for (int o = 0; o < kDecT; ++o) {
          const int q = o - e;
          if (q >= 0 && q < kDecQ) {
            const float2 cc = make_float2(c[q], c[q]);
            acc[o] = ffma2(cc, xe, q == 0 ? make_float2(0.f, 0.f) : acc[o]);
          }
        }
      }
    }
    // transpose through the warp's scratch in two passes of 8 outputs (row = (half, position),
    // column = output): lane (g, oo) of a half-warp sums positions 8g..8g+7 of output 8*pass + oo
    // in the canonical pairwise tree, one shuffle adds the two halves of the tree, and the lane
    // whose g equals the pass keeps the result, so lane l ends with output 32*warp + l
    {
      const int g8 = (lane >> 3) & 1, oo = lane & 7;
      float2 *wr = scratch + p * kStrScratchRow;
      const float2 *rd = scratch + (8 * g8) * kStrScratchRow + oo;
      float2 res = make_float2(0.f, 0.f);
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        __syncwarp();
#pragma unroll
        for (int o = 0; o < 8; ++o) wr[o] = acc[8 * pass + o];
        __syncwarp();
        float2 pp[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) pp[r] = rd[r * kStrScratchRow];
#pragma unroll
        for (int w2 = 1; w2 < 8; w2 <<= 1) {
#pragma unroll
          for (int r = 0; r < 8; r += 2 * w2) pp[r] = fadd2(pp[r], pp[r + w2]);
        }
        float2 other;
        other.x = __shfl_xor_sync(0xffffffffu, pp[0].x, 8);
        other.y = __shfl_xor_sync(0xffffffffu, pp[0].y, 8);
        const float2 tot = g8 ? fadd2(other, pp[0]) : fadd2(pp[0], other);   // (P0..7) + (P8..15)
        if (g8 == pass) res = tot;
      }
      const int k = S.k0 + 32 * warp + lane;
      if (k < n_out) y_ring[(size_t)S.stream * cap + (unsigned)((n_base + k) & cap_mask)] = res;
    }
    // release buffer b without a CTA barrier: the last of the warps to get here requests the
    // segment that goes into it next, so no warp ever waits for its siblings
    __syncwarp();
    if (i + kStrBufs < s_end) {
      if (lane == 0 && (atomicAdd(&s_done[b], 1u) % kStrWarps) == kStrWarps - 1 && is_fast(req)) request(req, b);
      advance(req);
    }
    advance(cur);
  }
}

// ------------------------------------------------------------------------------------
// K1s12: the streaming decimator at D = 12..15 (12: 23.04 Msps, the 15 MHz LTE rate).  The D = 16 kernel
// with D of the sixteen lanes of a half-warp at work: blocks are D samples, lanes D..15 run the same
// instruction stream on a window of zeros with zero taps, so their partial sums are exact +0 and the
// 16-row reduction below IS the canonical tree of a non-power-of-two rate (zero partials pad it).
// Up to a quarter of the FFMA2 lanes is idle, which still beats the tiled kernel's shared-memory limit.
// ------------------------------------------------------------------------------------
// blocks copied in front of a segment: 33 of filter history plus what starts the bulk copy on a 16-byte
// boundary and makes its size a multiple of 16 (segments start at multiples of 256 outputs)
template <int FMT, int D> __host__ __device__ constexpr int str12_lead() {
  int l = 33;
  while ((l * D * fmt_bytes(FMT)) % 16 != 0 || ((kStrSeg + l) * D * fmt_bytes(FMT)) % 16 != 0) ++l;
  return l;
}
template <int FMT, int D> __host__ __device__ constexpr size_t decim_stream12_smem_bytes() {
  return (size_t)kStrBufs * (kStrSeg + str12_lead<FMT, D>()) * D * fmt_bytes(FMT) + sizeof(float2) * kStrScratch * kStrWarps +
         ((kDecT + kDecQ + str12_lead<FMT, D>() - 33) * D + 16) * fmt_bytes(FMT);
}

template <int FMT, int D>
__global__ void __launch_bounds__(kStrThreads, 2)
decimate_stream12_kernel(const void *__restrict__ in, long long stride_bytes, int n_out, const float2 *__restrict__ tail_in,
                       float2 *__restrict__ y_ring, long long n_base, unsigned cap_mask, int cap, int segs_per_stream,
                       int total_segs, int dbg) {
  static_assert(D >= 12 && D <= 15, "D = 16: decimate_stream_kernel");
  constexpr int BPS = fmt_bytes(FMT);                              // bytes per input sample
  constexpr int LEAD = str12_lead<FMT, D>();                       // blocks copied in front of the segment
  constexpr int NB = kStrSeg + LEAD;
  constexpr int BUF = NB * D * BPS;
  constexpr int ZEROS = (kDecT + kDecQ + LEAD - 33) * D + 16;      // elements of the idle lanes' zero window
  static_assert(BUF % 16 == 0 && (LEAD * D * BPS) % 16 == 0, "cp.async.bulk size and alignment");
  typedef typename fmt_elem<FMT>::type elem_t;
  extern __shared__ __align__(128) unsigned char s_raw[];          // [kStrBufs][NB blocks][D positions], scratch, zeros
  __shared__ __align__(8) unsigned long long s_full[kStrBufs];
  __shared__ unsigned s_done[kStrBufs];                            // warps finished with a buffer (monotonic)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = lane & 15, half = lane >> 4;
  const bool active = p < D;                                       // lanes D..15 of a half-warp idle on zeros
  const long long n_in = (long long)n_out * D;

  const int s_begin = (int)((long long)total_segs * blockIdx.x / gridDim.x);
  const int s_end = (int)((long long)total_segs * (blockIdx.x + 1) / gridDim.x);
  if (s_begin >= s_end) return;

  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < kStrBufs; ++b) { mbar_init(&s_full[b], 1); s_done[b] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  elem_t *zeros = reinterpret_cast<elem_t *>(s_raw + kStrBufs * BUF + sizeof(float2) * kStrScratch * kStrWarps);
  for (int i = tid; i < ZEROS; i += kStrThreads) memset(&zeros[i], 0, sizeof(elem_t));
  __syncthreads();

  // this lane's taps: position p <-> polyphase branch v = (16 - p) % 16, c[q] = taps[16 q + v]
  // (sc16 / sc8: the 2^-15 / 2^-7 input scale is folded into the taps; both products are exact,
  //  so fma(c * 2^-15, s, acc) == fma(c, s * 2^-15, acc) bit for bit)
  float c[kDecQ];
  {
    const int v = active ? (D - p) % D : 0;
#pragma unroll
    for (int q = 0; q < kDecQ; ++q) {
      const float t = active ? c_decim_taps[decim_tap_offset(D) + q * D + v] : 0.f;   // zero padded beyond ntaps
      c[q] = FMT == LTB_FMT_FC32 ? t : __fmul_rn(t, fmt_scale(FMT));
      // keep the 33 taps in registers: without this the compiler re-reads them from the constant
      // bank inside the FMA loop, and a lane-indexed LDC replays once per distinct address
      asm volatile("" : "+f"(c[q]));
    }
  }

  // segment cursors (stream, k0), advanced incrementally: `cur` is computed on, `req` is the next
  // one to request (kStrBufs ahead)
  struct Seg { int stream, k0; };
  const int k_end = segs_per_stream * kStrSeg;
  auto advance = [&](Seg &S) { S.k0 += kStrSeg; if (S.k0 >= k_end) { S.k0 = 0; S.stream++; } };
  auto is_fast = [&](const Seg &S) { return S.k0 >= LEAD && (long long)D * (S.k0 + kStrSeg) <= n_in && !(dbg & 1); };
  auto request = [&](const Seg &S, int b) {                        // one thread: TMA for segment S into buffer b
    const char *src = (const char *)in + (long long)S.stream * stride_bytes;
    mbar_expect_tx(&s_full[b], BUF);
    bulk_copy_g2s(s_raw + b * BUF, src + (long long)D * (S.k0 - LEAD) * BPS, BUF, &s_full[b]);
  };
  Seg cur, req;
  cur.stream = s_begin / segs_per_stream;
  cur.k0 = (s_begin - cur.stream * segs_per_stream) * kStrSeg;
  req = cur;
  for (int j = 0; j < kStrBufs && s_begin + j < s_end; ++j) {
    if (tid == 0 && is_fast(req)) request(req, j);
    advance(req);
  }

  unsigned phase_bits = 0;                                         // parity of each buffer's barrier
  // window element e (= o - q) of this lane sits at block  32*warp + 16*half + 32 + (p == 0) + e
  const int lane_elem = (32 * warp + 16 * half + LEAD - 1 + (p == 0 ? 1 : 0)) * D + p;
  float2 *scratch = reinterpret_cast<float2 *>(s_raw + kStrBufs * BUF) + warp * kStrScratch + half * 16 * kStrScratchRow;

  for (int i = s_begin; i < s_end; ++i) {
    const int b = (i - s_begin) % kStrBufs;
    const Seg S = cur;
    elem_t *buf = reinterpret_cast<elem_t *>(s_raw + b * BUF);
    if (is_fast(S)) {
      mbar_wait(&s_full[b], (phase_bits >> b) & 1u);
      phase_bits ^= 1u << b;
    } else if (!(dbg & 1)) {
      // boundary segment: element-wise, with the carried tail before the chunk and zeros after it
      // (rare: two per stream and call; the only place where the warps of a CTA meet)
      __syncthreads();                                             // every warp has left buffer b
      const char *src = (const char *)in + (long long)S.stream * stride_bytes;
      const long long i_first = (long long)D * (S.k0 - LEAD);
      const float2 *tail = tail_in + (size_t)S.stream * kTailCap;
      for (int j = tid; j < NB * D; j += kStrThreads) {
        const long long idx = i_first + j;
        float2 val = make_float2(0.f, 0.f);
        if (idx >= 0) { if (idx < n_in) val = load_in_sample<FMT>(src, idx); }
        else if (idx >= -kTailCap) val = tail[kTailCap + idx];
        if (FMT == LTB_FMT_FC32) reinterpret_cast<float2 *>(buf)[j] = val;
        else if (FMT == LTB_FMT_SC16) reinterpret_cast<short2 *>(buf)[j] = make_short2((short)__fmul_rn(val.x, 32768.0f), (short)__fmul_rn(val.y, 32768.0f));
        else reinterpret_cast<char2 *>(buf)[j] = make_char2((signed char)__fmul_rn(val.x, 128.0f), (signed char)__fmul_rn(val.y, 128.0f));
      }
      __syncthreads();
    }

    // Element-major order: window element e (block offset from the lane's base) feeds output o
    // with tap q = o - e.  Walking e downwards keeps every accumulator's chain in ascending q
    // (the canonical order) while the up-to-16 consecutive FFMA2 of one element share their data
    // operand through the register reuse cache: 3 register reads per FFMA2 instead of 4, which is
    // what lets the pipe run at 2 cycles per FFMA2 (tools/ubench_issue.cu).
    float2 acc[kDecT];
    {
      const elem_t *base = active ? buf + lane_elem : zeros + (kDecQ - 1) * D;
      auto ld = [&](int e) -> float2 {                             // element e of the window
        if (FMT == LTB_FMT_FC32) return reinterpret_cast<const float2 *>(base)[e * D];
        const elem_t r = base[e * D];
        return make_float2((float)r.x, (float)r.y);
      };
      constexpr int PF = 4;                                        // elements loaded ahead of their use
      float2 x[PF];
#pragma unroll
      for (int j = 0; j < PF; ++j) x[j] = ld(kDecT - 1 - j);
#pragma unroll
      for (int e = kDecT - 1; e >= -(kDecQ - 1); --e) {
        const int slot = (kDecT - 1 - e) % PF;
        const float2 xe = x[slot];
        if (e - PF >= -(kDecQ - 1)) x[slot] = ld(e - PF);
#pragma unroll
        for (int o = 0; o < kDecT; ++o) {
          const int q = o - e;
          if (q >= 0 && q < kDecQ) {
            const float2 cc = make_float2(c[q], c[q]);
            acc[o] = ffma2(cc, xe, q == 0 ? make_float2(0.f, 0.f) : acc[o]);
          }
        }
      }
    }
    // transpose through the warp's scratch in two passes of 8 outputs (row = (half, position),
    // column = output): lane (g, oo) of a half-warp sums positions 8g..8g+7 of output 8*pass + oo
    // in the canonical pairwise tree, one shuffle adds the two halves of the tree, and the lane
    // whose g equals the pass keeps the result, so lane l ends with output 32*warp + l
    {
      const int g8 = (lane >> 3) & 1, oo = lane & 7;
      float2 *wr = scratch + p * kStrScratchRow;
      const float2 *rd = scratch + (8 * g8) * kStrScratchRow + oo;
      float2 res = make_float2(0.f, 0.f);
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        __syncwarp();
#pragma unroll
        for (int o = 0; o < 8; ++o) wr[o] = acc[8 * pass + o];
        __syncwarp();
        float2 pp[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) pp[r] = rd[r * kStrScratchRow];
#pragma unroll
        for (int w2 = 1; w2 < 8; w2 <<= 1) {
#pragma unroll
          for (int r = 0; r < 8; r += 2 * w2) pp[r] = fadd2(pp[r], pp[r + w2]);
        }
        float2 other;
        other.x = __shfl_xor_sync(0xffffffffu, pp[0].x, 8);
        other.y = __shfl_xor_sync(0xffffffffu, pp[0].y, 8);
        const float2 tot = g8 ? fadd2(other, pp[0]) : fadd2(pp[0], other);   // (P0..7) + (P8..15)
        if (g8 == pass) res = tot;
      }
      const int k = S.k0 + 32 * warp + lane;
      if (k < n_out) y_ring[(size_t)S.stream * cap + (unsigned)((n_base + k) & cap_mask)] = res;
    }
    // release buffer b without a CTA barrier: the last of the warps to get here requests the
    // segment that goes into it next, so no warp ever waits for its siblings
    __syncwarp();
    if (i + kStrBufs < s_end) {
      if (lane == 0 && (atomicAdd(&s_done[b], 1u) % kStrWarps) == kStrWarps - 1 && is_fast(req)) request(req, b);
      advance(req);
    }
    advance(cur);
  }
}

// ------------------------------------------------------------------------------------
// K1s2: the streaming decimator generalised to D = 8 and D = 4 (15.36 and 7.68 Msps), where the
// tiled kernel below reaches half the FFMA2 rate (shared-memory instruction queue, ncu).  Same
// machinery as decimate_stream_kernel -- TMA bulk copies of the input as it lies in memory, two
// buffers, two persistent 8-warp CTAs per SM, taps in registers, element-major FFMA2 order,
// per-warp scratch transposition for the pairwise tree -- with one more index in the lane
// mapping: a half-warp is D positions x R = 16 / D output residues, lane (r, p) runs the canonical
// chain of position p for the 16 outputs k = K0 + R j + r.  The 16 lanes of a half-warp then read
// R consecutive input blocks, 128 contiguous bytes, for every window element (conflict-free
// LDS.64 in natural layout, as at D = 16); the price is 15 R + 33 window elements per lane instead
// of 48, each feeding 33 / R taps.
// ------------------------------------------------------------------------------------
template <int D> __host__ __device__ constexpr int str2_R() { return 16 / D; }
template <int D> __host__ __device__ constexpr int str2_seg() { return kStrWarps * 32 * str2_R<D>(); }   // outputs per segment
// blocks copied in front of the segment: 33 of filter history, plus what it takes to start the
// bulk copy on a 16-byte boundary (segments start at multiples of 256 R outputs)
template <int FMT, int D> __host__ __device__ constexpr int str2_lead() {
  return (33 * D * fmt_bytes(FMT)) % 16 == 0 ? 33 : (34 * D * fmt_bytes(FMT)) % 16 == 0 ? 34 : 36;
}
template <int FMT, int D> __host__ __device__ constexpr int str2_buf_bytes() {
  return (str2_seg<D>() + str2_lead<FMT, D>()) * D * fmt_bytes(FMT);
}
template <int FMT, int D> __host__ __device__ constexpr size_t decim_stream2_smem_bytes() {
  return (size_t)kStrBufs * str2_buf_bytes<FMT, D>() + sizeof(float2) * kStrScratch * kStrWarps;
}

template <int FMT, int D>
__global__ void __launch_bounds__(kStrThreads, 2)
decimate_stream2_kernel(const void *__restrict__ in, long long stride_bytes, int n_out, const float2 *__restrict__ tail_in,
                        float2 *__restrict__ y_ring, long long n_base, unsigned cap_mask, int cap, int segs_per_stream,
                        int total_segs) {
  constexpr int R = str2_R<D>();
  constexpr int SEG = str2_seg<D>();
  constexpr int LEAD = str2_lead<FMT, D>();
  constexpr int NBLK = SEG + LEAD;
  constexpr int BPS = fmt_bytes(FMT);
  constexpr int BUF = str2_buf_bytes<FMT, D>();
  constexpr int NE = 15 * R + kDecQ;                               // window elements per lane
  static_assert(D == 4 || D == 8, "D = 16: decimate_stream_kernel");
  static_assert(BUF % 16 == 0 && (LEAD * D * BPS) % 16 == 0, "cp.async.bulk size and alignment");
  typedef typename fmt_elem<FMT>::type elem_t;
  extern __shared__ __align__(128) unsigned char s_raw[];          // [kStrBufs][NBLK blocks][D positions], scratch
  __shared__ __align__(8) unsigned long long s_full[kStrBufs];
  __shared__ unsigned s_done[kStrBufs];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = lane >> 4, q16 = lane & 15, r = q16 / D, p = q16 % D;
  const long long n_in = (long long)n_out * D;

  const int s_begin = (int)((long long)total_segs * blockIdx.x / gridDim.x);
  const int s_end = (int)((long long)total_segs * (blockIdx.x + 1) / gridDim.x);
  if (s_begin >= s_end) return;
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < kStrBufs; ++b) { mbar_init(&s_full[b], 1); s_done[b] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  float c[kDecQ];                                                  // this lane's taps: branch v = (D - p) % D
  {
    const int v = (D - p) % D;
#pragma unroll
    for (int q = 0; q < kDecQ; ++q) {
      const float t = c_decim_taps[decim_tap_offset(D) + q * D + v];   // zero padded beyond ntaps
      c[q] = FMT == LTB_FMT_FC32 ? t : __fmul_rn(t, fmt_scale(FMT));
      asm volatile("" : "+f"(c[q]));
    }
  }

  struct Seg { int stream, k0; };
  const int k_end = segs_per_stream * SEG;
  auto advance = [&](Seg &S) { S.k0 += SEG; if (S.k0 >= k_end) { S.k0 = 0; S.stream++; } };
  auto is_fast = [&](const Seg &S) { return S.k0 >= LEAD && (long long)D * (S.k0 + SEG) <= n_in; };
  auto request = [&](const Seg &S, int b) {
    const char *src = (const char *)in + (long long)S.stream * stride_bytes;
    mbar_expect_tx(&s_full[b], BUF);
    bulk_copy_g2s(s_raw + b * BUF, src + (long long)D * (S.k0 - LEAD) * BPS, BUF, &s_full[b]);
  };
  Seg cur, req;
  cur.stream = s_begin / segs_per_stream;
  cur.k0 = (s_begin - cur.stream * segs_per_stream) * SEG;
  req = cur;
  for (int j = 0; j < kStrBufs && s_begin + j < s_end; ++j) {
    if (tid == 0 && is_fast(req)) request(req, j);
    advance(req);
  }

  unsigned phase_bits = 0;
  // window element e of this lane: block K0 + r - (p > 0) + e, K0 = k0 + (2 warp + half) 16 R
  const int sub0 = (2 * warp + half) * 16 * R;                     // first output of this half-warp in the segment
  const int lane_elem = (sub0 + r - (p > 0 ? 1 : 0) + LEAD) * D + p;
  float2 *scratch = reinterpret_cast<float2 *>(s_raw + kStrBufs * BUF) + warp * kStrScratch + half * 16 * kStrScratchRow;

  for (int i = s_begin; i < s_end; ++i) {
    const int b = (i - s_begin) % kStrBufs;
    const Seg S = cur;
    elem_t *buf = reinterpret_cast<elem_t *>(s_raw + b * BUF);
    if (is_fast(S)) {
      mbar_wait(&s_full[b], (phase_bits >> b) & 1u);
      phase_bits ^= 1u << b;
    } else {
      // boundary segment: element-wise, with the carried tail before the chunk and zeros after it
      __syncthreads();
      const char *src = (const char *)in + (long long)S.stream * stride_bytes;
      const long long i_first = (long long)D * (S.k0 - LEAD);
      const float2 *tail = tail_in + (size_t)S.stream * kTailCap;
      for (int j = tid; j < NBLK * D; j += kStrThreads) {
        const long long idx = i_first + j;
        float2 val = make_float2(0.f, 0.f);
        if (idx >= 0) { if (idx < n_in) val = load_in_sample<FMT>(src, idx); }
        else if (idx >= -kTailCap) val = tail[kTailCap + idx];
        if (FMT == LTB_FMT_FC32) reinterpret_cast<float2 *>(buf)[j] = val;
        else if (FMT == LTB_FMT_SC16) reinterpret_cast<short2 *>(buf)[j] = make_short2((short)__fmul_rn(val.x, 32768.0f), (short)__fmul_rn(val.y, 32768.0f));
        else reinterpret_cast<char2 *>(buf)[j] = make_char2((signed char)__fmul_rn(val.x, 128.0f), (signed char)__fmul_rn(val.y, 128.0f));
      }
      __syncthreads();
    }

    // element-major: element e feeds output j with tap q = R j - e; walking e downwards keeps every
    // accumulator's chain in ascending q (the canonical order)
    float2 acc[kDecT];
    {
      const elem_t *base = buf + lane_elem;
      auto ld = [&](int e) -> float2 {
        const elem_t v = base[e * D];
        return make_float2((float)v.x, (float)v.y);
      };
      constexpr int PF = 4;
      constexpr int E_TOP = 15 * R, E_BOT = -(kDecQ - 1);
      float2 x[PF];
#pragma unroll
      for (int j = 0; j < PF; ++j) x[j] = ld(E_TOP - j);
#pragma unroll
      for (int e = E_TOP; e >= E_BOT; --e) {
        const int slot = (E_TOP - e) % PF;
        const float2 xe = x[slot];
        if (e - PF >= E_BOT) x[slot] = ld(e - PF);
#pragma unroll
        for (int j = 0; j < kDecT; ++j) {
          const int q = R * j - e;
          if (q >= 0 && q < kDecQ) {
            const float2 cc = make_float2(c[q], c[q]);
            acc[j] = ffma2(cc, xe, q == 0 ? make_float2(0.f, 0.f) : acc[j]);
          }
        }
      }
      (void)NE;
    }
    // transposition through the warp's scratch, two passes of 8 outputs per lane: row = (r, p),
    // column = j.  Lane (a, jj) of a half-warp then sums the D positions of R / 2 outputs
    // (residues r' = a R / 2 + u) in the canonical pairwise tree.
    {
      const int a = q16 >> 3, jj = q16 & 7;
      float2 *wr = scratch + q16 * kStrScratchRow;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        __syncwarp();
#pragma unroll
        for (int o = 0; o < 8; ++o) wr[o] = acc[8 * pass + o];
        __syncwarp();
        float2 res[R / 2];
#pragma unroll
        for (int u = 0; u < R / 2; ++u) {
          const int rr = a * (R / 2) + u;
          float2 pp[D];
#pragma unroll
          for (int t = 0; t < D; ++t) pp[t] = scratch[(rr * D + t) * kStrScratchRow + jj];
#pragma unroll
          for (int w2 = 1; w2 < D; w2 <<= 1) {
#pragma unroll
            for (int t = 0; t < D; t += 2 * w2) pp[t] = fadd2(pp[t], pp[t + w2]);
          }
          res[u] = pp[0];
        }
        // outputs K0 + R (8 pass + jj) + a R / 2 + u, u = 0 .. R/2 - 1: consecutive
        const int k = S.k0 + sub0 + R * (8 * pass + jj) + a * (R / 2);
        float2 *dst = y_ring + (size_t)S.stream * cap;
        if (R == 2) {
          if (k < n_out) dst[(unsigned)((n_base + k) & cap_mask)] = res[0];
        } else {
          // two consecutive outputs, k even: one 16-byte store (n_out, n_base and cap are multiples of 8)
          if (k < n_out)
            *reinterpret_cast<float4 *>(&dst[(unsigned)((n_base + k) & cap_mask)]) = make_float4(res[0].x, res[0].y, res[R / 2 - 1].x, res[R / 2 - 1].y);
        }
      }
    }
    __syncwarp();
    if (i + kStrBufs < s_end) {
      if (lane == 0 && (atomicAdd(&s_done[b], 1u) % kStrWarps) == kStrWarps - 1 && is_fast(req)) request(req, b);
      advance(req);
    }
    advance(cur);
  }
}

// ------------------------------------------------------------------------------------
// K1a: decimator for every other integer rate (D = 5, 7, 9..11, 13..15, 17..64): the reference
// accepts any multiple of 1.92 Msps (examples/cell_search_file.py:50-57).  One CTA = 256 outputs
// of one stream, one output per thread.  The 289 input blocks are staged transposed (row =
// position, odd row stride) so that the lanes of a warp read consecutive float2 for every tap;
// the branch taps [v][33] come from a per-rate table in global memory and sit in shared memory
// (broadcast reads).  Same canonical order: one fma chain per position over q ascending, then the
// balanced pairwise tree over the partials padded with zeros to a power of two, evaluated with a
// binary-counter stack so that no partial has to be kept.
// ------------------------------------------------------------------------------------
constexpr int kAnyOut = 256;
constexpr int kAnyBlocks = kAnyOut + kDecQ;           // 289
__host__ __device__ constexpr size_t decim_any_smem_bytes(int d) {
  return sizeof(float2) * (size_t)d * kAnyBlocks + sizeof(float) * (size_t)d * kDecQ;
}

template <int FMT>
__global__ void __launch_bounds__(kAnyOut)
decimate_any_kernel(const void *__restrict__ in, long long stride_bytes, int n_out, int D, const float *__restrict__ taps_vq,
                    const float2 *__restrict__ tail_in, float2 *__restrict__ y_ring, long long n_base,
                    unsigned cap_mask, int cap) {
  extern __shared__ __align__(16) float2 s_any[];     // [D][289] samples, then [D][33] taps
  float *s_taps = reinterpret_cast<float *>(s_any + (size_t)D * kAnyBlocks);
  const int stream = blockIdx.y, t = threadIdx.x;
  const int k0 = blockIdx.x * kAnyOut;
  const char *src = (const char *)in + (long long)stream * stride_bytes;
  const float2 *tail = tail_in + (size_t)stream * kTailCap;
  const long long n_in = (long long)n_out * D;
  const long long i_first = (long long)D * (k0 - kDecQ);
  for (int j = t; j < D * kDecQ; j += kAnyOut) s_taps[j] = taps_vq[j];
  for (int j = t; j < D * kAnyBlocks; j += kAnyOut) {
    const long long idx = i_first + j;
    float2 val = make_float2(0.f, 0.f);
    if (idx >= 0) { if (idx < n_in) val = load_in_sample<FMT>(src, idx); }
    else if (idx >= -kTailCap) val = tail[kTailCap + idx];
    const int b = j / D, p = j - b * D;
    s_any[p * kAnyBlocks + b] = val;
  }
  __syncthreads();
  int P2 = 1;
  while (P2 < D) P2 <<= 1;
  float2 stk[7], val = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int p = 0; p < P2; ++p) {
    val = make_float2(0.f, 0.f);
    if (p < D) {
      // X[k - q - (p > 0)][p], block index relative to k0 - 33
      const float2 *row = s_any + p * kAnyBlocks + t + kDecQ - (p > 0 ? 1 : 0);
      const float *cf = s_taps + ((D - p) % D) * kDecQ;
#pragma unroll
      for (int q = 0; q < kDecQ; ++q) {
        const float c = cf[q];
        val = ffma2(make_float2(c, c), row[-q], val);
      }
    }
    bool merging = true;                              // (P0+P1)+(P2+P3) ... : carry chain of p + 1
#pragma unroll
    for (int l = 0; l < 7; ++l) {
      if (merging) {
        if ((p >> l) & 1) val = fadd2(stk[l], val);
        else { stk[l] = val; merging = false; }
      }
    }
  }
  const int k = k0 + t;
  if (k < n_out) y_ring[(size_t)stream * cap + (unsigned)((n_base + k) & cap_mask)] = val;
}

// keep the last kTailCap converted input samples of each stream for the next call
template <int FMT>
__global__ void __launch_bounds__(256) tail_kernel(const void *__restrict__ in, long long stride_bytes,
                                                   long long n_in, const float2 *__restrict__ tail_old,
                                                   float2 *__restrict__ tail_new) {
  const int stream = blockIdx.x;
  const char *src = (const char *)in + (long long)stream * stride_bytes;
  for (int i = threadIdx.x; i < kTailCap; i += blockDim.x) {
    const long long idx = n_in - kTailCap + i;      // index into the new chunk (may be negative)
    float2 v;
    if (idx >= 0) v = load_in_sample<FMT>(src, idx);
    else v = (idx >= -kTailCap) ? tail_old[(size_t)stream * kTailCap + kTailCap + idx] : make_float2(0.f, 0.f);
    tail_new[(size_t)stream * kTailCap + i] = v;
  }
}

// ------------------------------------------------------------------------------------
// K2: three-root matched filter + |.|^2 over the new samples of every stream
//
// Folded direct form.  h[128-m] == h[m] and h_34 == conj(h_29), so per output n
//   s_0 = x[n],  s_m = x[n-m] + x[n-128+m] (m=1..63),  s_64 = x[n-64]
//   (A,B)_g += (hr,hi)_g[m] * (s.re, s.im)      (D,C)_g += (hi,hr)_g[m] * (s.re, s.im)
// for g = root 25, root 29: 1 FADD2 + 4 FFMA2 per folded tap instead of 12 FFMA per tap pair.
//   root 25 / 29: y = (A-B, C+D)     root 34: y = (A+B, C-D)     P = fma(re, re, im*im)
// Each thread owns 8 consecutive outputs and slides two 8-sample register windows over a
// shared-memory tile stored 8-way de-interleaved (row = index mod 8), which makes every
// LDS.64 of a warp hit 32 consecutive 8-byte words: no bank conflicts, 2 LDS per 40 FP ops.
// ------------------------------------------------------------------------------------
constexpr int kCorrT = 8;
constexpr int kCorrThreads = 256;
constexpr int kCorrTile = kCorrT * kCorrThreads;          // 2048 outputs per CTA
constexpr int kCorrCols = (kCorrTile + 128) / 8;          // 272
constexpr int kCorrRowStride = kCorrCols + 2;             // 274 float2 = 548 words == 4 (mod 16)

__device__ __forceinline__ float2 corr_lds(const float2 *row_base, int e) {
  // tile sample index e (compile-time after unrolling, relative to 8*t) -> row e&7, column t + (e>>3)
  return row_base[(e & 7) * kCorrRowStride + (e >> 3)];
}

// Persistent: a CTA walks a contiguous run of (stream, tile) pairs with two tile buffers; the
// 8-byte cp.async copies of tile i+1 (straight into the de-interleaved layout) are issued before
// the arithmetic of tile i, so staging latency -- 38 % of the warp time in the one-tile-per-CTA
// version (ncu) -- is hidden behind ~2600 FP instructions per thread.
__global__ void __launch_bounds__(kCorrThreads, 2)
pss_corr_kernel(const float2 *__restrict__ y_ring, float *__restrict__ p_ring, long long n_base, int n_new,
                unsigned cap_mask, int cap, int tiles_per_stream, int total_tiles) {
  __shared__ float2 sx2[2][8 * kCorrRowStride];
  const int t_begin = (int)((long long)total_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)total_tiles * (blockIdx.x + 1) / gridDim.x);
  // stage samples n0-128 .. n0+2048 of tile `tile` (tile index u = n - n0 + 128)
  auto stage = [&](int tile, float2 *dst) {
    const int st = tile / tiles_per_stream;
    const long long n0 = n_base + (long long)(tile - st * tiles_per_stream) * kCorrTile;
    const float2 *yr = y_ring + (size_t)st * cap;
#pragma unroll
    for (int j = 0; j < (kCorrTile + 128 + kCorrThreads - 1) / kCorrThreads; ++j) {
      const int u = threadIdx.x + j * kCorrThreads;
      if (u < kCorrTile + 128) cp_async_8(&dst[(u & 7) * kCorrRowStride + (u >> 3)], &yr[(unsigned)((n0 - 128 + u) & cap_mask)]);
    }
  };
  if (t_begin < t_end) stage(t_begin, sx2[0]);
  for (int tile = t_begin; tile < t_end; ++tile) {
  const float2 *sx = sx2[(tile - t_begin) & 1];
  const int stream = tile / tiles_per_stream;
  const int tile_x = tile - stream * tiles_per_stream;
  const long long n0 = n_base + (long long)tile_x * kCorrTile;   // absolute index of output 0
  cp_async_wait_all();
  __syncthreads();                       // tile landed; everyone is done with the other buffer
  if (tile + 1 < t_end) stage(tile + 1, sx2[(tile - t_begin + 1) & 1]);

  const int t = threadIdx.x;
  const float2 *rb = sx + t;          // column offset t; corr_lds adds row and the constant column
  float2 wa[8], wb[8];                // sliding windows, tile element e lives in slot e & 7
  float2 ab[8][2], dc[8][2];
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    ab[o][0] = ab[o][1] = dc[o][0] = dc[o][1] = make_float2(0.f, 0.f);
  }
  // m = 0 : s = x[n] -> tile element 8t + 128 + o
#pragma unroll
  for (int o = 0; o < 8; ++o) wa[o] = corr_lds(rb, 128 + o);
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const float2 s = wa[o];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      ab[o][g] = ffma2(c_pss_coef[g][0][0], s, ab[o][g]);
      dc[o][g] = ffma2(c_pss_coef[g][0][1], s, dc[o][g]);
    }
  }
  // b-window for m = 1: tile elements 8t + 1 .. 8t + 7 (element 8t + 8 arrives with step m = 1)
#pragma unroll
  for (int o = 1; o < 8; ++o) wb[o] = corr_lds(rb, o);

  // step m = 8*mb + m1 (m1 = 1..8): a-window slides down (new element 128-m: row (8-m1)&7,
  // column 15-mb), b-window slides up (new element m+7: row m1-1, column mb+1).  The slot
  // arithmetic only depends on m1, so an 8-step body is fully static and is rolled 7 times.
#define LTB_CORR_STEP(M1, PA, PB, COEF, WITH_B)                                        \
  {                                                                                    \
    wa[(8 - (M1)) & 7] = (PA)[((8 - (M1)) & 7) * kCorrRowStride];                      \
    if (WITH_B) wb[((M1) - 1) & 7] = (PB)[(((M1) - 1) & 7) * kCorrRowStride];          \
    const float2 c00 = (COEF)[0][0], c01 = (COEF)[0][1];                               \
    const float2 c10 = (COEF)[65][0], c11 = (COEF)[65][1];                         \
    _Pragma("unroll") for (int o = 0; o < 8; ++o) {                                    \
      const float2 s = (WITH_B) ? fadd2(wa[(o - (M1)) & 7], wb[(o + (M1)) & 7])        \
                                : wa[(o - (M1)) & 7];                                  \
      ab[o][0] = ffma2(c00, s, ab[o][0]);                                              \
      dc[o][0] = ffma2(c01, s, dc[o][0]);                                              \
      ab[o][1] = ffma2(c10, s, ab[o][1]);                                              \
      dc[o][1] = ffma2(c11, s, dc[o][1]);                                              \
    }                                                                                  \
  }
  typedef float2 coef_pair_t[2];
  const float2 *pa = rb + 15, *pb = rb + 1;
#pragma unroll 1
  for (int mb = 0; mb < 7; ++mb) {
    const coef_pair_t *cf = &c_pss_coef[0][8 * mb];
#pragma unroll
    for (int m1 = 1; m1 <= 8; ++m1) LTB_CORR_STEP(m1, pa, pb, cf + m1, true)
    pa -= 1;
    pb += 1;
  }
  {
    const coef_pair_t *cf = &c_pss_coef[0][56];
#pragma unroll
    for (int m1 = 1; m1 <= 7; ++m1) LTB_CORR_STEP(m1, pa, pb, cf + m1, true)
    LTB_CORR_STEP(8, pa, pb, cf + 8, false)      // m = 64 : s = x[n-64], no partner
  }
#undef LTB_CORR_STEP

  const int local = t * 8;
  if ((long long)tile_x * kCorrTile + local >= n_new) continue;   // n_new is a multiple of 8
  float p0[8], p1[8], p2[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    // ab = (A, B), dc = (D, C)
    float re = __fadd_rn(ab[o][0].x, -ab[o][0].y), im = __fadd_rn(dc[o][0].y, dc[o][0].x);
    p0[o] = __fmaf_rn(re, re, __fmul_rn(im, im));
    re = __fadd_rn(ab[o][1].x, -ab[o][1].y); im = __fadd_rn(dc[o][1].y, dc[o][1].x);
    p1[o] = __fmaf_rn(re, re, __fmul_rn(im, im));
    re = __fadd_rn(ab[o][1].x, ab[o][1].y); im = __fadd_rn(dc[o][1].y, -dc[o][1].x);
    p2[o] = __fmaf_rn(re, re, __fmul_rn(im, im));
  }
  const unsigned slot = (unsigned)((n0 + local) & cap_mask);
  float *pp = p_ring + (size_t)stream * 3 * cap + slot;
  reinterpret_cast<float4 *>(pp)[0] = make_float4(p0[0], p0[1], p0[2], p0[3]);
  reinterpret_cast<float4 *>(pp)[1] = make_float4(p0[4], p0[5], p0[6], p0[7]);
  reinterpret_cast<float4 *>(pp + cap)[0] = make_float4(p1[0], p1[1], p1[2], p1[3]);
  reinterpret_cast<float4 *>(pp + cap)[1] = make_float4(p1[4], p1[5], p1[6], p1[7]);
  reinterpret_cast<float4 *>(pp + 2 * (size_t)cap)[0] = make_float4(p2[0], p2[1], p2[2], p2[3]);
  reinterpret_cast<float4 *>(pp + 2 * (size_t)cap)[1] = make_float4(p2[4], p2[5], p2[6], p2[7]);
  }
}

// ------------------------------------------------------------------------------------
// K2f: the same three-root matched filter as overlap-save FFT blocks (corr_mode LTB_CORR_FFT).
//
// Block b of a stream = 1024 samples x[896 b - 128 .. 896 b + 896) -> the 896 correlation powers
// P_g[896 b .. 896 b + 896) of the three roots: X = FFT(x), Y_g = H_g X, y_g = IFFT(Y_g),
// P = |y_g[n]|^2 for n >= 128.  Blocks are aligned to absolute sample indices and only whole
// blocks are evaluated (the tail waits for the next call), so results do not depend on how the
// stream is cut into calls; a window's interior lags end 8765 samples before the data the
// scheduler rule requires, so the waiting tail is never needed.
//
// One warp per block, everything in registers: lane l holds 32 complex values.  FFT_1024 is a
// four-step transform (n = 32 n1 + n2, f = k1 + 32 k2): 32-point radix-2 DIF over the registers
// with W_32 twiddles as immediates, one table twiddle W_1024^(n2 k1) per value, a 32 x 32
// transpose through 8.4 kB of shared memory, another 32-point DIF.  The inverse runs the
// mirrored graph (DIT, conjugate table twiddle, transpose, DIF) three times on H_g X with X kept
// in registers.  Per lane and block: about 4700 FP32 operations against 20700 lane-cycles of the
// folded direct form for the same 896 x 3 outputs.  The arithmetic (every butterfly, the
// canonical complex product, exact shortcuts for w = 1, -+j) is the oracle's ORC_CONV_OS.
// ------------------------------------------------------------------------------------
constexpr int kOsStep = 896;
constexpr int kOsWarps = 4;

__device__ __forceinline__ float2 w32_const(int i) {
  switch (i) {
    case 1: return make_float2(9.807852507e-01f, -1.950903237e-01f);
    case 2: return make_float2(9.238795042e-01f, -3.826834261e-01f);
    case 3: return make_float2(8.314695954e-01f, -5.555702448e-01f);
    case 4: return make_float2(7.071067691e-01f, -7.071067691e-01f);
    case 5: return make_float2(5.555702448e-01f, -8.314695954e-01f);
    case 6: return make_float2(3.826834261e-01f, -9.238795042e-01f);
    case 7: return make_float2(1.950903237e-01f, -9.807852507e-01f);
    case 9: return make_float2(-1.950903237e-01f, -9.807852507e-01f);
    case 10: return make_float2(-3.826834261e-01f, -9.238795042e-01f);
    case 11: return make_float2(-5.555702448e-01f, -8.314695954e-01f);
    case 12: return make_float2(-7.071067691e-01f, -7.071067691e-01f);
    case 13: return make_float2(-8.314695954e-01f, -5.555702448e-01f);
    case 14: return make_float2(-9.238795042e-01f, -3.826834261e-01f);
    case 15: return make_float2(-9.807852507e-01f, -1.950903237e-01f);
    default: return make_float2(1.f, 0.f);
  }
}

// W_32^idx (conjugated if INV) times d; idx is a compile-time constant after unrolling
template <bool INV>
__device__ __forceinline__ float2 tw32_mul(int idx, float2 d) {
  if (idx == 0) return d;
  if (idx == 8) return INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
  float2 w = w32_const(idx);
  if (INV) w.y = -w.y;
  return cmul_canon(w, d);
}

// radix-2 DIF over the 32 registers: natural order in, bit-reversed order out.  One stage = 16
// butterflies in a flat loop of constant trip count (nested loops with stage-dependent bounds
// were left rolled by the compiler, which put the array in local memory).
template <bool INV, int S>
__device__ __forceinline__ void fft32_dif_stage(float2 (&v)[32]) {
  constexpr int half = 16 >> S;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const int k = t % half, i = (t / half) * 2 * half + k, j = i + half;
    // packed add; a - b as fma(b, -1, a), which rounds exactly like the subtraction: one issue
    // slot per complex add instead of two (the kernel is issue bound)
    const float2 a = v[i], b = v[j];
    v[i] = fadd2(a, b);
    v[j] = tw32_mul<INV>(k << S, ffma2(b, make_float2(-1.f, -1.f), a));
  }
}
template <bool INV>
__device__ __forceinline__ void fft32_dif(float2 (&v)[32]) {
  fft32_dif_stage<INV, 0>(v); fft32_dif_stage<INV, 1>(v); fft32_dif_stage<INV, 2>(v);
  fft32_dif_stage<INV, 3>(v); fft32_dif_stage<INV, 4>(v);
}

// radix-2 DIT: bit-reversed order in, natural order out
template <bool INV, int S>
__device__ __forceinline__ void fft32_dit_stage(float2 (&v)[32]) {
  constexpr int half = 1 << S;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const int k = t % half, i = (t / half) * 2 * half + k, j = i + half;
    const float2 w = tw32_mul<INV>(k * (16 >> S), v[j]);
    const float2 a = v[i];
    v[i] = fadd2(a, w);
    v[j] = ffma2(w, make_float2(-1.f, -1.f), a);
  }
}
template <bool INV>
__device__ __forceinline__ void fft32_dit(float2 (&v)[32]) {
  fft32_dit_stage<INV, 0>(v); fft32_dit_stage<INV, 1>(v); fft32_dit_stage<INV, 2>(v);
  fft32_dit_stage<INV, 3>(v); fft32_dit_stage<INV, 4>(v);
}

__host__ __device__ constexpr int bitrev5(int v) {
  return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}

// tw_perm[r][l] = W_1024^(l * bitrev5(r));  h_perm[g][r][l] = 2^-10 DFT_1024(h_g)[l + 32 bitrev5(r)]
//
// The work of one block is a loop over five stages that share ONE copy of the 32-point DIF code
// (and one of the DIT): stage 0 loads x and ends in the twiddle + transpose, stage 1 finishes the
// forward transform and parks X in shared memory, stages 2..4 are the three roots.  The inverse
// DIF with conjugated twiddles is evaluated as conj(DIF(conj(.))), which is exact (negation
// commutes with rounding): the conjugation of its input is folded into the table-twiddle
// product, and the conjugation of its output does not change |y|^2.  A first version with four
// inlined FFT bodies and X in registers spilled 400 B per thread to local memory and stalled on
// instruction fetch (ncu: long_scoreboard 3.8, no_instruction 1.9 warps per issue).
#ifndef LTB_OS_MIN_CTAS
#define LTB_OS_MIN_CTAS 3
#endif
struct OsShared {
  float2 tw[32 * 32];
  float2 tr[kOsWarps][32 * 33];
  float2 X[kOsWarps][32 * 32];
};

__global__ void __launch_bounds__(32 * kOsWarps, LTB_OS_MIN_CTAS)
pss_corr_fft_kernel(const float2 *__restrict__ y_ring, float *__restrict__ p_ring, long long blk_first, int blk_count,
                    unsigned cap_mask, int cap, int n_streams, const float2 *__restrict__ tw_perm,
                    const float2 *__restrict__ h_perm) {
  extern __shared__ __align__(16) unsigned char os_raw[];
  OsShared &S = *reinterpret_cast<OsShared *>(os_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 1024; i += 32 * kOsWarps) S.tw[i] = tw_perm[i];
  __syncthreads();
  float2 *tr = S.tr[warp];
  float2 *Xs = S.X[warp] + lane;
  const float2 *twl = S.tw + lane;
  const long long total = (long long)blk_count * n_streams;
  for (long long w = (long long)blockIdx.x * kOsWarps + warp; w < total; w += (long long)gridDim.x * kOsWarps) {
    const int stream = (int)(w / blk_count);
    const long long b = blk_first + (w - (long long)stream * blk_count);
    const long long base = b * kOsStep - 128;                       // absolute index of block sample 0
    const float2 *yr = y_ring + (size_t)stream * cap;
    float2 v[32];
#pragma unroll 1
    for (int stage = 0; stage < 5; ++stage) {
      if (stage == 0) {
        // lane = n2, j = n1; base is a multiple of 128 and so is the ring size: a 128-sample chunk
        // never straddles the ring end, one masked offset serves four loads
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float2 *src = yr + (unsigned)((base + 128 * q) & cap_mask) + lane;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) v[4 * q + jj] = src[32 * jj];
        }
      } else if (stage >= 2) {
        const float2 *hp = h_perm + (size_t)(stage - 2) * 1024 + lane;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          v[r] = cmul_canon(__ldg(hp + r * 32), Xs[r * 32]);
          if ((r & 7) == 7) asm volatile("" ::: "memory");          // at most 8 + 8 loads in flight: no spills
        }
        fft32_dit<true>(v);                                         // over k2: v[c], c = n2, lane = k1
        // conj(W_1024^(lane c)) v[c], conjugated: the DIF below then yields conj(y)
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float2 t = twl[bitrev5(c) * 32];
          const float2 d = v[c];
          v[c].x = __fmaf_rn(t.x, d.x, __fmul_rn(t.y, d.y));
          v[c].y = __fmaf_rn(t.x, -d.y, __fmul_rn(t.y, d.x));
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 32; ++c) tr[c * 33 + lane] = v[c];
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = tr[lane * 33 + c];      // lane = n2, c = k1
      }
      fft32_dif<false>(v);
      if (stage == 0) {                                             // v[r]: k1 = bitrev5(r), lane = n2
#pragma unroll
        for (int r = 0; r < 32; ++r) v[r] = cmul_canon(twl[r * 32], v[r]);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; ++r) tr[bitrev5(r) * 33 + lane] = v[r];
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = tr[lane * 33 + c];      // lane = k1, c = n2
      } else if (stage == 1) {                                      // v[r]: f = lane + 32 bitrev5(r)
#pragma unroll
        for (int r = 0; r < 32; ++r) Xs[r * 32] = v[r];
        __syncwarp();
      } else {                                                      // v[r] = conj(y[32 bitrev5(r) + lane])
        float *pr = p_ring + ((size_t)stream * 3 + (stage - 2)) * cap + lane;
#pragma unroll
        for (int q = 1; q < 8; ++q) {
          float *dst = pr + (unsigned)((base + 128 * q) & cap_mask);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float2 y = v[bitrev5(4 * q + jj)];                // n1 = 4 q + jj
            dst[32 * jj] = __fmaf_rn(y.x, y.x, __fmul_rn(y.y, y.y));
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// K3: per-chain sequential phase = pss::general_work + the head of sss::work
//     (lib/pss_impl.cc:154-223, lib/sss_impl.cc:83-110), one CTA per (stream, N_id_2) chain.
// ------------------------------------------------------------------------------------
constexpr int kTrackThreads = 256;
constexpr int kNEdge = 253;            // lags 0..126 and 9600..9725 see a truncated window

// srslte_cexptab_gen's phase recurrence  phase = fl(phase + inc)  (with the 4096 wraps) is
// sequential, but inside one float binade it is exactly linear in the bit pattern: a grid point
// plus a constant rounds to "bits + dk" with a constant dk (after one in-binade step fixes the
// ties-to-even parity).  Thread 0 therefore only walks binade/wrap boundaries with real float
// adds and emits (start, bits0, dk, len) segments; all threads expand them in parallel.  The
// result is bit-identical to the sample-by-sample loop (checked against the oracle, which runs
// that loop).
struct PhaseSeg { int start; unsigned bits0; int dk; int len; };
constexpr int kMaxSeg = 64;
constexpr int kSegPingPong = (int)0x80000000u;   // PhaseSeg::dk marker: indices alternate 0, 4096, 0, ... from the segment's start

__device__ int plan_phase_scan(float &phase, bool &stable, float inc, int n, PhaseSeg *segs,
                               unsigned short *idx_out) {
  int i = 0, ns = 0;
  while (i < n) {
    const float before = phase;
    while (phase >= 4096.0f) phase = __fadd_rn(phase, -4096.0f);
    while (phase < 0.f) phase = __fadd_rn(phase, 4096.0f);
    if (phase != before) stable = false;
    if (inc == 0.f && ns < kMaxSeg) {
      // no frequency offset at all (a noiseless capture at its native rate measures exactly zero): the phase never
      // moves, which the binade logic below would walk one sample at a time (phase 0 is in no binade) -- 960 serial
      // steps per window on thread 0, ten times the cost of the whole search (profiles/ncu_track_c2_r02.txt)
      segs[ns].start = i; segs[ns].bits0 = __float_as_uint(phase); segs[ns].dk = 0; segs[ns].len = n - i;
      return ns + 1;
    }
    if (inc < 0.f && phase == 0.f && ns < kMaxSeg && __fadd_rn(4096.0f, inc) == 4096.0f) {
      // an offset below zero and smaller than half an ulp of 4096 (|f| < 6e-8 cycles per sample: a noiseless capture
      // at its native rate, e.g. the 6 PRB fixture): the recurrence ping-pongs 0 -> inc -> (+4096, rounds to) 4096 ->
      // (-4096) 0 ..., indices 0, 4096, 0, 4096, ...; the binade logic below would take it one sample at a time
      segs[ns].start = i; segs[ns].bits0 = 0u; segs[ns].dk = kSegPingPong; segs[ns].len = n - i;
      phase = ((n - i) & 1) ? inc : 4096.0f;
      stable = false;
      return ns + 1;
    }
    if (ns == kMaxSeg) {                           // table full (huge |inc|): finish one by one
      idx_out[i++] = (unsigned short)(unsigned)phase;
      phase = __fadd_rn(phase, inc);
      stable = false;
      continue;
    }
    const float nxt = __fadd_rn(phase, inc);
    const unsigned pb = __float_as_uint(phase), nb = __float_as_uint(nxt);
    const bool same = phase > 0.f && nxt > 0.f && ((pb ^ nb) >> 23) == 0u;
    int len = 1, dk = 0;
    float phase_next = nxt;
    bool stable_next = same;
    if (stable && same) {
      dk = (int)nb - (int)pb;
      const unsigned lo = pb & 0xFF800000u, hi = lo | 0x007FFFFFu;
      unsigned avail;                              // further steps that stay inside the binade
      if (dk < 0) avail = (pb - lo) / (unsigned)(-dk);
      else if (dk > 0) avail = (hi - pb) / (unsigned)dk;
      else avail = (unsigned)n;
      len = (avail + 1u < (unsigned)(n - i)) ? (int)(avail + 1u) : (n - i);
      const unsigned lb = pb + (unsigned)((len - 1) * dk);
      const float last = __uint_as_float(lb);
      phase_next = __fadd_rn(last, inc);
      stable_next = phase_next > 0.f && ((lb ^ __float_as_uint(phase_next)) >> 23) == 0u;
    }
    segs[ns].start = i; segs[ns].bits0 = pb; segs[ns].dk = dk; segs[ns].len = len;
    ns++;
    i += len;
    phase = phase_next;
    stable = stable_next;
  }
  return ns;
}

struct TrackShared {
  float avg[kAvgLen];
  float2 lead[256];                    // window samples 0..126 at [128..255), zeros elsewhere
  float2 trail[256];                   // window samples 9473..9599 at [1..128), zeros elsewhere
  float edge[256];                     // power at the truncated lags
  float2 rot[kRotLen];                 // CFO-corrected samples kRotBase..959 of the emitted half-frame
  unsigned short ph_idx[960];          // phasor table index per sample
  PhaseSeg seg[kMaxSeg];
  int nseg;
  float red_v[8]; int red_i[8]; float red_l[8]; float red_r[8];
  float cp_part[12];
  float2 y01[2];
  // broadcast slots
  int p, lb, ub; float peak, lmax, rmax;
  int do_emit, do_track, zero_avg, frame_start, rec_slot, sss_slot, cp_len, chain;
  long long emit_abs;
  ChainState st;
};

// canonical folded correlation power at one lag of a zero-padded window; buf index of x[k]
// is `pos`, buf[pos-127 .. pos] readable.  G = coefficient group (0: root 25, 1: roots 29/34) is a
// template parameter and the tap loop is fully unrolled so that every coefficient is a
// constant-bank operand of its FFMA2: with a run-time group the loop issued two indexed constant
// loads per tap and took 14k cycles per window (measured with clock64 instrumentation).
template <int G>
__device__ __forceinline__ float edge_power(const float2 *buf, int pos, int n_id_2) {
  float2 ab = make_float2(0.f, 0.f), dc = make_float2(0.f, 0.f);
  {
    const float2 s = buf[pos];
    ab = ffma2(c_pss_coef[G][0][0], s, ab);
    dc = ffma2(c_pss_coef[G][0][1], s, dc);
  }
#pragma unroll
  for (int m = 1; m <= 63; ++m) {
    const float2 s = fadd2(buf[pos - m], buf[pos - 128 + m]);
    ab = ffma2(c_pss_coef[G][m][0], s, ab);
    dc = ffma2(c_pss_coef[G][m][1], s, dc);
  }
  {
    const float2 s = buf[pos - 64];
    ab = ffma2(c_pss_coef[G][64][0], s, ab);
    dc = ffma2(c_pss_coef[G][64][1], s, dc);
  }
  float re, im;
  if (n_id_2 == 2) { re = __fadd_rn(ab.x, ab.y); im = __fadd_rn(dc.y, -dc.x); }
  else             { re = __fadd_rn(ab.x, -ab.y); im = __fadd_rn(dc.y, dc.x); }
  return __fmaf_rn(re, re, __fmul_rn(im, im));
}

// Queue order for pss_track_kernel: a tracking chain (CFO, CP metric and SSS extraction on every
// window) runs ~1.5x as long as a searching one, and 1536 chains over 592 resident CTAs are 2.6
// rounds, so the long jobs go first and the short ones fill the tail.  Tracking chains are placed
// from the front, the others from the back; the order inside a class does not matter (chains are
// independent and write to slots fixed by their own index).  counters: [0] front, [1] back.
__global__ void __launch_bounds__(256) chain_order_kernel(const ChainState *__restrict__ state, int n_chains,
                                                          int *__restrict__ order, int *__restrict__ counters) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_chains; i += gridDim.x * blockDim.x) {
    if (state[i].tracking) order[atomicAdd(&counters[0], 1)] = i;
    else order[n_chains - 1 - atomicAdd(&counters[1], 1)] = i;
  }
}

__global__ void __launch_bounds__(kTrackThreads, 4) pss_track_kernel(TrackParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TrackShared &S = *reinterpret_cast<TrackShared *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Persistent CTAs pull chains from a queue: a tracking chain costs ~1.5x a searching one, and
  // 1536 chains over 444 resident CTAs were 3.46 waves, i.e. four rounds of the slowest chain.
  for (;;) {
  __syncthreads();                                                  // previous chain fully written back
  if (tid == 0) {
    const int q = atomicAdd(P.chain_counter, 1);
    S.chain = q < P.n_chains ? P.chain_order[q] : P.n_chains;
  }
  __syncthreads();
  const int chain = S.chain;
  if (chain >= P.n_chains) break;
  const int stream = chain / 3, n_id_2 = chain % 3;
  if (!((P.root_mask >> n_id_2) & 1)) { if (tid == 0) P.rec_count[chain] = 0; continue; }

  const float2 *yr = P.y_ring + (size_t)stream * P.cap;
  const float *pr = P.p_ring + ((size_t)stream * 3 + n_id_2) * P.cap;
  float *avg_g = P.avg + (size_t)chain * kAvgLen;
  const float thr = P.thr[chain];

  if (tid == 0) S.st = P.state[chain];
  {
    // 39 kB of EMA state per chain: all of a thread's loads in flight at once (one round trip), 16 bytes each
    constexpr int kN4 = (kAvgLen / 4 + kTrackThreads - 1) / kTrackThreads;   // 10
    static_assert(kAvgLen % 4 == 0, "EMA state moves as float4");
    const float4 *ag4 = reinterpret_cast<const float4 *>(avg_g);
    float4 v[kN4];
#pragma unroll
    for (int j = 0; j < kN4; ++j) { const int g = tid + j * kTrackThreads; v[j] = g < kAvgLen / 4 ? ag4[g] : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
    for (int j = 0; j < kN4; ++j) { const int g = tid + j * kTrackThreads; if (g < kAvgLen / 4) reinterpret_cast<float4 *>(S.avg)[g] = v[j]; }
  }
  S.lead[tid] = make_float2(0.f, 0.f);
  S.trail[tid] = make_float2(0.f, 0.f);
  int n_rec = 0;
  __syncthreads();

  for (;;) {
    const long long R = S.st.next_win;
    if (R + kLookahead > P.n_total) break;
    const bool search = !S.st.tracking || S.st.timer == 0;          // lib/pss_impl.cc:163
    __syncthreads();                                                // everyone has read S.st
    if (search) {
      // ---- srslte_pss_find_pss ------------------------------------------------
      // thread t owns the lags 4 (t + 256 j) .. + 3, j = 0..9: four consecutive lags move as one 16-byte load of the power
      // ring (when the window start is a multiple of four samples: always for a chain that has not triggered yet) and one
      // LDS.128 / STS.128 of the moving average.  They are requested in two batches of five (register budget for four
      // CTAs per SM); the first batch is in flight during the edge computation, and both mostly hit L2 thanks to the
      // prefetch issued at the end of the previous window.  Lags below 127 and from 9600 on come from the truncated
      // window (S.edge) instead: groups 0..31 (warp 0, j = 0) and 2400..2431 (warp 3, j = 9).
      constexpr int kGrp = (kNLag + 3) / 4;                                     // 2432 groups of four lags
      constexpr int kPerThread = (kGrp + kTrackThreads - 1) / kTrackThreads;    // 10
      constexpr int kBatch = kPerThread / 2;                                    // 5
      static_assert(kPerThread == 2 * kBatch, "two equal batches");
      const bool aligned4 = ((unsigned)R & 3u) == 0u;
      float4 a[kBatch];
      auto load_batch = [&](const int h) {
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          const int k0 = 4 * (tid + (h * kBatch + j) * kTrackThreads);
          float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k0 >= 128 && k0 < kHalf) {
            if (aligned4) {
              pv = __ldg(reinterpret_cast<const float4 *>(&pr[(unsigned)((R + k0) & P.cap_mask)]));
            } else {
              pv.x = __ldg(&pr[(unsigned)((R + k0) & P.cap_mask)]);
              pv.y = __ldg(&pr[(unsigned)((R + k0 + 1) & P.cap_mask)]);
              pv.z = __ldg(&pr[(unsigned)((R + k0 + 2) & P.cap_mask)]);
              pv.w = __ldg(&pr[(unsigned)((R + k0 + 3) & P.cap_mask)]);
            }
          } else if (k0 == 124) {
            pv.w = __ldg(&pr[(unsigned)((R + 127) & P.cap_mask)]);               // lag 127: the first whole one
          }
          a[j] = pv;
        }
      };
      load_batch(0);
      if (tid < 127) {
        S.lead[128 + tid] = yr[(unsigned)((R + tid) & P.cap_mask)];
        S.trail[1 + tid] = yr[(unsigned)((R + 9473 + tid) & P.cap_mask)];
      }
      __syncthreads();
      if (tid < kNEdge) {
        // lag k<127: x[k] at lead[128+k];  lag k>=9600: x[k] at trail[k-9472]
        const float2 *eb = (tid < 127) ? S.lead : S.trail;
        const int ep = (tid < 127) ? 128 + tid : 128 + (tid - 127);
        S.edge[tid] = (n_id_2 == 0) ? edge_power<0>(eb, ep, n_id_2) : edge_power<1>(eb, ep, n_id_2);
      }
      __syncthreads();
      float best = -3.402823466e+38f; int bi = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1) load_batch(1);
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          const int k0 = 4 * (tid + (h * kBatch + j) * kTrackThreads);
          if (k0 < kNLag) {
            float4 av = a[j];
            if (k0 < 128) {                                          // leading edge: lags k0 .. k0 + 3 < 127 (127 itself is whole)
              av.x = S.edge[k0]; av.y = S.edge[k0 + 1]; av.z = S.edge[k0 + 2];
              if (k0 + 3 < 127) av.w = S.edge[k0 + 3];
            } else if (k0 >= kHalf) {                                // trailing edge: lags 9600 .. 9725
              av.x = S.edge[127 + (k0 - kHalf)]; av.y = S.edge[128 + (k0 - kHalf)];
              if (k0 + 3 < kNLag) { av.z = S.edge[129 + (k0 - kHalf)]; av.w = S.edge[130 + (k0 - kHalf)]; }
            }
            float4 *slot = reinterpret_cast<float4 *>(&S.avg[k0]);
            const float4 o = *slot;
            float4 v;                                                // EMA, alpha 0.2
            v.x = __fadd_rn(__fmul_rn(av.x, 0.2f), __fmul_rn(o.x, 0.8f));
            v.y = __fadd_rn(__fmul_rn(av.y, 0.2f), __fmul_rn(o.y, 0.8f));
            v.z = __fadd_rn(__fmul_rn(av.z, 0.2f), __fmul_rn(o.z, 0.8f));
            v.w = __fadd_rn(__fmul_rn(av.w, 0.2f), __fmul_rn(o.w, 0.8f));
            if (k0 + 3 < kNLag) {
              *slot = v;
              if (v.x > best) { best = v.x; bi = k0; }
              if (v.y > best) { best = v.y; bi = k0 + 1; }
              if (v.z > best) { best = v.z; bi = k0 + 2; }
              if (v.w > best) { best = v.w; bi = k0 + 3; }
            } else {                                                 // the last group: lags 9724, 9725 only
              S.avg[k0] = v.x; S.avg[k0 + 1] = v.y;
              if (v.x > best) { best = v.x; bi = k0; }
              if (v.y > best) { best = v.y; bi = k0 + 1; }
            }
          }
        }
      }
      // first-index argmax over the block
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_down_sync(0xffffffffu, best, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) { S.red_v[warp] = best; S.red_i[warp] = bi; }
      __syncthreads();
      if (warp == 0) {
        float bv = S.red_v[0]; int bidx = S.red_i[0];
        for (int w = 1; w < kTrackThreads / 32; ++w)
          if (S.red_v[w] > bv || (S.red_v[w] == bv && S.red_i[w] < bidx)) { bv = S.red_v[w]; bidx = S.red_i[w]; }
        const int p = bidx;
        // main-lobe walk (srslte_pss_find_pss: ub climbs from p + 1 while avg[ub + 1] <= avg[ub], lb descends from p - 1
        // while avg[lb - 1] <= avg[lb]), 32 lags per step by the lanes of one warp: the first lane whose lag ends the walk
        // gives the same bound as the one-by-one loop, and a plateau (silence, or a noiseless capture with empty symbols,
        // where the loop runs over thousands of equal values) costs 1/32 of the steps
        const int conv_output_len = kNLag + 1;
        int ub = p + 1;
        for (;;) {
          const int i = ub + lane;
          bool stop = true;                                            // beyond the array: the loop has ended before
          if (i + 1 < kAvgLen) stop = !(S.avg[i + 1] <= S.avg[i] && i < conv_output_len);
          const unsigned m = __ballot_sync(0xffffffffu, stop);
          if (m) { ub += __ffs(m) - 1; break; }
          ub += 32;
        }
        int lb = 0;
        if (p > 2) {
          lb = p - 1;
          for (;;) {
            const int i = lb - lane;
            bool stop = true;
            if (i >= 1) stop = !(S.avg[i - 1] <= S.avg[i] && i > 1);
            const unsigned m = __ballot_sync(0xffffffffu, stop);
            if (m) { lb -= __ffs(m) - 1; break; }
            lb -= 32;
          }
        }
        if (lane == 0) { S.p = p; S.peak = S.avg[p]; S.lb = lb; S.ub = ub; }
      }
      __syncthreads();
      {
        // largest side lobe: the maximum over the lags outside [lb, ub).  Thread 0 below wants the maxima to the left
        // and to the right separately only to take the larger (and to replace an empty side by its fall-back value), so
        // one maximum over everything outside the main lobe serves as both
        const int lb = S.lb, ub = S.ub;
        float om = -3.402823466e+38f;
#pragma unroll
        for (int j = 0; j < (kNLag + 3) / 4 / kTrackThreads + 1; ++j) {
          const int k0 = 4 * (tid + j * kTrackThreads);
          if (k0 < kNLag) {
            const float4 v = *reinterpret_cast<const float4 *>(&S.avg[k0]);
            if ((k0 + 3 < lb || k0 >= ub) && k0 + 3 < kNLag) {
              om = fmaxf(om, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
            } else {
              if (k0 < lb || k0 >= ub) om = fmaxf(om, v.x);
              if (k0 + 1 < kNLag && (k0 + 1 < lb || k0 + 1 >= ub)) om = fmaxf(om, v.y);
              if (k0 + 2 < kNLag && (k0 + 2 < lb || k0 + 2 >= ub)) om = fmaxf(om, v.z);
              if (k0 + 3 < kNLag && (k0 + 3 < lb || k0 + 3 >= ub)) om = fmaxf(om, v.w);
            }
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) om = fmaxf(om, __shfl_down_sync(0xffffffffu, om, off));
        const float lm = om, rm = om;
        if (lane == 0) { S.red_l[warp] = lm; S.red_r[warp] = rm; }
      }
      __syncthreads();
    }
    if (tid == 0) {
      ChainState &st = S.st;
      unsigned flags = 0;
      if (search) {
        float lm = S.red_l[0], rm = S.red_r[0];
        for (int w = 1; w < kTrackThreads / 32; ++w) { lm = fmaxf(lm, S.red_l[w]); rm = fmaxf(rm, S.red_r[w]); }
        // vec_max_fi over an empty range returns index 0
        const float left = (S.lb > 0) ? lm : S.avg[0];
        const float right = (S.ub < kNLag) ? rm : S.avg[S.ub];
        const float side = left > right ? left : right;
        st.timer = P.track_every;
        st.peak_pos = S.p;
        st.peak_value = S.peak;
        st.psr = __fdiv_rn(S.peak, side);
        st.psr_data[st.psr_i++ % kMavg] = st.psr;
        flags |= LTB_F_SEARCHED;
      } else {
        st.timer--;
      }
      const bool over = st.psr > thr;                               // :174
      int zero_avg = 0;
      if (over) {                                                   // incr_score :111-127
        flags |= LTB_F_OVER;
        if (!(st.tracking && st.score == P.track_after)) {
          st.score++;
          if (!st.tracking && st.score == P.track_after) { st.tracking = 1; zero_avg = 1; }
        }
      } else if (st.score != 0) {                                   // reset_score :129-152
        st.score = 0; st.timer = 0; st.tracking = 0;
        zero_avg = 1;
        for (int i = 0; i < kMavg; ++i) { st.psr_data[i] = 0.f; st.cfo_data[i] = 0.f; }
        st.psr_i = 0; st.cfo_i = 0; st.cfo_last_freq = 0.f;
        st.lost = 1;
      }
      if (st.psr > st.psr_max) st.psr_max = st.psr;                 // :181
      const int peak_used = st.peak_pos;
      int nconsume = kHalf, do_emit = 0, do_track = 0, frame_start = 0;
      if (over || st.lost) {                                        // :184
        frame_start = st.peak_pos - kSlot;
        st.peak_pos = kSlot;
        nconsume = frame_start + kHalf;
        do_emit = 1; flags |= LTB_F_EMIT;
        if (st.tracking) { do_track = 1; flags |= LTB_F_TRACKING; }
        else { flags |= LTB_F_TAG_LOST; st.lost = 0; }
      }
      S.do_emit = do_emit; S.do_track = do_track; S.zero_avg = zero_avg;
      S.frame_start = frame_start; S.emit_abs = R + frame_start;
      S.rec_slot = -1;
      if ((do_emit || P.record_all) && n_rec < P.w_max) {
        S.rec_slot = chain * P.w_max + n_rec;
        ltb_window_rec r;
        r.win_start = R; r.emit_start = do_emit ? R + frame_start : -1;
        r.stream = stream; r.n_id_2 = n_id_2; r.win_index = st.win_index; r.flags = flags;
        r.peak_pos = peak_used; r.score = st.score; r.psr = st.psr; r.peak_value = st.peak_value;
        r.cfo = 0.f; r.mean_cfo = 0.f; r.m0 = -1; r.m1 = -1; r.m0_val = 0.f; r.m1_val = 0.f;
        r.n_id_1 = -1; r.cell_id = -1; r.cp_norm_avg = 0.f; r.cp_ext_avg = 0.f;
        P.recs[S.rec_slot] = r;
      }
      if (do_emit && !do_track) { st.cp_norm_avg = 0.f; st.cp_ext_avg = 0.f; }   // sss: srslte_sync_reset on the tag
      st.win_index++;
      st.next_win = R + nconsume;
    }
    __syncthreads();
    if (S.rec_slot >= 0) n_rec++;
    {
      // the next window starts at S.st.next_win: pull its power lags and edge samples into L2 now
      // (prefetch only: no registers, harmless if the chain stops or skips the search), so the
      // round trip to HBM overlaps the rest of this window instead of opening the next one
      const long long Rn = S.st.next_win;
      if (Rn + kLookahead <= P.n_total && (!S.st.tracking || S.st.timer == 0)) {
        for (int i = tid; i < (kHalf - 127 + 31) / 32 + 1; i += kTrackThreads)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(&pr[(unsigned)((Rn + 127 + 32 * i) & P.cap_mask)]));
        if (tid < 18) {
          const long long o = (tid < 9) ? Rn + 16 * tid : Rn + 9473 + 16 * (tid - 9);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(&yr[(unsigned)(o & P.cap_mask)]));
        }
      }
    }
    if (S.zero_avg) {                                               // srslte_pss_reset
      for (int g = tid; g < kAvgLen / 4; g += kTrackThreads) reinterpret_cast<float4 *>(S.avg)[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (S.do_emit) {
      const long long E = S.emit_abs;
      float2 *hf = nullptr;
      if (P.hf_out != nullptr && S.rec_slot >= 0) hf = P.hf_out + (size_t)S.rec_slot * kHalf;
      if (S.do_track) {
        // ---- srslte_pss_cfo_compute on out[832..960) (lib/pss_impl.cc:199) --------------
        for (int i = tid; i < kRotLen; i += kTrackThreads) S.rot[i] = yr[(unsigned)((E + kRotBase + i) & P.cap_mask)];   // raw
        __syncthreads();
        if (tid < 2) {
          float yr_ = 0.f, yi_ = 0.f;
          const int nb = tid * 64;
          for (int n = nb; n < nb + 64; ++n) {
            const float2 h = c_pss_taps[n_id_2][n];
            const float2 r = S.rot[832 - kRotBase + n];                          // out[832 + n]
            yr_ = __fmaf_rn(h.x, r.x, yr_); yr_ = __fmaf_rn(-h.y, r.y, yr_);
            yi_ = __fmaf_rn(h.x, r.y, yi_); yi_ = __fmaf_rn(h.y, r.x, yi_);
          }
          S.y01[tid] = make_float2(yr_, yi_);
        }
        __syncthreads();
        if (tid == 0) {
          ChainState &st = S.st;
          const float2 y0 = S.y01[0], y1 = S.y01[1];
          const float pre = __fmaf_rn(y0.x, y1.x, __fmul_rn(y0.y, y1.y));       // conj(y0)*y1
          const float pim = __fmaf_rn(y0.x, y1.y, -__fmul_rn(y0.y, y1.x));
          const float cfo = (float)__ddiv_rn(canon_atan2((double)pim, (double)pre), 3.14159265358979323846);
          st.cfo_data[st.cfo_i++ % kMavg] = cfo;                                 // :200
          unsigned npts = st.cfo_i > (unsigned)kMavg ? (unsigned)kMavg : st.cfo_i;
          double acc = 0.0;
          for (unsigned i = 0; i < npts; ++i) acc = __dadd_rn(acc, (double)st.cfo_data[i]);
          const float mcfo = (float)__ddiv_rn(acc, (double)npts);
          const float freq = __fdiv_rn(-mcfo, 128.0f);                           // :204
          if (fabsf(__fadd_rn(st.cfo_last_freq, -freq)) > 0.0f) { st.cfo_last_freq = freq; st.cfo_table_freq = freq; }
          if (S.rec_slot >= 0) { P.recs[S.rec_slot].cfo = cfo; P.recs[S.rec_slot].mean_cfo = mcfo; }
        }
        __syncthreads();
        // ---- srslte_cfo_correct: phasor index scan (sequential float phase), then rotate -----
        const float phase_inc = __fmul_rn(S.st.cfo_table_freq, 4096.0f);
        const int n_blocks = hf ? 10 : 1;
        float phase = 0.f;       // carried by thread 0 across blocks
        bool stable = false;
        for (int b = 0; b < n_blocks; ++b) {
          if (tid == 0) S.nseg = plan_phase_scan(phase, stable, phase_inc, 960, S.seg, S.ph_idx);
          __syncthreads();
          for (int sg = 0; sg < S.nseg; ++sg) {
            const PhaseSeg q = S.seg[sg];
            for (int j = tid; j < q.len; j += kTrackThreads)
              S.ph_idx[q.start + j] = q.dk == kSegPingPong ? (unsigned short)((j & 1) ? 4096 : 0)
                                                           : (unsigned short)(unsigned)__uint_as_float(q.bits0 + (unsigned)(j * q.dk));
          }
          __syncthreads();
          if (b == 0) {
            for (int i = tid; i < kRotLen; i += kTrackThreads) S.rot[i] = cmul_canon(P.cexp[S.ph_idx[kRotBase + i]], S.rot[i]);
          }
          if (hf) {
            for (int i = tid; i < 960; i += kTrackThreads) {
              const float2 x = yr[(unsigned)((E + b * 960 + i) & P.cap_mask)];
              hf[b * 960 + i] = cmul_canon(P.cexp[S.ph_idx[i]], x);
            }
          }
          __syncthreads();
        }
        // ---- srslte_sync_detect_cp(in, 960) (lib/sss_impl.cc:104) ----------------------------
        if (tid < 12) {
          const int h = tid / 6, sy = (tid % 6) / 2, kind = tid & 1;
          const int cp = h ? 32 : 9;
          const int j0 = 960 - 3 * (128 + cp) + sy * (128 + cp) - kRotBase;   // index into rot
          float acc = 0.f;
          if (kind == 0) {
            for (int i = 0; i < cp; ++i) {
              acc = __fmaf_rn(S.rot[j0 + 128 + i].x, S.rot[j0 + i].x, acc);
              acc = __fmaf_rn(S.rot[j0 + 128 + i].y, S.rot[j0 + i].y, acc);
            }
          } else {
            for (int i = 0; i < cp; ++i) {
              acc = __fmaf_rn(S.rot[j0 + i].x, S.rot[j0 + i].x, acc);
              acc = __fmaf_rn(S.rot[j0 + i].y, S.rot[j0 + i].y, acc);
            }
          }
          S.cp_part[tid] = acc;
        }
        __syncthreads();
        if (tid == 0) {
          ChainState &st = S.st;
          float Rv[2], Mv[2];
          for (int h = 0; h < 2; ++h) {
            const float cpf = h ? 32.0f : 9.0f;
            float Rs = 0.f, Cs = 0.f;
            for (int sy = 0; sy < 3; ++sy) {
              Rs = __fadd_rn(Rs, S.cp_part[h * 6 + sy * 2]);
              Cs = __fadd_rn(Cs, __fmul_rn(cpf, __fdiv_rn(S.cp_part[h * 6 + sy * 2 + 1], cpf)));
            }
            Rv[h] = Rs;
            Mv[h] = (Cs > 0.f) ? __fdiv_rn(Rs, Cs) : 0.f;
          }
          const float mn = __fdiv_rn(Mv[0], 3.0f), me = __fdiv_rn(Mv[1], 3.0f);
          const double one_m = __dadd_rn(1.0, -0.1);
          st.cp_norm_avg = (float)__dadd_rn(__dmul_rn(0.1, (double)mn), __dmul_rn(one_m, (double)st.cp_norm_avg));
          st.cp_ext_avg = (float)__dadd_rn(__dmul_rn(0.1, (double)me), __dmul_rn(one_m, (double)st.cp_ext_avg));
          int cp_norm;
          if (st.cp_norm_avg > st.cp_ext_avg) cp_norm = 1;
          else if (st.cp_norm_avg < st.cp_ext_avg) cp_norm = 0;
          else cp_norm = Rv[0] > Rv[1] ? 1 : 0;
          S.cp_len = cp_norm ? 9 : 32;
          S.sss_slot = -1;
          if (S.rec_slot >= 0) {
            ltb_window_rec &r = P.recs[S.rec_slot];
            r.flags |= LTB_F_SSS | (cp_norm ? LTB_F_CP_NORM : 0u);
            r.cp_norm_avg = st.cp_norm_avg; r.cp_ext_avg = st.cp_ext_avg;
            const int slot = atomicAdd(P.sss_count, 1);
            if (slot < P.sss_cap) { S.sss_slot = slot; P.sss_rec[slot] = S.rec_slot; }
          }
        }
        __syncthreads();
        if (S.sss_slot >= 0 && tid < 128) {
          const int sss_idx = sss_symbol_start(S.cp_len, P.tdd);       // lib/sss_impl.cc:110
          P.sss_sym[(size_t)S.sss_slot * 128 + tid] = S.rot[sss_idx - kRotBase + tid];
        }
      } else if (hf) {
        for (int i = tid; i < kHalf; i += kTrackThreads) hf[i] = yr[(unsigned)((E + i) & P.cap_mask)];
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) { P.state[chain] = S.st; P.rec_count[chain] = n_rec; }
  for (int g = tid; g < kAvgLen / 4; g += kTrackThreads) reinterpret_cast<float4 *>(avg_g)[g] = reinterpret_cast<const float4 *>(S.avg)[g];
  }
}

// ------------------------------------------------------------------------------------
// K4: batched SSS decode -- srslte_sss_m0m1_partial(M=1, ce=NULL) + srslte_sss_N_id_1 +
//     srslte_sync_get_cell_id (lib/sss_impl.cc:112-124) -- one warp per candidate symbol.
// ------------------------------------------------------------------------------------
constexpr int kSssWarps = 4;

__device__ __forceinline__ void fft128_warp(float2 *a, int lane, const float2 *tw) {
  // radix-2 DIT on bit-reversed input in shared memory a[128]; 64 butterflies per stage,
  // two per lane.  Same butterfly as the oracle: t = w*b (canonical), a' = a+t, b' = a-t.
#pragma unroll
  for (int half = 1; half < 128; half <<= 1) {
    const int step = 64 / half;
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      const int bf = lane + rep * 32;            // butterfly id 0..63
      const int j = bf & (half - 1);
      const int base = (bf / half) * 2 * half;
      const float2 w = tw[j * step];
      const float2 b = a[base + j + half];
      const float2 u = a[base + j];
      const float2 tw = cmul_canon(w, b);
      a[base + j] = make_float2(__fadd_rn(u.x, tw.x), __fadd_rn(u.y, tw.y));
      a[base + j + half] = make_float2(__fadd_rn(u.x, -tw.x), __fadd_rn(u.y, -tw.y));
    }
    __syncwarp();
  }
}

__device__ __forceinline__ int sss_corr_argmax(const float2 *y, int lane, float *val_out, const float *s2) {
  // lane m < 31: |sum_i y[i] * s_tilde[(i+m)%31]|^2, sequential i; first-index argmax.
  // s2 = s_tilde repeated twice in shared memory: no modulo, no lane-indexed constant loads
  float v = -3.402823466e+38f;
  if (lane < 31) {
    float ar = 0.f, ai = 0.f;
#pragma unroll
    for (int i = 0; i < 31; ++i) {
      const float sv = s2[i + lane];
      ar = __fmaf_rn(y[i].x, sv, ar);
      ai = __fmaf_rn(y[i].y, sv, ai);
    }
    v = __fmaf_rn(ar, ar, __fmul_rn(ai, ai));
  }
  int idx = lane;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_down_sync(0xffffffffu, v, off);
    const int oi = __shfl_down_sync(0xffffffffu, idx, off);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  idx = __shfl_sync(0xffffffffu, idx, 0);
  *val_out = __shfl_sync(0xffffffffu, v, 0);
  return idx;
}

__global__ void __launch_bounds__(kSssWarps * 32) sss_kernel(const float2 *__restrict__ sss_sym,
                                                              const int *__restrict__ sss_rec,
                                                              const int *__restrict__ sss_count, int sss_cap,
                                                              ltb_window_rec *__restrict__ recs) {
  __shared__ float2 sa[kSssWarps][128];
  __shared__ float2 sy[kSssWarps][2][32];
  __shared__ float2 s_tw[64];
  __shared__ float s_s2[62];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // lane-indexed reads of __constant__ tables replay once per distinct address: stage them
  if (threadIdx.x < 64) s_tw[threadIdx.x] = c_fft128_tw[threadIdx.x];
  if (threadIdx.x < 62) s_s2[threadIdx.x] = c_sss_s[threadIdx.x % 31];
  __syncthreads();
  int count = *sss_count;
  if (count > sss_cap) count = sss_cap;
  for (int slot = blockIdx.x * kSssWarps + warp; slot < count; slot += gridDim.x * kSssWarps) {
    float2 *a = sa[warp];
    for (int i = lane; i < 128; i += 32) a[__brev((unsigned)i) >> 25] = sss_sym[(size_t)slot * 128 + i];
    __syncwarp();
    fft128_warp(a, lane, s_tw);
    ltb_window_rec &r = recs[sss_rec[slot]];
    const int n_id_2 = r.n_id_2;
    if (lane < 31) {
      const int e = 2 * lane, o = 2 * lane + 1;
      const int be = (e < 31) ? 97 + e : e - 30;
      const int bo = (o < 31) ? 97 + o : o - 30;
      const float c0 = c_sss_c0[n_id_2][lane], c1 = c_sss_c1[n_id_2][lane];
      sy[warp][0][lane] = make_float2(__fmul_rn(a[be].x, c0), __fmul_rn(a[be].y, c0));
      sy[warp][1][lane] = make_float2(__fmul_rn(a[bo].x, c1), __fmul_rn(a[bo].y, c1));
    }
    __syncwarp();
    float m0v, m1v;
    const int m0 = sss_corr_argmax(sy[warp][0], lane, &m0v, s_s2);
    if (lane < 31) {
      const float z = c_sss_z[(lane + (m0 % 8)) % 31];
      sy[warp][1][lane] = make_float2(__fmul_rn(sy[warp][1][lane].x, z), __fmul_rn(sy[warp][1][lane].y, z));
    }
    __syncwarp();
    const int m1 = sss_corr_argmax(sy[warp][1], lane, &m1v, s_s2);
    if (lane == 0) {
      int nid = -1;
      const unsigned um0 = (unsigned)m0, um1 = (unsigned)m1;
      if (um1 > um0) { if (um0 < 30u && um1 - 1u < 30u) nid = c_sss_nid1[um0 * 30 + (um1 - 1)]; }
      else           { if (um1 < 30u && um0 - 1u < 30u) nid = c_sss_nid1[um1 * 30 + (um0 - 1)]; }
      r.m0 = m0; r.m1 = m1; r.m0_val = m0v; r.m1_val = m1v; r.n_id_1 = nid;
      if (nid >= 0) { r.cell_id = 3 * nid + n_id_2; r.flags |= LTB_F_CELL; }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------
// standalone sss block: CP detection + symbol extraction for n consecutive half-frames of one
// chain (sequential EMA), feeding sss_kernel.  One CTA.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) sss_block_front_kernel(const float2 *__restrict__ hf, const int *__restrict__ tag_lost,
                                                              int n_hf, int n_id_2, int tdd, float *cp_state /*2*/,
                                                              ltb_window_rec *recs, float2 *sss_sym, int *sss_rec,
                                                              int *sss_count) {
  __shared__ float part[12];
  __shared__ int s_cp_len, s_slot;
  const int tid = threadIdx.x;
  float cpn = cp_state[0], cpe = cp_state[1];
  for (int f = 0; f < n_hf; ++f) {
    const float2 *in = hf + (size_t)f * kHalf;
    if (tag_lost[f]) { cpn = 0.f; cpe = 0.f; __syncthreads(); continue; }
    if (tid < 12) {
      const int h = tid / 6, sy = (tid % 6) / 2, kind = tid & 1;
      const int cp = h ? 32 : 9;
      const int j0 = 960 - 3 * (128 + cp) + sy * (128 + cp);
      float acc = 0.f;
      if (kind == 0) {
        for (int i = 0; i < cp; ++i) {
          acc = __fmaf_rn(in[j0 + 128 + i].x, in[j0 + i].x, acc);
          acc = __fmaf_rn(in[j0 + 128 + i].y, in[j0 + i].y, acc);
        }
      } else {
        for (int i = 0; i < cp; ++i) {
          acc = __fmaf_rn(in[j0 + i].x, in[j0 + i].x, acc);
          acc = __fmaf_rn(in[j0 + i].y, in[j0 + i].y, acc);
        }
      }
      part[tid] = acc;
    }
    __syncthreads();
    if (tid == 0) {
      float Rv[2], Mv[2];
      for (int h = 0; h < 2; ++h) {
        const float cpf = h ? 32.0f : 9.0f;
        float Rs = 0.f, Cs = 0.f;
        for (int sy = 0; sy < 3; ++sy) {
          Rs = __fadd_rn(Rs, part[h * 6 + sy * 2]);
          Cs = __fadd_rn(Cs, __fmul_rn(cpf, __fdiv_rn(part[h * 6 + sy * 2 + 1], cpf)));
        }
        Rv[h] = Rs;
        Mv[h] = (Cs > 0.f) ? __fdiv_rn(Rs, Cs) : 0.f;
      }
      const float mn = __fdiv_rn(Mv[0], 3.0f), me = __fdiv_rn(Mv[1], 3.0f);
      const double one_m = __dadd_rn(1.0, -0.1);
      cpn = (float)__dadd_rn(__dmul_rn(0.1, (double)mn), __dmul_rn(one_m, (double)cpn));
      cpe = (float)__dadd_rn(__dmul_rn(0.1, (double)me), __dmul_rn(one_m, (double)cpe));
      int cp_norm;
      if (cpn > cpe) cp_norm = 1; else if (cpn < cpe) cp_norm = 0; else cp_norm = Rv[0] > Rv[1] ? 1 : 0;
      s_cp_len = cp_norm ? 9 : 32;
      ltb_window_rec &r = recs[f];
      r.n_id_2 = n_id_2;
      r.flags |= LTB_F_SSS | (cp_norm ? LTB_F_CP_NORM : 0u);
      r.cp_norm_avg = cpn; r.cp_ext_avg = cpe;
      r.m0 = r.m1 = -1; r.n_id_1 = -1; r.cell_id = -1;
      s_slot = atomicAdd(sss_count, 1);
      sss_rec[s_slot] = f;
    }
    __syncthreads();
    // broadcast the EMA state held by thread 0 to the others through shared memory
    {
      __shared__ float s_cpn, s_cpe;
      if (tid == 0) { s_cpn = cpn; s_cpe = cpe; }
      __syncthreads();
      cpn = s_cpn; cpe = s_cpe;
    }
    const int sss_idx = sss_symbol_start(s_cp_len, tdd);
    sss_sym[(size_t)s_slot * 128 + tid] = in[sss_idx + tid];
    __syncthreads();
  }
  if (tid == 0) { cp_state[0] = cpn; cp_state[1] = cpe; }
}

}  // namespace ltb
