// C ABI of libltetrigger_b200.so (declared in include/ltetrigger_b200.h): context
// management, launch sequencing on one CUDA stream, record hand-back.  No CPU fallback:
// every compute entry point fails with LTB_ERROR when no CUDA device is usable.
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ltetrigger_b200.h"
#include "ltb_kernels.cuh"
#include "ltb_tables.h"
#include "ltb_tc_frontend.cuh"

using namespace ltb;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define LTB_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      return fail(LTB_ERROR, std::string(#call) + ": " + cudaGetErrorString(e__));           \
    }                                                                                       \
  } while (0)

int next_pow2(long long v) {
  long long p = 1;
  while (p < v) p <<= 1;
  return (int)p;
}

// ---- constant tables, uploaded once per device ------------------------------------------
std::mutex g_const_mu;
bool g_const_done[64] = {false};
int g_sm_count[64] = {0};

int ensure_constants(int device) {
  std::lock_guard<std::mutex> lk(g_const_mu);
  if (device < 0 || device >= 64) return fail(LTB_ERROR_INVALID_INPUTS, "device ordinal out of range");
  if (g_const_done[device]) return LTB_SUCCESS;
  PssTaps taps[3];
  for (int r = 0; r < 3; ++r) make_pss_taps(r, taps[r]);
  float2 coef[2][65][2];
  for (int g = 0; g < 2; ++g)
    for (int m = 0; m <= 64; ++m) {
      coef[g][m][0] = make_float2(taps[g].re[m], taps[g].im[m]);
      coef[g][m][1] = make_float2(taps[g].im[m], taps[g].re[m]);
    }
  LTB_CUDA(cudaMemcpyToSymbol(c_pss_coef, coef, sizeof coef));
  float2 full[3][128];
  for (int r = 0; r < 3; ++r)
    for (int n = 0; n < 128; ++n) full[r][n] = make_float2(taps[r].re[n], taps[r].im[n]);
  LTB_CUDA(cudaMemcpyToSymbol(c_pss_taps, full, sizeof full));
  static float dt[sizeof(c_decim_taps) / sizeof(float)];
  std::memset(dt, 0, sizeof dt);
  const int stream_rates[7] = {4, 8, 12, 13, 14, 15, 16};
  for (int d : stream_rates) {
    std::vector<float> v = make_decim_taps(d);
    if ((int)v.size() != decim_ntaps(d)) return fail(LTB_ERROR, "unexpected decimator tap count");
    std::memcpy(dt + decim_tap_offset(d), v.data(), v.size() * sizeof(float));
  }
  LTB_CUDA(cudaMemcpyToSymbol(c_decim_taps, dt, sizeof dt));
  {
    static float2 pairs[sizeof(c_decim_pairs) / sizeof(float2)];
    std::memset(pairs, 0, sizeof pairs);
    for (int d = 2; d <= 15; ++d) {
      if (!decim_is_tiled(d)) continue;
      std::vector<float> vq = make_decim_branch_taps(d);
      if (vq.empty()) return fail(LTB_ERROR, "unexpected decimator tap count");
      for (int j = 0; j < d * kDecQ; ++j) pairs[decim_pair_offset(d) + j] = make_float2(vq[j], vq[j]);
    }
    LTB_CUDA(cudaMemcpyToSymbol(c_decim_pairs, pairs, sizeof pairs));
  }
  float twr[64], twi[64];
  make_fft128_twiddles(twr, twi);
  float2 tw[64];
  for (int k = 0; k < 64; ++k) tw[k] = make_float2(twr[k], twi[k]);
  LTB_CUDA(cudaMemcpyToSymbol(c_fft128_tw, tw, sizeof tw));
  float c0[3][32], c1[3][32], sv[32], zv[32];
  short nid[900];
  std::memset(c0, 0, sizeof c0); std::memset(c1, 0, sizeof c1);
  std::memset(sv, 0, sizeof sv); std::memset(zv, 0, sizeof zv);
  for (int r = 0; r < 3; ++r) {
    SssTables st;
    make_sss_tables(r, st);
    for (int i = 0; i < 31; ++i) { c0[r][i] = (float)st.c0[i]; c1[r][i] = (float)st.c1[i]; }
    if (r == 0) {
      for (int i = 0; i < 31; ++i) { sv[i] = (float)st.s_tilde[i]; zv[i] = (float)st.z_tilde[i]; }
      for (int i = 0; i < 900; ++i) nid[i] = (short)st.n_id_1[i];
    }
  }
  LTB_CUDA(cudaMemcpyToSymbol(c_sss_c0, c0, sizeof c0));
  LTB_CUDA(cudaMemcpyToSymbol(c_sss_c1, c1, sizeof c1));
  LTB_CUDA(cudaMemcpyToSymbol(c_sss_s, sv, sizeof sv));
  LTB_CUDA(cudaMemcpyToSymbol(c_sss_z, zv, sizeof zv));
  LTB_CUDA(cudaMemcpyToSymbol(c_sss_nid1, nid, sizeof nid));
  LTB_CUDA(cudaFuncSetAttribute(pss_track_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(TrackShared)));
  LTB_CUDA(cudaDeviceGetAttribute(&g_sm_count[device], cudaDevAttrMultiProcessorCount, device));
#define LTB_SMEM_ATTR(kernel, bytes) LTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)))
#define LTB_SMEM_ATTR_FMT(FMT)                                                                  \
  LTB_SMEM_ATTR(decimate_stream_kernel<FMT>, decim_stream_smem_bytes<FMT>());                   \
  LTB_SMEM_ATTR(decimate_any_kernel<FMT>, decim_any_smem_bytes(kMaxDecim));                     \
  LTB_SMEM_ATTR((decimate_stream12_kernel<FMT, 12>), (decim_stream12_smem_bytes<FMT, 12>()));   \
  LTB_SMEM_ATTR((decimate_stream12_kernel<FMT, 13>), (decim_stream12_smem_bytes<FMT, 13>()));   \
  LTB_SMEM_ATTR((decimate_stream12_kernel<FMT, 14>), (decim_stream12_smem_bytes<FMT, 14>()));   \
  LTB_SMEM_ATTR((decimate_stream12_kernel<FMT, 15>), (decim_stream12_smem_bytes<FMT, 15>()));   \
  LTB_SMEM_ATTR((decimate_stream2_kernel<FMT, 8>), (decim_stream2_smem_bytes<FMT, 8>()));       \
  LTB_SMEM_ATTR((decimate_stream2_kernel<FMT, 4>), (decim_stream2_smem_bytes<FMT, 4>()));       \
  LTB_SMEM_ATTR((decimate_kernel<FMT, 4>), decim_smem_bytes(4)); \
  LTB_SMEM_ATTR((decimate_kernel<FMT, 5>), decim_smem_bytes(5)); \
  LTB_SMEM_ATTR((decimate_kernel<FMT, 6>), decim_smem_bytes(6)); \
  LTB_SMEM_ATTR((decimate_kernel<FMT, 7>), decim_smem_bytes(7)); \
  LTB_SMEM_ATTR((decimate_kernel<FMT, 8>), decim_smem_bytes(8)); \
  LTB_SMEM_ATTR((decimate_kernel<FMT, 9>), decim_smem_bytes(9)); \
  LTB_SMEM_ATTR((decimate_kernel<FMT, 10>), decim_smem_bytes(10)); \
  LTB_SMEM_ATTR((decimate_kernel<FMT, 11>), decim_smem_bytes(11)); \
  LTB_SMEM_ATTR((decimate_kernel<FMT, 12>), decim_smem_bytes(12)); \
  LTB_SMEM_ATTR((decimate_kernel<FMT, 13>), decim_smem_bytes(13)); \
  LTB_SMEM_ATTR((decimate_kernel<FMT, 14>), decim_smem_bytes(14)); \
  LTB_SMEM_ATTR((decimate_kernel<FMT, 15>), decim_smem_bytes(15));
  LTB_SMEM_ATTR_FMT(LTB_FMT_FC32)
  LTB_SMEM_ATTR_FMT(LTB_FMT_SC16)
  LTB_SMEM_ATTR_FMT(LTB_FMT_SC8)
#undef LTB_SMEM_ATTR_FMT
#undef LTB_SMEM_ATTR
  g_const_done[device] = true;
  return LTB_SUCCESS;
}

// overlap-save correlator tables in the kernel's access order (one copy per device):
// tw_perm[r][l] = W_1024^(l * bitrev5(r)),  h_perm[g][r][l] = H_g[l + 32 * bitrev5(r)]
float2 *g_os_tw[64] = {nullptr};
float2 *g_os_h[64] = {nullptr};

int ensure_os_tables(int device) {
  std::lock_guard<std::mutex> lk(g_const_mu);
  if (g_os_tw[device]) return LTB_SUCCESS;
  std::vector<float> wr(1024), wi(1024), hr(1024), hi(1024);
  make_fft1024_twiddles(wr.data(), wi.data());
  std::vector<float2> tw(1024), hp(3 * 1024);
  for (int r = 0; r < 32; ++r)
    for (int l = 0; l < 32; ++l) {
      const int t = (l * bitrev5(r)) & 1023;
      tw[r * 32 + l] = make_float2(wr[t], wi[t]);
    }
  for (int g = 0; g < 3; ++g) {
    make_os_filter(g, hr.data(), hi.data());
    for (int r = 0; r < 32; ++r)
      for (int l = 0; l < 32; ++l) hp[(size_t)g * 1024 + r * 32 + l] = make_float2(hr[l + 32 * bitrev5(r)], hi[l + 32 * bitrev5(r)]);
  }
  float2 *d_tw = nullptr, *d_h = nullptr;
  LTB_CUDA(cudaMalloc(&d_tw, sizeof(float2) * tw.size()));
  LTB_CUDA(cudaMalloc(&d_h, sizeof(float2) * hp.size()));
  LTB_CUDA(cudaMemcpy(d_tw, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
  LTB_CUDA(cudaMemcpy(d_h, hp.data(), sizeof(float2) * hp.size(), cudaMemcpyHostToDevice));
  LTB_CUDA(cudaFuncSetAttribute(pss_corr_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OsShared)));
  g_os_h[device] = d_h;
  g_os_tw[device] = d_tw;
  return LTB_SUCCESS;
}

// whole overlap-save blocks [blk_first, blk_first + blk_count) of every stream
int launch_corr_fft(int device, const float2 *y_ring, float *p_ring, long long blk_first, int blk_count, unsigned mask,
                    int cap, int n_streams, cudaStream_t st) {
  if (blk_count <= 0) return LTB_SUCCESS;
  const long long total = (long long)blk_count * n_streams;
  long long ctas = (long long)LTB_OS_MIN_CTAS * (g_sm_count[device] > 0 ? g_sm_count[device] : 148);
  if (ctas > (total + kOsWarps - 1) / kOsWarps) ctas = (total + kOsWarps - 1) / kOsWarps;
  pss_corr_fft_kernel<<<(unsigned)ctas, 32 * kOsWarps, sizeof(OsShared), st>>>(y_ring, p_ring, blk_first, blk_count, mask, cap, n_streams,
                                                                g_os_tw[device], g_os_h[device]);
  return LTB_SUCCESS;
}

int make_cexp_device(float2 **out) {
  std::vector<float> re(4097), im(4097);
  make_cexp_table(re.data(), im.data());
  std::vector<float2> tab(4097);
  for (int i = 0; i < 4097; ++i) tab[i] = make_float2(re[i], im[i]);
  LTB_CUDA(cudaMalloc(out, sizeof(float2) * 4097));
  LTB_CUDA(cudaMemcpy(*out, tab.data(), sizeof(float2) * 4097, cudaMemcpyHostToDevice));
  return LTB_SUCCESS;
}

// debug build only (-DLTB_DEBUG, ltb_debug_set_flag; all zero and constant in the release library):
// [0] decimator dissection bits, [1] bit 0: decimate with the general kernel at
// every rate, bit 1: D = 12..15, 8, 4 with the tiled kernel instead of the streaming one, [2] extra dynamic smem for the tiled decimator (occupancy experiments), [3] unused
#ifdef LTB_DEBUG
int g_debug_flags[4] = {0, 0, 0, 0};
#else
constexpr int g_debug_flags[4] = {0, 0, 0, 0};
#endif

bool valid_decim(int d) { return d >= 1 && d <= kMaxDecim; }
bool valid_format(int f) { return f == LTB_FMT_FC32 || f == LTB_FMT_SC16 || f == LTB_FMT_SC8; }

// branch-major taps [v][33] of rational_resampler_ccc(1, decim) on the device (decimate_any_kernel)
int make_branch_taps_device(int decim, float **out) {
  *out = nullptr;
  if (decim < 2) return LTB_SUCCESS;
  std::vector<float> vq = make_decim_branch_taps(decim);
  if (vq.empty()) return fail(LTB_ERROR, "unexpected decimator tap count");
  LTB_CUDA(cudaMalloc(out, sizeof(float) * vq.size()));
  LTB_CUDA(cudaMemcpy(*out, vq.data(), sizeof(float) * vq.size(), cudaMemcpyHostToDevice));
  return LTB_SUCCESS;
}

// ---- front-end launchers ---------------------------------------------------------------
template <int FMT>
int launch_frontend(int decim, const void *d_iq, long long stride, int n_streams, int m, float2 *tail_old,
                    float2 *tail_new, const float *branch_taps, float2 *y_ring, long long n_base, unsigned mask,
                    int cap, cudaStream_t st, int *launches) {
  const uintptr_t addr_bits = reinterpret_cast<uintptr_t>(d_iq) | (uintptr_t)stride;
  if (addr_bits & (uintptr_t)(fmt_bytes(FMT) - 1))
    return fail(LTB_ERROR_INVALID_INPUTS, "input pointer and row stride must be multiples of the sample size");
  if (decim == 1) {
    int gx = (m / 2 + 255) / 256;
    if (gx > 1024) gx = 1024;
    const int pairs = (addr_bits & (uintptr_t)(2 * fmt_bytes(FMT) - 1)) == 0;    // two samples per load
    ingest_kernel<FMT><<<dim3(gx, n_streams), 256, 0, st>>>(d_iq, stride, m, y_ring, n_base, mask, cap, pairs);
    *launches += 1;
    return LTB_SUCCESS;
  }
  const dim3 grid((m + kDecOut - 1) / kDecOut, n_streams);
  const bool force_any = (g_debug_flags[1] & 1) != 0;      // parity tests: run the general kernel at every rate
  // the streaming kernels move whole segments with 16-byte bulk copies: a base or row stride that is
  // only sample aligned takes the tiled / general kernel instead (same bits, lower rate)
  const bool aligned16 = (addr_bits & 15u) == 0;
  if (decim == 16 && !force_any && aligned16) {
    // streaming variant: two persistent 8-warp CTAs per SM, each a contiguous run of 256-output segments
    int dev = 0;
    LTB_CUDA(cudaGetDevice(&dev));
    const int sps = (m + kStrSeg - 1) / kStrSeg;
    const long long total = (long long)sps * n_streams;
    if (total > 0x7fffffffLL) return fail(LTB_ERROR_INVALID_INPUTS, "too many decimator segments in one call");
    long long ctas = 2LL * (g_sm_count[dev] > 0 ? g_sm_count[dev] : 148);
    if (ctas > total) ctas = total;
    decimate_stream_kernel<FMT><<<(unsigned)ctas, kStrThreads, decim_stream_smem_bytes<FMT>(), st>>>(
        d_iq, stride, m, tail_old, y_ring, n_base, mask, cap, sps, (int)total, g_debug_flags[0]);
  } else if (decim >= 12 && decim <= 15 && !force_any && aligned16 && !(g_debug_flags[1] & 2)) {
    // streaming variant for D = 12..15: the D = 16 kernel with D of sixteen lanes at work
    int dev = 0;
    LTB_CUDA(cudaGetDevice(&dev));
    const int sps = (m + kStrSeg - 1) / kStrSeg;
    const long long total = (long long)sps * n_streams;
    if (total > 0x7fffffffLL) return fail(LTB_ERROR_INVALID_INPUTS, "too many decimator segments in one call");
    long long ctas = 2LL * (g_sm_count[dev] > 0 ? g_sm_count[dev] : 148);
    if (ctas > total) ctas = total;
#define LTB_STREAM12_CASE(D)                                                                                        \
  case D:                                                                                                           \
    decimate_stream12_kernel<FMT, D><<<(unsigned)ctas, kStrThreads, decim_stream12_smem_bytes<FMT, D>(), st>>>(    \
        d_iq, stride, m, tail_old, y_ring, n_base, mask, cap, sps, (int)total, g_debug_flags[0]);                  \
    break;
    switch (decim) { LTB_STREAM12_CASE(12) LTB_STREAM12_CASE(13) LTB_STREAM12_CASE(14) LTB_STREAM12_CASE(15) }
#undef LTB_STREAM12_CASE
  } else if ((decim == 8 || decim == 4) && !force_any && aligned16 && !(g_debug_flags[1] & 2)) {
    // streaming variant for D = 8, 4 (debug flag 1 bit 1 selects the tiled kernel instead)
    int dev = 0;
    LTB_CUDA(cudaGetDevice(&dev));
    const int seg = decim == 8 ? str2_seg<8>() : str2_seg<4>();
    const int sps = (m + seg - 1) / seg;
    const long long total = (long long)sps * n_streams;
    if (total > 0x7fffffffLL) return fail(LTB_ERROR_INVALID_INPUTS, "too many decimator segments in one call");
    long long ctas = 2LL * (g_sm_count[dev] > 0 ? g_sm_count[dev] : 148);
    if (ctas > total) ctas = total;
    if (decim == 8)
      decimate_stream2_kernel<FMT, 8><<<(unsigned)ctas, kStrThreads, decim_stream2_smem_bytes<FMT, 8>(), st>>>(
          d_iq, stride, m, tail_old, y_ring, n_base, mask, cap, sps, (int)total);
    else
      decimate_stream2_kernel<FMT, 4><<<(unsigned)ctas, kStrThreads, decim_stream2_smem_bytes<FMT, 4>(), st>>>(
          d_iq, stride, m, tail_old, y_ring, n_base, mask, cap, sps, (int)total);
  } else if (decim_is_tiled(decim) && !force_any) {
#define LTB_DECIM_CASE(D)                                                                             \
  case D: {                                                                                           \
    decimate_kernel<FMT, D><<<grid, 32 * D, decim_smem_bytes(D) + g_debug_flags[2], st>>>(            \
        d_iq, stride, m, tail_old, y_ring, n_base, mask, cap, n_streams, g_debug_flags[0]);           \
  } break;
    switch (decim) {
      LTB_DECIM_CASE(2) LTB_DECIM_CASE(3) LTB_DECIM_CASE(4) LTB_DECIM_CASE(5) LTB_DECIM_CASE(6) LTB_DECIM_CASE(7)
      LTB_DECIM_CASE(8) LTB_DECIM_CASE(9) LTB_DECIM_CASE(10) LTB_DECIM_CASE(11) LTB_DECIM_CASE(12)
      LTB_DECIM_CASE(13) LTB_DECIM_CASE(14) LTB_DECIM_CASE(15)
      default: return fail(LTB_ERROR_INVALID_INPUTS, "unsupported decimation");
    }
#undef LTB_DECIM_CASE
  } else {
    if (!branch_taps) return fail(LTB_ERROR, "decimator tap table missing");
    const dim3 g2((m + kAnyOut - 1) / kAnyOut, n_streams);
    decimate_any_kernel<FMT><<<g2, kAnyOut, decim_any_smem_bytes(decim), st>>>(
        d_iq, stride, m, decim, branch_taps, tail_old, y_ring, n_base, mask, cap);
  }
  tail_kernel<FMT><<<n_streams, 256, 0, st>>>(d_iq, stride, (long long)m * decim, tail_old, tail_new);
  *launches += 2;
  return LTB_SUCCESS;
}

// ---- LTB_FRONTEND_TC_INT: integer tensor-core front end (sc16 / sc8, decim 16) ----------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
struct TcTable { int8_t *d = nullptr; long long sum_t = 0; };
TcTable g_tc_tab[64][3][LTB_MAX_DECIM + 1];      // per device, input format and rate

// every supported (format, rate) pair; X(fmt, D)
#define LTB_TC_VARIANTS(X)                                                                                    \
  X(LTB_FMT_FC32, 2) X(LTB_FMT_FC32, 4) X(LTB_FMT_FC32, 8) X(LTB_FMT_FC32, 12) X(LTB_FMT_FC32, 16)              \
  X(LTB_FMT_FC32, 24) X(LTB_FMT_FC32, 32)                                                                       \
  X(LTB_FMT_SC16, 4) X(LTB_FMT_SC16, 8) X(LTB_FMT_SC16, 12) X(LTB_FMT_SC16, 16) X(LTB_FMT_SC16, 24) X(LTB_FMT_SC16, 32) \
  X(LTB_FMT_SC8, 8) X(LTB_FMT_SC8, 16) X(LTB_FMT_SC8, 24) X(LTB_FMT_SC8, 32)

int ensure_tc_tables(int device, int fmt, int decim) {
  std::lock_guard<std::mutex> lk(g_const_mu);
  static_assert(LTB_FMT_FC32 == 0 && LTB_FMT_SC16 == 1 && LTB_FMT_SC8 == 2, "tap tables are indexed by format");
  if (!tc_supported(fmt, decim)) return fail(LTB_ERROR_INVALID_INPUTS, "no tensor-core front end for this format and rate");
  TcTable &tt = g_tc_tab[device][fmt][decim];
  if (tt.d) return LTB_SUCCESS;
  if (!g_encode_tiled) {
    cudaDriverEntryPointQueryResult qres;
    void *fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
      return fail(LTB_ERROR, "cuTensorMapEncodeTiled is not available from this driver");
    g_encode_tiled = (EncodeTiledFn)fn;
  }
  long long sum = 0;
  make_tc_taps(decim, &sum);
  const std::vector<int8_t> tab = make_tc_btab(fmt, decim);
  if (tab.empty()) return fail(LTB_ERROR, "decimator taps do not fit three base-256 digits");
  int8_t *d = nullptr;
  LTB_CUDA(cudaMalloc(&d, tab.size()));
  LTB_CUDA(cudaMemcpy(d, tab.data(), tab.size(), cudaMemcpyHostToDevice));
#define X(F, DD) if (fmt == F && decim == DD) LTB_CUDA(cudaFuncSetAttribute(decimate_tc_kernel<F, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes()));
  LTB_TC_VARIANTS(X)
#undef X
  tt.sum_t = sum;
  tt.d = d;
  return LTB_SUCCESS;
}

// n_in new samples per stream (multiple of 8 decim) -> n_in / decim search-rate samples in y_ring; tail_old holds the
// 48 decim raw samples before the chunk, tail_new receives the last 48 decim for the next call.  full_scale: fc32 only
// (the input is taken as 23-bit fixed point of that range, ltb_tc_frontend.cuh)
int launch_frontend_tc(int device, int fmt, int decim, float full_scale, const void *d_iq, long long stride, int n_streams, int n_in,
                       const void *tail_old, void *tail_new, float2 *y_ring, long long n_base, unsigned mask, int cap, int *d_err,
                       cudaStream_t st, int *launches) {
  if ((reinterpret_cast<uintptr_t>(d_iq) | (uintptr_t)stride) & 15u)
    return fail(LTB_ERROR_INVALID_INPUTS, "LTB_FRONTEND_TC_INT needs a 16-byte aligned input pointer and row stride (TMA)");
  if (!tc_supported(fmt, decim) || !g_tc_tab[device][fmt][decim].d) return fail(LTB_ERROR, "tensor-core front end not initialised for this format and rate");
  const int bps = tc_sample_bytes(fmt);
  const int row = 16 * decim;
  const int full_rows = n_in / row;
  CUtensorMap map;
  const cuuint64_t row_bytes = (cuuint64_t)row * bps;
  const cuuint64_t gdim[3] = {row_bytes, (cuuint64_t)(full_rows > 0 ? full_rows : 1), (cuuint64_t)n_streams};
  const cuuint64_t gstr[2] = {row_bytes, (cuuint64_t)stride};
  const cuuint32_t box[3] = {256, (cuuint32_t)kTcTileRows, 1}, estr[3] = {1, 1, 1};
  if (n_streams > 1 && stride < (long long)n_in * bps) return fail(LTB_ERROR_INVALID_INPUTS, "row stride smaller than a row");
  const CUresult r = g_encode_tiled(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(d_iq), gdim, gstr, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LTB_ERROR, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  const TcTable &tt = g_tc_tab[device][fmt][decim];
  TcParams P;
  P.in = d_iq; P.stride_bytes = stride; P.n_in = n_in; P.n_streams = n_streams; P.tail = tail_old; P.y_ring = y_ring;
  P.n_base = n_base; P.cap_mask = mask; P.cap = cap; P.m_out = n_in / decim;
  const int rows = (n_in + row - 1) / row;
  P.tiles_per_stream = (rows + kTcUseful - 1) / kTcUseful;
  const long long total = (long long)P.tiles_per_stream * n_streams;
  if (total > 0x7fffffffLL) return fail(LTB_ERROR_INVALID_INPUTS, "too many decimator tiles in one call");
  P.total_tiles = (int)total;
  P.btab = tt.d; P.err = d_err; P.dbg_acc = nullptr;
  const int shift = tc_tap_shift_for(decim);
  P.c_const = fmt == LTB_FMT_SC16 ? 128 * tt.sum_t : fmt == LTB_FMT_FC32 ? -16384 * tt.sum_t : 0;
  P.q_inv = fmt == LTB_FMT_FC32 ? (float)(0.5 / (double)full_scale) : 0.f;
  P.out_scale = fmt == LTB_FMT_FC32 ? (float)((double)full_scale / 4194303.0 * std::ldexp(1.0, 8 - shift))
                                    : (float)std::ldexp(1.0, -(shift + (fmt == LTB_FMT_SC16 ? 15 : 7)));
  const int sms = g_sm_count[device] > 0 ? g_sm_count[device] : 148;
  const int grid = P.total_tiles < sms ? P.total_tiles : sms;
  bool launched = false;
#define X(F, DD)                                                                                             \
  if (fmt == F && decim == DD) {                                                                             \
    decimate_tc_kernel<F, DD><<<grid, kTcThreads, tc_smem_bytes(), st>>>(map, P);                            \
    tc_tail_kernel<F, DD><<<n_streams, 256, 0, st>>>(d_iq, stride, n_in, tail_old, tail_new);                \
    launched = true;                                                                                         \
  }
  LTB_TC_VARIANTS(X)
#undef X
  if (!launched) return fail(LTB_ERROR, "no tensor-core kernel for this format and rate");
  *launches += 2;
  return LTB_SUCCESS;
}

// format dispatch
int launch_frontend_fmt(int fmt, int decim, const void *d_iq, long long stride, int n_streams, int m, float2 *tail_old,
                        float2 *tail_new, const float *branch_taps, float2 *y_ring, long long n_base, unsigned mask,
                        int cap, cudaStream_t st, int *launches) {
  switch (fmt) {
    case LTB_FMT_FC32: return launch_frontend<LTB_FMT_FC32>(decim, d_iq, stride, n_streams, m, tail_old, tail_new, branch_taps, y_ring, n_base, mask, cap, st, launches);
    case LTB_FMT_SC16: return launch_frontend<LTB_FMT_SC16>(decim, d_iq, stride, n_streams, m, tail_old, tail_new, branch_taps, y_ring, n_base, mask, cap, st, launches);
    case LTB_FMT_SC8:  return launch_frontend<LTB_FMT_SC8>(decim, d_iq, stride, n_streams, m, tail_old, tail_new, branch_taps, y_ring, n_base, mask, cap, st, launches);
    default: return fail(LTB_ERROR_INVALID_INPUTS, "unknown input format");
  }
}

}  // namespace

// ==========================================================================================
// batched trigger engine
// ==========================================================================================
struct ltb_trigger {
  ltb_trigger_config cfg;
  int n_chains = 0, cap = 0, max_m = 0, w_cap = 0, w_cur = 0;
  unsigned cap_mask = 0;
  // Two launch streams.  `stream` (highest priority) carries the stateless part of a call -- front end and
  // all-lag correlator; `track_stream` (lowest priority) the per-chain sequential part -- chain order, track
  // kernel, SSS.  With two calls in flight the track kernel of call i runs while the front end of call i+1
  // does: the front end's CTAs are placed first (priority) and leave registers for one track CTA per SM,
  // which fills issue slots the FFMA2-bound decimator cannot use.  LTB_PIPE_SERIAL puts both on one stream.
  cudaStream_t stream = nullptr;
  cudaStream_t track_stream = nullptr;
  cudaStream_t user_stream = nullptr;     // cfg.cuda_stream: the front end of a call waits for work queued there
  bool overlap = false;
  // up to two submitted-but-not-collected calls: the host enqueues call i+1 while the records
  // of call i are still in flight, so the stream never idles between calls
  struct Slot {
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, done = nullptr;
    cudaEvent_t ev_k[4] = {nullptr, nullptr, nullptr, nullptr};   // after front end, corr, track, sss
    cudaEvent_t ev_b0 = nullptr;          // track stream reaches this call (after waiting for its correlator)
    cudaEvent_t ev_user = nullptr;        // recorded on cfg.cuda_stream at submit
    ltb_window_rec *h_recs = nullptr;     // pinned
    int *h_rec_count = nullptr;           // pinned
    ltb_window_rec *d_recs = nullptr;     // this call's records on the device
    int *d_rec_count = nullptr;
    int w_cur = 0, launches = 0;
  } slot[2];
  int n_pending = 0, head = 0;            // slot[head] is the oldest pending call
  int last = 0;                           // slot of the most recently collected call
  float last_kernel_ms[4] = {0.f, 0.f, 0.f, 0.f};
  void *d_in[2] = {nullptr, nullptr};     // host-input staging, one per call in flight
  size_t d_in_stride = 0;
  cudaStream_t in_stream = nullptr;       // host->device copies, overlap the previous call's kernels
  cudaEvent_t in_ready[2] = {nullptr, nullptr};
  float2 *d_y = nullptr;
  float *d_p = nullptr;
  ChainState *d_state = nullptr;
  float *d_avg = nullptr;
  float *d_thr = nullptr;
  cudaStream_t copy_stream = nullptr;     // record read-back, overlaps the next call's kernels
  float2 *d_sss_sym = nullptr;
  int *d_sss_rec = nullptr;
  int *d_sss_count = nullptr;     // [0] SSS candidate count, [1] track-kernel chain queue head, [2..3] chain_order_kernel
  int *d_chain_order = nullptr;   // [n_chains]
  int sss_cap = 0;
  float2 *d_hf = nullptr;
  float2 *d_tail[2] = {nullptr, nullptr};
  int tail_cur = 0;
  void *d_tc_tail[2] = {nullptr, nullptr};     // LTB_FRONTEND_TC_INT: raw history, [n_streams][768] samples
  int *d_tc_err = nullptr;
  float2 *d_cexp = nullptr;
  float *d_branch_taps = nullptr;         // [decim][33], decimate_any_kernel
  long long n_total = 0;
  long long os_blocks_done = 0;           // LTB_CORR_FFT: overlap-save blocks already evaluated
  std::vector<float> h_thr;
  int last_launches = 0;
  float last_ms = 0.f;
};

namespace {

int trigger_zero_state(ltb_trigger *t) {
  const int S = t->cfg.n_streams;
  LTB_CUDA(cudaMemsetAsync(t->d_y, 0, sizeof(float2) * (size_t)S * t->cap, t->stream));
  LTB_CUDA(cudaMemsetAsync(t->d_p, 0, sizeof(float) * (size_t)S * 3 * t->cap, t->stream));
  LTB_CUDA(cudaMemsetAsync(t->d_state, 0, sizeof(ChainState) * t->n_chains, t->stream));
  LTB_CUDA(cudaMemsetAsync(t->d_avg, 0, sizeof(float) * (size_t)t->n_chains * kAvgLen, t->stream));
  LTB_CUDA(cudaMemsetAsync(t->d_tail[0], 0, sizeof(float2) * (size_t)S * kTailCap, t->stream));
  LTB_CUDA(cudaMemsetAsync(t->d_tail[1], 0, sizeof(float2) * (size_t)S * kTailCap, t->stream));
  if (t->d_tc_tail[0]) {
    LTB_CUDA(cudaMemsetAsync(t->d_tc_tail[0], 0, 8 * (size_t)S * kTcTailSamples, t->stream));
    LTB_CUDA(cudaMemsetAsync(t->d_tc_tail[1], 0, 8 * (size_t)S * kTcTailSamples, t->stream));
    LTB_CUDA(cudaMemsetAsync(t->d_tc_err, 0, sizeof(int), t->stream));
  }
  LTB_CUDA(cudaMemcpyAsync(t->d_thr, t->h_thr.data(), sizeof(float) * t->n_chains, cudaMemcpyHostToDevice, t->stream));
  LTB_CUDA(cudaStreamSynchronize(t->stream));
  LTB_CUDA(cudaStreamSynchronize(t->track_stream));
  t->n_total = 0;
  t->os_blocks_done = 0;
  t->tail_cur = 0;
  t->n_pending = 0; t->head = 0;
  return LTB_SUCCESS;
}

void trigger_free(ltb_trigger *t) {
  if (!t) return;
  cudaSetDevice(t->cfg.device);
  cudaFree(t->d_in[0]); cudaFree(t->d_in[1]); cudaFree(t->d_y); cudaFree(t->d_p); cudaFree(t->d_state); cudaFree(t->d_avg);
  cudaFree(t->d_thr); cudaFree(t->d_sss_sym);
  cudaFree(t->d_sss_rec); cudaFree(t->d_sss_count); cudaFree(t->d_chain_order); cudaFree(t->d_hf); cudaFree(t->d_tail[0]);
  cudaFree(t->d_tail[1]); cudaFree(t->d_cexp); cudaFree(t->d_branch_taps);
  cudaFree(t->d_tc_tail[0]); cudaFree(t->d_tc_tail[1]); cudaFree(t->d_tc_err);
  for (auto &sl : t->slot) {
    cudaFree(sl.d_recs); cudaFree(sl.d_rec_count);
    if (sl.h_recs) cudaFreeHost(sl.h_recs);
    if (sl.h_rec_count) cudaFreeHost(sl.h_rec_count);
    if (sl.ev0) cudaEventDestroy(sl.ev0);
    if (sl.ev1) cudaEventDestroy(sl.ev1);
    if (sl.done) cudaEventDestroy(sl.done);
    for (int i = 0; i < 4; ++i) if (sl.ev_k[i]) cudaEventDestroy(sl.ev_k[i]);
    if (sl.ev_b0) cudaEventDestroy(sl.ev_b0);
    if (sl.ev_user) cudaEventDestroy(sl.ev_user);
  }
  if (t->copy_stream) cudaStreamDestroy(t->copy_stream);
  if (t->in_stream) cudaStreamDestroy(t->in_stream);
  for (auto &e : t->in_ready) if (e) cudaEventDestroy(e);
  if (t->track_stream && t->track_stream != t->stream) cudaStreamDestroy(t->track_stream);
  if (t->stream) cudaStreamDestroy(t->stream);
  delete t;
}

int trigger_enqueue(ltb_trigger *t, const void *d_iq, long long stride, long long n_samples) {
  const ltb_trigger_config &c = t->cfg;
  // emitted half-frames live in one device buffer, so keep_halfframes allows one call in flight
  if (t->n_pending >= (c.keep_halfframes ? 1 : 2)) return fail(LTB_ERROR_INVALID_INPUTS, "too many submits not collected");
  ltb_trigger::Slot &sl = t->slot[(t->head + t->n_pending) & 1];
  if (!d_iq || n_samples <= 0 || n_samples > c.max_chunk || (n_samples % (8 * c.decim)) != 0)
    return fail(LTB_ERROR_INVALID_INPUTS, "n_samples must be a positive multiple of 8*decim and <= max_chunk");
  const int m = (int)(n_samples / c.decim);
  const int S = c.n_streams;
  const long long n_base = t->n_total;
  int launches = 0;
  if (t->user_stream) {                    // inputs produced on the caller's stream are ordered before our reads
    LTB_CUDA(cudaEventRecord(sl.ev_user, t->user_stream));
    LTB_CUDA(cudaStreamWaitEvent(t->stream, sl.ev_user, 0));
  }
  LTB_CUDA(cudaEventRecord(sl.ev0, t->stream));
  int rc;
  if (c.frontend_mode == LTB_FRONTEND_TC_INT)
    rc = launch_frontend_tc(c.device, c.input_format, c.decim, c.fc32_full_scale, d_iq, stride, S, (int)n_samples, t->d_tc_tail[t->tail_cur], t->d_tc_tail[t->tail_cur ^ 1],
                            t->d_y, n_base, t->cap_mask, t->cap, t->d_tc_err, t->stream, &launches);
  else
    rc = launch_frontend_fmt(c.input_format, c.decim, d_iq, stride, S, m, t->d_tail[t->tail_cur], t->d_tail[t->tail_cur ^ 1],
                             t->d_branch_taps, t->d_y, n_base, t->cap_mask, t->cap, t->stream, &launches);
  if (rc) return rc;
  if (c.decim > 1) t->tail_cur ^= 1;
  LTB_CUDA(cudaEventRecord(sl.ev_k[0], t->stream));
  if (c.corr_mode == LTB_CORR_FFT) {
    // whole 896-output blocks that end inside the received samples; the tail waits for the next call
    const long long blk_end = (n_base + m) / kOsStep;
    rc = launch_corr_fft(c.device, t->d_y, t->d_p, t->os_blocks_done, (int)(blk_end - t->os_blocks_done), t->cap_mask, t->cap,
                         S, t->stream);
    if (rc) return rc;
    t->os_blocks_done = blk_end;
  } else {
    const int tps = (m + kCorrTile - 1) / kCorrTile;
    const long long total = (long long)tps * S;
    if (total > 0x7fffffffLL) return fail(LTB_ERROR_INVALID_INPUTS, "too many correlator tiles in one call");
    long long ctas = 2LL * (g_sm_count[c.device] > 0 ? g_sm_count[c.device] : 148);
    if (ctas > total) ctas = total;
    pss_corr_kernel<<<(unsigned)ctas, kCorrThreads, 0, t->stream>>>(t->d_y, t->d_p, n_base, m, t->cap_mask, t->cap, tps, (int)total);
  }
  launches++;
  LTB_CUDA(cudaEventRecord(sl.ev_k[1], t->stream));
  t->n_total += m;
  t->w_cur = m / (kHalf - kSlot) + 4;
  if (t->w_cur > t->w_cap) t->w_cur = t->w_cap;
  // ---- the sequential part, on the track stream: it needs this call's correlator, nothing of the next call ----
  cudaStream_t ts = t->track_stream;
  if (ts != t->stream) LTB_CUDA(cudaStreamWaitEvent(ts, sl.ev_k[1], 0));
  LTB_CUDA(cudaEventRecord(sl.ev_b0, ts));
  LTB_CUDA(cudaMemsetAsync(t->d_sss_count, 0, 4 * sizeof(int), ts));
  chain_order_kernel<<<(t->n_chains + 255) / 256, 256, 0, ts>>>(t->d_state, t->n_chains, t->d_chain_order, t->d_sss_count + 2);
  launches++;
  TrackParams P;
  P.y_ring = t->d_y; P.p_ring = t->d_p; P.state = t->d_state; P.avg = t->d_avg; P.thr = t->d_thr;
  P.recs = sl.d_recs; P.rec_count = sl.d_rec_count; P.sss_sym = t->d_sss_sym; P.sss_rec = t->d_sss_rec;
  P.sss_count = t->d_sss_count; P.sss_cap = t->sss_cap; P.hf_out = t->d_hf; P.cexp = t->d_cexp;
  P.n_total = t->n_total; P.cap_mask = t->cap_mask; P.cap = t->cap; P.w_max = t->w_cur;
  P.track_after = c.track_after; P.track_every = c.track_every; P.record_all = c.record_all;
  P.root_mask = c.root_mask;
  P.tdd = c.frame_type == LTB_FRAME_TDD;
  P.chain_counter = t->d_sss_count + 1; P.n_chains = t->n_chains;
  P.chain_order = t->d_chain_order;
  {
    int ctas = 4 * (g_sm_count[c.device] > 0 ? g_sm_count[c.device] : 148);
    if (ctas > t->n_chains) ctas = t->n_chains;
    pss_track_kernel<<<ctas, kTrackThreads, sizeof(TrackShared), ts>>>(P);
  }
  launches++;
  LTB_CUDA(cudaEventRecord(sl.ev_k[2], ts));
  int sss_grid = (t->n_chains * t->w_cur + kSssWarps - 1) / kSssWarps;
  if (sss_grid > 148 * 8) sss_grid = 148 * 8;
  sss_kernel<<<sss_grid, kSssWarps * 32, 0, ts>>>(t->d_sss_sym, t->d_sss_rec, t->d_sss_count, t->sss_cap, sl.d_recs);
  launches++;
  LTB_CUDA(cudaEventRecord(sl.ev1, ts));
  // read the records back on the copy stream: the next call's kernels need not wait for it
  LTB_CUDA(cudaStreamWaitEvent(t->copy_stream, sl.ev1, 0));
  LTB_CUDA(cudaMemcpyAsync(sl.h_rec_count, sl.d_rec_count, sizeof(int) * t->n_chains, cudaMemcpyDeviceToHost, t->copy_stream));
  LTB_CUDA(cudaMemcpyAsync(sl.h_recs, sl.d_recs, sizeof(ltb_window_rec) * (size_t)t->n_chains * t->w_cur,
                           cudaMemcpyDeviceToHost, t->copy_stream));
  LTB_CUDA(cudaEventRecord(sl.done, t->copy_stream));
  LTB_CUDA(cudaGetLastError());
  sl.launches = launches;
  sl.w_cur = t->w_cur;
  t->n_pending++;
  return LTB_SUCCESS;
}

}  // namespace

extern "C" {

const char *ltb_last_error(void) { return g_err.c_str(); }
const char *ltb_version(void) { return "ltetrigger_b200 0.1 (sm_100a)"; }

int ltb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int ltb_trigger_create(const ltb_trigger_config *cfg, ltb_trigger **out) {
  // ABI evolution: a caller built against the header before corr_mode was appended passes the
  // shorter size and gets the defaults of the fields it does not know
  constexpr size_t kMinConfig = offsetof(ltb_trigger_config, corr_mode);   // later fields default to 0
  if (!cfg || !out || cfg->struct_size < kMinConfig || cfg->struct_size > sizeof(ltb_trigger_config))
    return fail(LTB_ERROR_INVALID_INPUTS, "bad config pointer or struct_size");
  *out = nullptr;
  ltb_trigger_config c = ltb_trigger_config();
  std::memcpy(&c, cfg, cfg->struct_size);
  c.struct_size = sizeof c;
  if (c.n_streams <= 0 || !valid_decim(c.decim) || !valid_format(c.input_format) ||
      c.max_chunk <= 0 || (c.root_mask & ~7) || (c.corr_mode != LTB_CORR_DIRECT && c.corr_mode != LTB_CORR_FFT) ||
      (c.frame_type != LTB_FRAME_FDD && c.frame_type != LTB_FRAME_TDD) ||
      (c.pipeline != LTB_PIPE_OVERLAP && c.pipeline != LTB_PIPE_SERIAL) ||
      (c.frontend_mode != LTB_FRONTEND_FP32 && c.frontend_mode != LTB_FRONTEND_TC_INT))
    return fail(LTB_ERROR_INVALID_INPUTS, "invalid trigger configuration");
  if (c.frontend_mode == LTB_FRONTEND_TC_INT && !tc_supported(c.input_format, c.decim))
    return fail(LTB_ERROR_INVALID_INPUTS, "LTB_FRONTEND_TC_INT is available at decim 2 / 4 / 8 / 12 / 16 / 24 / 32 for fc32, from 4 for sc16 "
                                          "and from 8 for sc8 input (a 16-output row must be whole 256-byte pieces)");
  if (c.frontend_mode == LTB_FRONTEND_TC_INT && c.input_format == LTB_FMT_FC32 &&
      !(c.fc32_full_scale > 0.f && c.fc32_full_scale < 1e30f))
    return fail(LTB_ERROR_INVALID_INPUTS, "LTB_FRONTEND_TC_INT on fc32 input takes the samples as 23-bit fixed point: "
                                          "set fc32_full_scale > 0 (the largest |re|, |im| the source delivers)");
  if (c.root_mask == 0) c.root_mask = 7;
  if (c.track_after <= 0) c.track_after = 16;
  if (c.track_every <= 0) c.track_every = 8;
  if (!(c.psr_threshold > LTB_MIN_PSR_THRESHOLD)) c.psr_threshold = LTB_MIN_PSR_THRESHOLD;   // _ensure_safe_threshold
  c.max_chunk = (c.max_chunk + 8 * c.decim - 1) / (8 * c.decim) * (8 * c.decim);
  if (ltb_device_count() <= c.device || c.device < 0) return fail(LTB_ERROR, "no such CUDA device");
  LTB_CUDA(cudaSetDevice(c.device));
  int rc = ensure_constants(c.device);
  if (!rc && c.corr_mode == LTB_CORR_FFT) rc = ensure_os_tables(c.device);
  if (!rc && c.frontend_mode == LTB_FRONTEND_TC_INT) rc = ensure_tc_tables(c.device, c.input_format, c.decim);
  if (rc) return rc;

  ltb_trigger *t = new ltb_trigger();
  t->cfg = c;
  const int S = c.n_streams;
  t->n_chains = S * 3;
  t->max_m = (int)(c.max_chunk / c.decim);
  t->overlap = c.pipeline == LTB_PIPE_OVERLAP && !c.keep_halfframes;
  // the rings are indexed by absolute sample number.  One call in flight: the chunk being written plus what the
  // oldest unfinished window still reads.  Overlapped: the front end of call i+1 writes while the track kernel
  // of call i reads back to its chains' oldest window (about one chunk + a look-ahead + a half-frame behind).
  t->cap = next_pow2((long long)t->max_m * (t->overlap ? 2 : 1) + kLookahead + (t->overlap ? 2 * kHalf : 0) + kSlot + 256);
  t->cap_mask = (unsigned)(t->cap - 1);
  t->w_cap = t->max_m / (kHalf - kSlot) + 4;
  t->sss_cap = t->n_chains * t->w_cap;
  t->h_thr.assign(t->n_chains, c.psr_threshold);
#define LTB_CUDA_T(call)                                                                   \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      trigger_free(t);                                                                     \
      return fail(LTB_ERROR, std::string(#call) + ": " + cudaGetErrorString(e__));          \
    }                                                                                      \
  } while (0)
  {
    int prio_least = 0, prio_greatest = 0;
    LTB_CUDA_T(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    t->user_stream = (cudaStream_t)c.cuda_stream;
    LTB_CUDA_T(cudaStreamCreateWithPriority(&t->stream, cudaStreamNonBlocking, prio_greatest));
    if (t->overlap) LTB_CUDA_T(cudaStreamCreateWithPriority(&t->track_stream, cudaStreamNonBlocking, prio_least));
    else t->track_stream = t->stream;
  }
  for (auto &sl : t->slot) {
    LTB_CUDA_T(cudaEventCreate(&sl.ev_b0));
    LTB_CUDA_T(cudaEventCreateWithFlags(&sl.ev_user, cudaEventDisableTiming));
    LTB_CUDA_T(cudaEventCreate(&sl.ev0));
    LTB_CUDA_T(cudaEventCreate(&sl.ev1));
    LTB_CUDA_T(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    for (int i = 0; i < 4; ++i) LTB_CUDA_T(cudaEventCreate(&sl.ev_k[i]));
  }
  t->d_in_stride = (size_t)c.max_chunk * fmt_bytes(c.input_format);
  LTB_CUDA_T(cudaMalloc(&t->d_y, sizeof(float2) * (size_t)S * t->cap));
  LTB_CUDA_T(cudaMalloc(&t->d_p, sizeof(float) * (size_t)S * 3 * t->cap));
  LTB_CUDA_T(cudaMalloc(&t->d_state, sizeof(ChainState) * t->n_chains));
  LTB_CUDA_T(cudaMalloc(&t->d_avg, sizeof(float) * (size_t)t->n_chains * kAvgLen));
  LTB_CUDA_T(cudaMalloc(&t->d_thr, sizeof(float) * t->n_chains));
  LTB_CUDA_T(cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking));
  LTB_CUDA_T(cudaStreamCreateWithFlags(&t->in_stream, cudaStreamNonBlocking));
  for (auto &e : t->in_ready) LTB_CUDA_T(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto &sl : t->slot) {
    LTB_CUDA_T(cudaMalloc(&sl.d_recs, sizeof(ltb_window_rec) * (size_t)t->n_chains * t->w_cap));
    LTB_CUDA_T(cudaMalloc(&sl.d_rec_count, sizeof(int) * t->n_chains));
  }
  LTB_CUDA_T(cudaMalloc(&t->d_sss_sym, sizeof(float2) * 128 * (size_t)t->sss_cap));
  LTB_CUDA_T(cudaMalloc(&t->d_sss_rec, sizeof(int) * t->sss_cap));
  LTB_CUDA_T(cudaMalloc(&t->d_sss_count, 4 * sizeof(int)));
  LTB_CUDA_T(cudaMalloc(&t->d_chain_order, sizeof(int) * t->n_chains));
  LTB_CUDA_T(cudaMalloc(&t->d_tail[0], sizeof(float2) * (size_t)S * kTailCap));
  LTB_CUDA_T(cudaMalloc(&t->d_tail[1], sizeof(float2) * (size_t)S * kTailCap));
  if (c.frontend_mode == LTB_FRONTEND_TC_INT) {
    LTB_CUDA_T(cudaMalloc(&t->d_tc_tail[0], 8 * (size_t)S * kTcTailSamples));
    LTB_CUDA_T(cudaMalloc(&t->d_tc_tail[1], 8 * (size_t)S * kTcTailSamples));
    LTB_CUDA_T(cudaMalloc(&t->d_tc_err, sizeof(int)));
  }
  if (c.keep_halfframes) LTB_CUDA_T(cudaMalloc(&t->d_hf, sizeof(float2) * kHalf * (size_t)t->n_chains * t->w_cap));
  for (auto &sl : t->slot) {
    LTB_CUDA_T(cudaMallocHost(&sl.h_recs, sizeof(ltb_window_rec) * (size_t)t->n_chains * t->w_cap));
    LTB_CUDA_T(cudaMallocHost(&sl.h_rec_count, sizeof(int) * t->n_chains));
  }
#undef LTB_CUDA_T
  rc = make_cexp_device(&t->d_cexp);
  if (!rc) rc = make_branch_taps_device(c.decim, &t->d_branch_taps);
  if (!rc) rc = trigger_zero_state(t);
  if (rc) { trigger_free(t); return rc; }
  *out = t;
  return LTB_SUCCESS;
}

int ltb_trigger_destroy(ltb_trigger *t) {
  if (!t) return LTB_ERROR_INVALID_INPUTS;
  cudaSetDevice(t->cfg.device);
  if (t->in_stream) cudaStreamSynchronize(t->in_stream);
  cudaStreamSynchronize(t->stream);
  if (t->track_stream) cudaStreamSynchronize(t->track_stream);
  if (t->copy_stream) cudaStreamSynchronize(t->copy_stream);
  trigger_free(t);
  return LTB_SUCCESS;
}

int ltb_trigger_reset(ltb_trigger *t) {
  if (!t) return LTB_ERROR_INVALID_INPUTS;
  LTB_CUDA(cudaSetDevice(t->cfg.device));
  LTB_CUDA(cudaStreamSynchronize(t->in_stream));
  LTB_CUDA(cudaStreamSynchronize(t->stream));
  LTB_CUDA(cudaStreamSynchronize(t->track_stream));
  LTB_CUDA(cudaStreamSynchronize(t->copy_stream));
  return trigger_zero_state(t);
}

int ltb_trigger_set_psr_threshold(ltb_trigger *t, int stream, int n_id_2, float thr, int clamp) {
  if (!t || stream < -1 || stream >= t->cfg.n_streams || n_id_2 < -1 || n_id_2 > 2)
    return fail(LTB_ERROR_INVALID_INPUTS, "bad chain selector");
  if (clamp && !(thr > LTB_MIN_PSR_THRESHOLD)) thr = LTB_MIN_PSR_THRESHOLD;
  for (int s = 0; s < t->cfg.n_streams; ++s)
    for (int r = 0; r < 3; ++r)
      if ((stream < 0 || stream == s) && (n_id_2 < 0 || n_id_2 == r)) t->h_thr[s * 3 + r] = thr;
  LTB_CUDA(cudaSetDevice(t->cfg.device));
  // the track kernel reads d_thr: the new value applies to the calls submitted after this returns
  LTB_CUDA(cudaStreamSynchronize(t->stream));
  LTB_CUDA(cudaMemcpyAsync(t->d_thr, t->h_thr.data(), sizeof(float) * t->n_chains, cudaMemcpyHostToDevice, t->track_stream));
  LTB_CUDA(cudaStreamSynchronize(t->track_stream));
  return LTB_SUCCESS;
}

int ltb_trigger_submit_device(ltb_trigger *t, const void *d_iq, int64_t stride, int64_t n_samples) {
  if (!t) return LTB_ERROR_INVALID_INPUTS;
  LTB_CUDA(cudaSetDevice(t->cfg.device));
  return trigger_enqueue(t, d_iq, stride, n_samples);
}

int ltb_trigger_collect(ltb_trigger *t, ltb_window_rec *recs, int max_recs, int *n_recs) {
  if (!t || !n_recs || (!recs && max_recs > 0)) return LTB_ERROR_INVALID_INPUTS;
  if (t->n_pending == 0) return fail(LTB_ERROR_INVALID_INPUTS, "nothing submitted");
  LTB_CUDA(cudaSetDevice(t->cfg.device));
  ltb_trigger::Slot &sl = t->slot[t->head];
  LTB_CUDA(cudaEventSynchronize(sl.done));
  // count first: a buffer that is too small leaves the call pending, so the caller can retry with
  // *n_recs entries instead of losing the records of a call whose chain state has already advanced
  int total = 0;
  for (int ch = 0; ch < t->n_chains; ++ch) total += sl.h_rec_count[ch];
  *n_recs = total;
  if (total > max_recs) return fail(LTB_ERROR_INVALID_INPUTS, "record buffer too small: *n_recs entries are needed; the call stays pending");
  t->last = t->head;
  t->head ^= 1;
  t->n_pending--;
  t->last_launches = sl.launches;
  cudaEventElapsedTime(&t->last_ms, sl.ev0, sl.ev1);
  cudaEventElapsedTime(&t->last_kernel_ms[0], sl.ev0, sl.ev_k[0]);
  cudaEventElapsedTime(&t->last_kernel_ms[1], sl.ev_k[0], sl.ev_k[1]);
  cudaEventElapsedTime(&t->last_kernel_ms[2], sl.ev_b0, sl.ev_k[2]);
  cudaEventElapsedTime(&t->last_kernel_ms[3], sl.ev_k[2], sl.ev1);
  int written = 0;
  for (int ch = 0; ch < t->n_chains; ++ch) {
    const int n = sl.h_rec_count[ch];
    const ltb_window_rec *src = sl.h_recs + (size_t)ch * sl.w_cur;
    for (int i = 0; i < n; ++i) recs[written++] = src[i];
  }
  return LTB_SUCCESS;
}

int ltb_trigger_process_device(ltb_trigger *t, const void *d_iq, int64_t stride, int64_t n_samples,
                               ltb_window_rec *recs, int max_recs, int *n_recs) {
  if (t && t->n_pending) return fail(LTB_ERROR_INVALID_INPUTS, "collect the submitted calls before a synchronous process call");
  int rc = ltb_trigger_submit_device(t, d_iq, stride, n_samples);
  if (rc) return rc;
  return ltb_trigger_collect(t, recs, max_recs, n_recs);
}

int ltb_trigger_submit_host(ltb_trigger *t, const void *iq, int64_t stride, int64_t n_samples) {
  if (!t || !iq) return LTB_ERROR_INVALID_INPUTS;
  LTB_CUDA(cudaSetDevice(t->cfg.device));
  if (n_samples <= 0 || n_samples > t->cfg.max_chunk) return fail(LTB_ERROR_INVALID_INPUTS, "n_samples out of range");
  if (t->n_pending >= (t->cfg.keep_halfframes ? 1 : 2)) return fail(LTB_ERROR_INVALID_INPUTS, "too many submits not collected");
  // staging buffer of the slot this call will occupy; the call that used it last has been
  // collected (at most two in flight), so its front end is done with it
  const int slot = (t->head + t->n_pending) & 1;
  if (!t->d_in[slot]) LTB_CUDA(cudaMalloc(&t->d_in[slot], t->d_in_stride * (size_t)t->cfg.n_streams));
  const size_t row = (size_t)n_samples * fmt_bytes(t->cfg.input_format);
  if (t->cfg.n_streams == 1) {
    // one row: the stride is irrelevant (and may be 0), which a pitched copy would reject
    LTB_CUDA(cudaMemcpyAsync(t->d_in[slot], iq, row, cudaMemcpyHostToDevice, t->in_stream));
  } else {
    if ((size_t)stride < row) return fail(LTB_ERROR_INVALID_INPUTS, "stream_stride_bytes is smaller than one row of n_samples");
    LTB_CUDA(cudaMemcpy2DAsync(t->d_in[slot], t->d_in_stride, iq, (size_t)stride, row, (size_t)t->cfg.n_streams,
                               cudaMemcpyHostToDevice, t->in_stream));
  }
  LTB_CUDA(cudaEventRecord(t->in_ready[slot], t->in_stream));
  LTB_CUDA(cudaStreamWaitEvent(t->stream, t->in_ready[slot], 0));
  return trigger_enqueue(t, t->d_in[slot], (long long)t->d_in_stride, n_samples);
}

int ltb_trigger_process_host(ltb_trigger *t, const void *iq, int64_t stride, int64_t n_samples,
                             ltb_window_rec *recs, int max_recs, int *n_recs) {
  if (t && t->n_pending) return fail(LTB_ERROR_INVALID_INPUTS, "collect the submitted calls before a synchronous process call");
  int rc = ltb_trigger_submit_host(t, iq, stride, n_samples);
  if (rc) return rc;
  return ltb_trigger_collect(t, recs, max_recs, n_recs);
}

int ltb_trigger_get_stats(ltb_trigger *t, int stream, int n_id_2, ltb_pss_stats *out) {
  if (!t || !out || stream < 0 || stream >= t->cfg.n_streams || n_id_2 < 0 || n_id_2 > 2)
    return fail(LTB_ERROR_INVALID_INPUTS, "bad chain selector");
  LTB_CUDA(cudaSetDevice(t->cfg.device));
  ChainState st;
  LTB_CUDA(cudaStreamSynchronize(t->stream));
  LTB_CUDA(cudaMemcpyAsync(&st, t->d_state + (stream * 3 + n_id_2), sizeof st, cudaMemcpyDeviceToHost, t->track_stream));
  LTB_CUDA(cudaStreamSynchronize(t->track_stream));
  auto mean = [](const float *d, unsigned n) -> float {      // compute_moving_avg, lib/pss_impl.cc:94-109
    if (!n) return 0.0f;
    if (n > (unsigned)kMavg) n = kMavg;
    double acc = 0.0;
    for (unsigned i = 0; i < n; ++i) acc += d[i];
    return (float)(acc / (double)n);
  };
  out->max_psr = st.psr_max;
  out->mean_psr = mean(st.psr_data, st.psr_i);
  out->mean_cfo = mean(st.cfo_data, st.cfo_i);
  out->psr_threshold = t->h_thr[stream * 3 + n_id_2];
  out->tracking_score = (float)st.score;
  out->tracking = st.tracking;
  out->next_window = st.next_win;
  return LTB_SUCCESS;
}

int ltb_trigger_fetch_halfframes(ltb_trigger *t, ltb_cf *out, int max_hf, int *n_hf) {
  if (!t || !n_hf || !t->d_hf) return fail(LTB_ERROR_INVALID_INPUTS, "keep_halfframes not enabled");
  LTB_CUDA(cudaSetDevice(t->cfg.device));
  LTB_CUDA(cudaStreamSynchronize(t->stream));
  int total = 0;
  for (int ch = 0; ch < t->n_chains; ++ch) {
    const ltb_trigger::Slot &sl = t->slot[t->last];
    const int n = sl.h_rec_count[ch];
    for (int i = 0; i < n; ++i) {
      const ltb_window_rec &r = sl.h_recs[(size_t)ch * sl.w_cur + i];
      if (!(r.flags & LTB_F_EMIT)) continue;
      if (total < max_hf)
        LTB_CUDA(cudaMemcpyAsync(out + (size_t)total * kHalf, t->d_hf + ((size_t)ch * sl.w_cur + i) * kHalf,
                                 sizeof(float2) * kHalf, cudaMemcpyDeviceToHost, t->stream));
      total++;
    }
  }
  LTB_CUDA(cudaStreamSynchronize(t->stream));
  *n_hf = total;
  return total > max_hf ? fail(LTB_ERROR_INVALID_INPUTS, "half-frame buffer too small") : LTB_SUCCESS;
}

int ltb_trigger_last_timing(ltb_trigger *t, float *ms_total, int *n_launches) {
  if (!t) return LTB_ERROR_INVALID_INPUTS;
  if (ms_total) *ms_total = t->last_ms;
  if (n_launches) *n_launches = t->last_launches;
  return LTB_SUCCESS;
}

#ifdef LTB_DEBUG
// only in lib/libltetrigger_b200_debug.so (make debug): the release library has no way to set the flags
int ltb_debug_set_flag(int flag, int value) {
  if (flag < 0 || flag >= 4) return LTB_ERROR_INVALID_INPUTS;
  g_debug_flags[flag] = value;
  return LTB_SUCCESS;
}
#endif

int ltb_trigger_last_kernel_times(ltb_trigger *t, float ms[4]) {
  if (!t || !ms) return LTB_ERROR_INVALID_INPUTS;
  for (int i = 0; i < 4; ++i) ms[i] = t->last_kernel_ms[i];
  return LTB_SUCCESS;
}

// ==========================================================================================
// standalone sss block
// ==========================================================================================
struct ltb_sss {
  int device, n_id_2;
  int frame_type = LTB_FRAME_FDD;
  float *d_cp = nullptr;
  int *d_count = nullptr;
};

int ltb_sss_create(int device, int n_id_2, ltb_sss **out) {
  if (!out || n_id_2 < 0 || n_id_2 > 2) return fail(LTB_ERROR_INVALID_INPUTS, "Error initializing SSS N_id_2");
  *out = nullptr;
  if (ltb_device_count() <= device || device < 0) return fail(LTB_ERROR, "no such CUDA device");
  LTB_CUDA(cudaSetDevice(device));
  int rc = ensure_constants(device);
  if (rc) return rc;
  ltb_sss *s = new ltb_sss();
  s->device = device; s->n_id_2 = n_id_2;
  if (cudaMalloc(&s->d_cp, 2 * sizeof(float)) != cudaSuccess || cudaMalloc(&s->d_count, sizeof(int)) != cudaSuccess ||
      cudaMemset(s->d_cp, 0, 2 * sizeof(float)) != cudaSuccess) {
    cudaFree(s->d_cp); cudaFree(s->d_count); delete s;
    return fail(LTB_ERROR, "Error initializing SSS SYNC");
  }
  *out = s;
  return LTB_SUCCESS;
}

int ltb_sss_destroy(ltb_sss *s) {
  if (!s) return LTB_ERROR_INVALID_INPUTS;
  cudaSetDevice(s->device);
  cudaFree(s->d_cp); cudaFree(s->d_count);
  delete s;
  return LTB_SUCCESS;
}

int ltb_sss_set_frame_type(ltb_sss *s, int frame_type) {
  if (!s || (frame_type != LTB_FRAME_FDD && frame_type != LTB_FRAME_TDD)) return LTB_ERROR_INVALID_INPUTS;
  s->frame_type = frame_type;
  return LTB_SUCCESS;
}

int ltb_sss_work(ltb_sss *s, const ltb_cf *in, const int32_t *tag_lost, int n, ltb_window_rec *recs) {
  if (!s || !in || !tag_lost || !recs || n <= 0) return LTB_ERROR_INVALID_INPUTS;
  LTB_CUDA(cudaSetDevice(s->device));
  float2 *d_hf = nullptr, *d_sym = nullptr; int *d_tag = nullptr, *d_rec = nullptr; ltb_window_rec *d_recs = nullptr;
  int rc = LTB_SUCCESS;
  cudaError_t e = cudaSuccess;
#define STEP(call) if (e == cudaSuccess) e = (call)
  STEP(cudaMalloc(&d_hf, sizeof(float2) * kHalf * (size_t)n));
  STEP(cudaMalloc(&d_sym, sizeof(float2) * 128 * (size_t)n));
  STEP(cudaMalloc(&d_tag, sizeof(int) * n));
  STEP(cudaMalloc(&d_rec, sizeof(int) * n));
  STEP(cudaMalloc(&d_recs, sizeof(ltb_window_rec) * n));
  STEP(cudaMemcpy(d_hf, in, sizeof(float2) * kHalf * (size_t)n, cudaMemcpyHostToDevice));
  STEP(cudaMemcpy(d_tag, tag_lost, sizeof(int) * n, cudaMemcpyHostToDevice));
  STEP(cudaMemcpy(d_recs, recs, sizeof(ltb_window_rec) * n, cudaMemcpyHostToDevice));
  STEP(cudaMemset(s->d_count, 0, sizeof(int)));
  if (e == cudaSuccess) {
    sss_block_front_kernel<<<1, 128>>>(d_hf, d_tag, n, s->n_id_2, s->frame_type == LTB_FRAME_TDD, s->d_cp, d_recs, d_sym, d_rec, s->d_count);
    sss_kernel<<<(n + kSssWarps - 1) / kSssWarps, kSssWarps * 32>>>(d_sym, d_rec, s->d_count, n, d_recs);
    e = cudaGetLastError();
  }
  STEP(cudaMemcpy(recs, d_recs, sizeof(ltb_window_rec) * n, cudaMemcpyDeviceToHost));
#undef STEP
  if (e != cudaSuccess) rc = fail(LTB_ERROR, std::string("ltb_sss_work: ") + cudaGetErrorString(e));
  cudaFree(d_hf); cudaFree(d_sym); cudaFree(d_tag); cudaFree(d_rec); cudaFree(d_recs);
  return rc;
}

// ==========================================================================================
// kernel-level entry points
// ==========================================================================================
int ltb_kernel_pss_corr_host(int device, const ltb_cf *x, int n_streams, int64_t n, float *power) {
  if (!x || !power || n_streams <= 0 || n <= 0 || (n % 8) != 0) return fail(LTB_ERROR_INVALID_INPUTS, "n must be a positive multiple of 8");
  if (ltb_device_count() <= device || device < 0) return fail(LTB_ERROR, "no such CUDA device");
  LTB_CUDA(cudaSetDevice(device));
  int rc = ensure_constants(device);
  if (rc) return rc;
  const int cap = next_pow2(n + kCorrTile + 256);
  float2 *d_y = nullptr; float *d_p = nullptr;
  cudaError_t e = cudaMalloc(&d_y, sizeof(float2) * (size_t)n_streams * cap);
  if (e == cudaSuccess) e = cudaMalloc(&d_p, sizeof(float) * (size_t)n_streams * 3 * cap);
  if (e == cudaSuccess) e = cudaMemset(d_y, 0, sizeof(float2) * (size_t)n_streams * cap);
  if (e == cudaSuccess) e = cudaMemcpy2D(d_y, sizeof(float2) * (size_t)cap, x, sizeof(float2) * (size_t)n, sizeof(float2) * (size_t)n, n_streams, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    const int tps = (int)((n + kCorrTile - 1) / kCorrTile);
    const long long total = (long long)tps * n_streams;
    long long ctas = 2LL * (g_sm_count[device] > 0 ? g_sm_count[device] : 148);
    if (ctas > total) ctas = total;
    pss_corr_kernel<<<(unsigned)ctas, kCorrThreads>>>(d_y, d_p, 0, (int)n, (unsigned)(cap - 1), cap, tps, (int)total);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy2D(power, sizeof(float) * (size_t)n, d_p, sizeof(float) * (size_t)cap, sizeof(float) * (size_t)n, (size_t)n_streams * 3, cudaMemcpyDeviceToHost);
  cudaFree(d_y); cudaFree(d_p);
  if (e != cudaSuccess) return fail(LTB_ERROR, std::string("ltb_kernel_pss_corr_host: ") + cudaGetErrorString(e));
  return LTB_SUCCESS;
}

int ltb_kernel_pss_corr_fft_host(int device, const ltb_cf *x, int n_streams, int64_t n, float *power) {
  if (!x || !power || n_streams <= 0 || n < kOsStep) return fail(LTB_ERROR_INVALID_INPUTS, "n must hold at least one 896-sample block");
  if (ltb_device_count() <= device || device < 0) return fail(LTB_ERROR, "no such CUDA device");
  LTB_CUDA(cudaSetDevice(device));
  int rc = ensure_constants(device);
  if (!rc) rc = ensure_os_tables(device);
  if (rc) return rc;
  const int nblk = (int)(n / kOsStep);
  const int64_t n_out = (int64_t)nblk * kOsStep;
  const int cap = next_pow2(n + 1024);
  float2 *d_y = nullptr; float *d_p = nullptr;
  cudaError_t e = cudaMalloc(&d_y, sizeof(float2) * (size_t)n_streams * cap);
  if (e == cudaSuccess) e = cudaMalloc(&d_p, sizeof(float) * (size_t)n_streams * 3 * cap);
  if (e == cudaSuccess) e = cudaMemset(d_y, 0, sizeof(float2) * (size_t)n_streams * cap);
  if (e == cudaSuccess) e = cudaMemcpy2D(d_y, sizeof(float2) * (size_t)cap, x, sizeof(float2) * (size_t)n, sizeof(float2) * (size_t)n, n_streams, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = launch_corr_fft(device, d_y, d_p, 0, nblk, (unsigned)(cap - 1), cap, n_streams, 0);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy2D(power, sizeof(float) * (size_t)n_out, d_p, sizeof(float) * (size_t)cap, sizeof(float) * (size_t)n_out, (size_t)n_streams * 3, cudaMemcpyDeviceToHost);
  cudaFree(d_y); cudaFree(d_p);
  if (e != cudaSuccess) return fail(LTB_ERROR, std::string("ltb_kernel_pss_corr_fft_host: ") + cudaGetErrorString(e));
  return rc;
}

int ltb_kernel_decimate_host(int device, const void *x, int fmt, int n_streams, int64_t n_in, int decim, ltb_cf *y) {
  if (!x || !y || n_streams <= 0 || n_in <= 0 || !valid_decim(decim) || (n_in % decim) != 0 || !valid_format(fmt))
    return fail(LTB_ERROR_INVALID_INPUTS, "bad decimate arguments");
  if (ltb_device_count() <= device || device < 0) return fail(LTB_ERROR, "no such CUDA device");
  LTB_CUDA(cudaSetDevice(device));
  int rc = ensure_constants(device);
  if (rc) return rc;
  const int m = (int)(n_in / decim);
  const int cap = next_pow2(m + 8);
  const size_t in_row = (size_t)n_in * fmt_bytes(fmt);
  const size_t dev_row = (in_row + 127) / 128 * 128;   // the streaming kernels' bulk copies need 16-byte aligned rows
  void *d_in = nullptr; float2 *d_y = nullptr, *d_t0 = nullptr, *d_t1 = nullptr;
  float *d_bt = nullptr;
  rc = make_branch_taps_device(decim, &d_bt);
  if (rc) return rc;
  cudaError_t e = cudaMalloc(&d_in, dev_row * n_streams);
  if (e == cudaSuccess) e = cudaMalloc(&d_y, sizeof(float2) * (size_t)n_streams * cap);
  if (e == cudaSuccess) e = cudaMalloc(&d_t0, sizeof(float2) * (size_t)n_streams * kTailCap);
  if (e == cudaSuccess) e = cudaMalloc(&d_t1, sizeof(float2) * (size_t)n_streams * kTailCap);
  if (e == cudaSuccess) e = cudaMemset(d_t0, 0, sizeof(float2) * (size_t)n_streams * kTailCap);
  if (e == cudaSuccess) e = cudaMemcpy2D(d_in, dev_row, x, in_row, in_row, n_streams, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    int launches = 0;
    rc = launch_frontend_fmt(fmt, decim, d_in, (long long)dev_row, n_streams, m, d_t0, d_t1, d_bt, d_y, 0, (unsigned)(cap - 1), cap, 0, &launches);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy2D(y, sizeof(float2) * (size_t)m, d_y, sizeof(float2) * (size_t)cap, sizeof(float2) * (size_t)m, n_streams, cudaMemcpyDeviceToHost);
  cudaFree(d_in); cudaFree(d_y); cudaFree(d_t0); cudaFree(d_t1); cudaFree(d_bt);
  if (e != cudaSuccess) return fail(LTB_ERROR, std::string("ltb_kernel_decimate_host: ") + cudaGetErrorString(e));
  return rc;
}

int ltb_kernel_decimate_tc_host(int device, const void *x, int fmt, int decim, float full_scale, int n_streams, int64_t n_in,
                                int64_t chunk, ltb_cf *y) {
  if (!x || !y || n_streams <= 0 || !valid_format(fmt) || !tc_supported(fmt, decim) || n_in <= 0 || (n_in % (8 * decim)) != 0 ||
      chunk <= 0 || (chunk % (8 * decim)) != 0 || (fmt == LTB_FMT_FC32 && !(full_scale > 0.f && full_scale < 1e30f)))
    return fail(LTB_ERROR_INVALID_INPUTS, "a supported (format, rate) pair, n_in and chunk positive multiples of 8 decim; fc32 needs full_scale > 0");
  if (ltb_device_count() <= device || device < 0) return fail(LTB_ERROR, "no such CUDA device");
  LTB_CUDA(cudaSetDevice(device));
  int rc = ensure_constants(device);
  if (!rc) rc = ensure_tc_tables(device, fmt, decim);
  if (rc) return rc;
  const int bps = tc_sample_bytes(fmt);
  const int m = (int)(n_in / decim);
  const int cap = next_pow2(m + 8);
  const size_t in_row = (size_t)n_in * bps, dev_row = (in_row + 127) / 128 * 128;
  void *d_in = nullptr; float2 *d_y = nullptr; void *d_t[2] = {nullptr, nullptr}; int *d_err = nullptr;
  cudaError_t e = cudaMalloc(&d_in, dev_row * n_streams);
  if (e == cudaSuccess) e = cudaMalloc(&d_y, sizeof(float2) * (size_t)n_streams * cap);
  if (e == cudaSuccess) e = cudaMalloc(&d_t[0], 8 * (size_t)n_streams * kTcTailSamples);
  if (e == cudaSuccess) e = cudaMalloc(&d_t[1], 8 * (size_t)n_streams * kTcTailSamples);
  if (e == cudaSuccess) e = cudaMalloc(&d_err, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(d_err, 0, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(d_t[0], 0, 8 * (size_t)n_streams * kTcTailSamples);
  if (e == cudaSuccess) e = cudaMemcpy2D(d_in, dev_row, x, in_row, in_row, n_streams, cudaMemcpyHostToDevice);
  int cur = 0;
  for (int64_t c0 = 0; c0 < n_in && e == cudaSuccess && !rc; c0 += chunk) {
    const int nc = (int)(n_in - c0 < chunk ? n_in - c0 : chunk);
    int launches = 0;
    rc = launch_frontend_tc(device, fmt, decim, full_scale, (const char *)d_in + c0 * bps, (long long)dev_row, n_streams, nc, d_t[cur], d_t[cur ^ 1], d_y,
                            c0 / decim, (unsigned)(cap - 1), cap, d_err, 0, &launches);
    cur ^= 1;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy2D(y, sizeof(float2) * (size_t)m, d_y, sizeof(float2) * (size_t)cap, sizeof(float2) * (size_t)m, n_streams, cudaMemcpyDeviceToHost);
  cudaFree(d_in); cudaFree(d_y); cudaFree(d_t[0]); cudaFree(d_t[1]); cudaFree(d_err);
  if (e != cudaSuccess) return fail(LTB_ERROR, std::string("ltb_kernel_decimate_tc_host: ") + cudaGetErrorString(e));
  return rc;
}

// ==========================================================================================
// tables (host only)
// ==========================================================================================
int ltb_table_pss_taps(int n_id_2, float h_re[128], float h_im[128]) {
  if (n_id_2 < 0 || n_id_2 > 2 || !h_re || !h_im) return LTB_ERROR_INVALID_INPUTS;
  PssTaps t;
  make_pss_taps(n_id_2, t);
  std::memcpy(h_re, t.re, sizeof t.re);
  std::memcpy(h_im, t.im, sizeof t.im);
  return LTB_SUCCESS;
}

int ltb_table_decim_taps(int decim, float *taps, int max_taps) {
  if (!taps || decim < 1) return LTB_ERROR_INVALID_INPUTS;
  std::vector<float> v = make_decim_taps(decim);
  if ((int)v.size() > max_taps) return LTB_ERROR_INVALID_INPUTS;
  std::memcpy(taps, v.data(), v.size() * sizeof(float));
  return (int)v.size();
}

int ltb_table_sss(int n_id_2, int32_t c0[31], int32_t c1[31], int32_t s_tilde[31], int32_t z_tilde[31], int32_t n_id_1_table[900]) {
  if (n_id_2 < 0 || n_id_2 > 2) return LTB_ERROR_INVALID_INPUTS;
  SssTables st;
  make_sss_tables(n_id_2, st);
  std::memcpy(c0, st.c0, sizeof st.c0); std::memcpy(c1, st.c1, sizeof st.c1);
  std::memcpy(s_tilde, st.s_tilde, sizeof st.s_tilde); std::memcpy(z_tilde, st.z_tilde, sizeof st.z_tilde);
  std::memcpy(n_id_1_table, st.n_id_1, sizeof st.n_id_1);
  return LTB_SUCCESS;
}

int ltb_table_fft1024_twiddles(float w_re[1024], float w_im[1024]) {
  if (!w_re || !w_im) return LTB_ERROR_INVALID_INPUTS;
  make_fft1024_twiddles(w_re, w_im);
  return LTB_SUCCESS;
}

int ltb_table_os_filter(int n_id_2, float H_re[1024], float H_im[1024]) {
  if (n_id_2 < 0 || n_id_2 > 2 || !H_re || !H_im) return LTB_ERROR_INVALID_INPUTS;
  make_os_filter(n_id_2, H_re, H_im);
  return LTB_SUCCESS;
}

int ltb_table_tc_btab(int fmt, int decim, int8_t tab[208 * 128], int64_t *sum_t) {
  if (!valid_format(fmt) || !tab || !tc_supported(fmt, decim)) return LTB_ERROR_INVALID_INPUTS;
  long long sum = 0;
  make_tc_taps(decim, &sum);
  const std::vector<int8_t> t = make_tc_btab(fmt, decim);
  if (t.size() != 208 * 128) return LTB_ERROR;
  std::memcpy(tab, t.data(), t.size());
  if (sum_t) *sum_t = sum;
  return LTB_SUCCESS;
}

int ltb_table_cexp(float tab_re[4097], float tab_im[4097]) {
  if (!tab_re || !tab_im) return LTB_ERROR_INVALID_INPUTS;
  make_cexp_table(tab_re, tab_im);
  return LTB_SUCCESS;
}

int ltb_table_fft128_twiddles(float w_re[64], float w_im[64]) {
  if (!w_re || !w_im) return LTB_ERROR_INVALID_INPUTS;
  make_fft128_twiddles(w_re, w_im);
  return LTB_SUCCESS;
}

}  // extern "C"
