// Stand-alone driver of the integer tensor-core front end (gr-ltetrigger_b200/csrc/ltb_tc_frontend.cuh):
// random sc16 streams through decimate_tc_kernel<G>, every checked output compared EXACTLY with an int64
// evaluation on the host, then timed.  One variant per process (a watchdog trap poisons the context):
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/ubench_tc_i8 tools/ubench_tc_i8.cu
//   ./tools/ubench_tc_i8 <fmt = 0 (fc32 as 23-bit fixed point) | 1 (sc16) | 2 (sc8)> [n_streams] [n_in per stream] [chunks] [0] [decim = 16 | 12 | 8 | 4 | 2]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef UBENCH_NO_DBG         // -DUBENCH_NO_DBG: time the kernel as the library builds it (the accumulator check then reports mismatches)
#define LTB_TC_DBG_ACC 1      // this tool reads back the raw accumulators of tile 0
#endif
#include "../gr-ltetrigger_b200/csrc/ltb_tc_frontend.cuh"

using namespace ltb;

static double izero(double v) {
  double sum = 1, u = 1, h = v / 2; int n = 1;
  do { double t = h / n; n++; t *= t; u *= t; sum += u; } while (u >= 1e-21 * sum);
  return sum;
}
// the float32 taps of rational_resampler_ccc(1, D) (same design as ltb_tables.cpp)
static std::vector<float> taps_for(int D) {
  const int ntaps = tc_ntaps(D), M = (ntaps - 1) / 2;
  const double beta = 7.0, tw = 0.1 / D, mid = 0.5 / D - tw / 2, fw = 2 * M_PI * mid;
  std::vector<float> w(ntaps), t(ntaps);
  for (int i = 0; i < ntaps; ++i) { const double x = 2.0 * i / (ntaps - 1) - 1; w[i] = (float)(izero(beta * sqrt(1 - x * x)) / izero(beta)); }
  for (int n = -M; n <= M; ++n) t[n + M] = (float)((n == 0 ? fw / M_PI : sin(n * fw) / (n * M_PI)) * w[n + M]);
  double g = t[M];
  for (int n = 1; n <= M; ++n) g += 2 * t[n + M];
  for (auto &v : t) v = (float)(v / g);
  return t;
}

static int sw_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }

// tap table in the kernel's shared-memory image (see ltb_tc_frontend.cuh, ltb_tables.cpp make_tc_btab)
static int digit(int t, int v) {              // balanced base-256 digits v = 0..2
  int d0 = ((t + 128) & 255) - 128; int t1 = (t - d0) >> 8;
  int d1 = ((t1 + 128) & 255) - 128; int t2 = (t1 - d1) >> 8;
  if (t2 < -128 || t2 > 127) { fprintf(stderr, "tap does not fit three digits\n"); exit(2); }
  return v == 0 ? d0 : v == 1 ? d1 : v == 2 ? t2 : 0;
}
static int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }
static std::vector<int8_t> make_btab(int fmt, int D, const std::vector<int> &T) {
  std::vector<int8_t> tab((size_t)kTcBTileBytes, 0);
  auto tapq = [&](int j) { return (j >= 0 && j < (int)T.size()) ? T[j] : 0; };
  const int bpc = fmt == 0 ? 4 : fmt == 1 ? 2 : 1, spk = 32 / bpc, g = gcd_i(spk, D), nph = D / g;
  for (int n = 0; n < kTcBRows; ++n) {
    const int d = n / 4, v = n % 4;
    for (int i = 0; i < nph; ++i)
      for (int p = 0; p < spk; ++p) {
        const int t = tapq(D * d - i * g - p);
        if (!t) continue;
        for (int byte = 0; byte < (fmt == 0 ? 3 : bpc); ++byte) {
          const int dg = fmt == 0 ? v + 1 - byte : v - byte;
          if (dg < 0 || dg > 2) continue;
          const int kb = 32 * i + bpc * p + byte;
          tab[sw_off(n, kb >> 4) + (kb & 15)] = (int8_t)digit(t, dg);
        }
      }
  }
  return tab;
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int g_ldg = 0;                                   // (the load-from-global variant was measured and removed: DESIGN.md section 10)
template <int FMT, int D>
static void launch(const CUtensorMap &map, const TcParams &P, int grid, int S, int n_chunk, void *t_old, void *t_new) {
  cudaFuncSetAttribute(decimate_tc_kernel<FMT, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes());
  decimate_tc_kernel<FMT, D><<<grid, kTcThreads, tc_smem_bytes()>>>(map, P);
  tc_tail_kernel<FMT, D><<<S, 256>>>(P.in, P.stride_bytes, n_chunk, t_old, t_new);
}
#define UB_VARIANTS(X) X(0, 2) X(0, 4) X(0, 8) X(0, 12) X(0, 16) X(0, 24) X(0, 32) X(1, 4) X(1, 8) X(1, 12) X(1, 16) X(1, 24) X(1, 32) X(2, 8) X(2, 16) X(2, 24) X(2, 32)

int main(int argc, char **argv) {
  const int G = argc > 1 ? atoi(argv[1]) : 1;            // input format: 0 fc32, 1 sc16, 2 sc8
  const int bps = G == 0 ? 8 : G == 1 ? 4 : 2;
  const float FS = 2.0f;                                 // fc32: declared range
  const float q_inv = (float)(0.5 / (double)FS);
  const int S = argc > 2 ? atoi(argv[2]) : 64;
  const int n_in = argc > 3 ? atoi(argv[3]) : 3072000;
  const int chunks = argc > 4 ? atoi(argv[4]) : 1;       // > 1: feed the stream in `chunks` calls (tail carried)
  (void)g_ldg;
  const int DEC = argc > 6 ? atoi(argv[6]) : 16;         // decimation
  if (!tc_supported(G, DEC)) { fprintf(stderr, "unsupported (fmt, decim)\n"); return 2; }
  const int ROWS = 16 * DEC, NTAPS = tc_ntaps(DEC), TAILS = kTcHalo * ROWS, SHIFT = tc_tap_shift(DEC);
  if (G < 0 || G > 2) { fprintf(stderr, "fmt must be 0 (fc32), 1 (sc16) or 2 (sc8)\n"); return 2; }
  const std::vector<float> tf = taps_for(DEC);
  std::vector<int> T(tf.size());
  long long sumT = 0;
  for (size_t j = 0; j < tf.size(); ++j) { T[j] = (int)llrint(ldexp((double)tf[j], SHIFT)); sumT += T[j]; }
  const std::vector<int8_t> btab = make_btab(G, DEC, T);
  const float out_scale = G == 0 ? (float)((double)FS / 4194303.0 * ldexp(1.0, 8 - SHIFT)) : (float)ldexp(1.0, -(SHIFT + (G == 1 ? 15 : 7)));

  const size_t row_bytes = (size_t)n_in * bps;
  std::vector<short> x((size_t)S * n_in * 2);        // sample values (sc8: within -128..127), packed below
  srand(12345);
  for (auto &v : x) v = G != 2 ? (short)((rand() & 0xffff) - 32768) : (short)((rand() & 0xff) - 128);
  // a few structured streams: extremes exercise the digit bounds
  for (int i = 0; i < n_in * 2 && S > 2; ++i) { x[(size_t)1 * n_in * 2 + i] = G != 2 ? 32767 : 127; x[(size_t)2 * n_in * 2 + i] = G != 2 ? -32768 : -128; }
  std::vector<signed char> x8;
  if (G == 2) { x8.resize(x.size()); for (size_t i = 0; i < x.size(); ++i) x8[i] = (signed char)x[i]; }
  // fc32: floats over +-1.25 x the declared range (some saturate); mq = the 23-bit fixed-point value the kernel must form
  std::vector<float> xf; std::vector<int> mq;
  auto quant = [&](float v) {
    float u = fmaf(v, q_inv, 0.5f); if (!(u > 0.f)) u = 0.f; if (u > 1.f) u = 1.f;
    const float t = fmaf(u, kTcQMul, kTcQAdd); unsigned b; memcpy(&b, &t, 4); return (int)(b & 0x7fffffu);
  };
  if (G == 0) {
    xf.resize(x.size()); mq.resize(x.size());
    for (size_t i = 0; i < x.size(); ++i) { xf[i] = (float)x[i] * (1.25f * FS / 32768.0f) * (1.0f + 1e-3f * (float)(rand() & 1023)); mq[i] = quant(xf[i]); }
  }

  void *d_x; void *d_tail[2]; float2 *d_y; int8_t *d_b; int *d_err; int *d_acc;
  cudaMalloc(&d_acc, 128 * kTcBRows * 4); cudaMemset(d_acc, 0x7f, 128 * kTcBRows * 4);
  const int m_total = n_in / DEC;
  int cap = 1; while (cap < m_total + 64) cap <<= 1;
  cudaMalloc(&d_x, (size_t)S * row_bytes);
  cudaMalloc(&d_tail[0], (size_t)S * kTcTailSamples * 8); cudaMalloc(&d_tail[1], (size_t)S * kTcTailSamples * 8); (void)TAILS;
  cudaMemset(d_tail[0], 0, (size_t)S * kTcTailSamples * 8);
  cudaMalloc(&d_y, (size_t)S * cap * 8); cudaMemset(d_y, 0xff, (size_t)S * cap * 8);
  cudaMalloc(&d_b, btab.size()); cudaMalloc(&d_err, 4); cudaMemset(d_err, 0, 4);
  cudaMemcpy(d_x, G == 0 ? (const void *)xf.data() : G == 1 ? (const void *)x.data() : (const void *)x8.data(), (size_t)S * row_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(d_b, btab.data(), btab.size(), cudaMemcpyHostToDevice);

  EncodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
    printf("{\"error\": \"cuTensorMapEncodeTiled not found\"}\n"); return 1;
  }
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);

  auto run_chunk = [&](int c0, int n_chunk, int tail_cur) -> int {
    CUtensorMap map;
    const int full_rows = n_chunk / ROWS;
    const cuuint64_t gdim[3] = {(cuuint64_t)ROWS * bps, (cuuint64_t)(full_rows > 0 ? full_rows : 1), (cuuint64_t)S};
    const cuuint64_t gstr[2] = {(cuuint64_t)ROWS * bps, (cuuint64_t)row_bytes};
    const cuuint32_t box[3] = {256, (cuuint32_t)kTcTileRows, 1}, estr[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (char *)d_x + (size_t)c0 * bps, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("{\"error\": \"cuTensorMapEncodeTiled -> %d\"}\n", (int)r); return 1; }
    TcParams P;
    P.in = (char *)d_x + (size_t)c0 * bps; P.stride_bytes = (long long)row_bytes; P.n_in = n_chunk; P.n_streams = S;
    P.tail = d_tail[tail_cur]; P.y_ring = d_y; P.n_base = c0 / DEC; P.m_out = n_chunk / DEC; P.cap_mask = (unsigned)(cap - 1); P.cap = cap;
    const int rows = (n_chunk + ROWS - 1) / ROWS;
    P.tiles_per_stream = (rows + kTcUseful - 1) / kTcUseful; P.total_tiles = P.tiles_per_stream * S;
    P.btab = d_b; P.c_const = G == 1 ? 128 * sumT : G == 0 ? -16384 * sumT : 0; P.err = d_err; P.dbg_acc = c0 == 0 ? d_acc : nullptr;
    P.q_inv = q_inv; P.out_scale = out_scale;
    const int grid = P.total_tiles < sms ? P.total_tiles : sms;
#define X(F, DD) if (G == F && DEC == DD) launch<F, DD>(map, P, grid, S, n_chunk, d_tail[tail_cur], d_tail[tail_cur ^ 1]);
    UB_VARIANTS(X)
#undef X
    return 0;
  };
  auto run_all = [&]() -> int {
    cudaMemset(d_tail[0], 0, (size_t)S * kTcTailSamples * 8);
    int tc = 0, c0 = 0;
    for (int c = 0; c < chunks; ++c) {
      int n_chunk = (c == chunks - 1) ? n_in - c0 : (n_in / chunks) / (8 * DEC) * (8 * DEC);
      if (c < chunks - 1 && (c & 1)) n_chunk += 8 * DEC;                 // ragged: not always a multiple of a row
      if (run_chunk(c0, n_chunk, tc)) return 1;
      c0 += n_chunk; tc ^= 1;
    }
    return 0;
  };
  if (run_all()) return 1;
  cudaError_t e = cudaDeviceSynchronize();
  int herr = 0; cudaMemcpy(&herr, d_err, 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("{\"fmt\": %d, \"error\": \"%s\", \"watchdog\": %d}\n", G, cudaGetErrorString(e), herr); return 1; }

  // ---- raw accumulators of tile 0 (stream 0, rows -3..60; lanes 0..63 re, 64..127 im) against the host ----
  long long acc_bad = 0; int acc_first[4] = {-1, -1, 0, 0};
  {
    std::vector<int> acc(128 * kTcBRows);
    cudaMemcpy(acc.data(), d_acc, acc.size() * 4, cudaMemcpyDeviceToHost);
    for (int lanei = 0; lanei < 128; ++lanei) {
      const int comp = lanei >> 6, row = (lanei & 63) - kTcHalo;
      for (int col = 0; col < 196; ++col) {
        const int u = col >> 2, v = col & 3;
        long long want = 0;
        for (int p = 0; p < ROWS; ++p) {
          const int j = DEC * u - p;
          if (j < 0 || j >= NTAPS) continue;
          const long long nidx = (long long)row * ROWS + p;
          const int xv = nidx >= 0 && nidx < n_in ? x[2 * nidx + comp] : 0;
          if (G == 0) {
            const int m = nidx >= 0 && nidx < n_in ? mq[2 * nidx + comp] : kTcQMid;
            for (int bi = 0; bi < 3; ++bi) { const int i = v + 1 - bi; if (i >= 0 && i <= 2) want += (long long)((m >> (8 * bi)) & 255) * digit(T[j], i); }
          } else if (G == 1) {
            const int lo = (xv & 255) - 128, hi = xv >> 8;
            want += (long long)lo * (v <= 2 ? digit(T[j], v) : 0) + (long long)hi * (v >= 1 ? digit(T[j], v - 1) : 0);
          } else {
            want += (long long)xv * (v <= 2 ? digit(T[j], v) : 0);
          }
        }
        const int got = acc[lanei * kTcBRows + col];
        if ((long long)got != want) { if (!acc_bad) { acc_first[0] = lanei; acc_first[1] = col; acc_first[2] = got; acc_first[3] = (int)want; } acc_bad++; }
      }
    }
  }
  // ---- exact check against int64 on the host ----
  std::vector<float2> y((size_t)S * cap);
  cudaMemcpy(y.data(), d_y, y.size() * 8, cudaMemcpyDeviceToHost);
  long long checked = 0, bad = 0; int first_bad_s = -1, first_bad_k = -1; float gb = 0, wb = 0;
  for (int s = 0; s < S; ++s) {
    const short *xs = x.data() + (size_t)s * n_in * 2;
    const int step = (s < 4) ? 1 : 97;                                   // four streams in full, the others sampled
    for (int k = 0; k < m_total; k += step) {
      long long are = 0, aim = 0;
      for (int j = 0; j < NTAPS; ++j) {
        const long long n = (long long)DEC * k - j;
        if (n < 0) break;
        if (G == 0) {
          const int mr = mq[(size_t)s * n_in * 2 + 2 * n], mi = mq[(size_t)s * n_in * 2 + 2 * n + 1];
          are += (long long)T[j] * (mr - kTcQMid) - (long long)(mr & 255) * digit(T[j], 0);
          aim += (long long)T[j] * (mi - kTcQMid) - (long long)(mi & 255) * digit(T[j], 0);
        } else { are += (long long)T[j] * xs[2 * n]; aim += (long long)T[j] * xs[2 * n + 1]; }
      }
      if (G == 0) { are /= 256; aim /= 256; }
      const float sc = out_scale;
      const float wre = (float)are * sc, wim = (float)aim * sc;
      const float2 g = y[(size_t)s * cap + k];
      checked++;
      if (memcmp(&g.x, &wre, 4) || memcmp(&g.y, &wim, 4)) {
        if (!bad) { first_bad_s = s; first_bad_k = k; gb = g.x; wb = wre; }
        bad++;
      }
    }
  }
  // ---- timing ----
  float ms = 0;
  if ((!bad || getenv("UBENCH_FORCE_TIME")) && chunks == 1) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    run_all();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) run_all();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  }
  e = cudaDeviceSynchronize();
  printf("{\"fmt\": %d, \"decim\": %d, \"acc_tile0_mismatches\": %lld, \"acc_first_bad\": [%d, %d, %d, %d], \"streams\": %d, \"n_in\": %d, \"chunks\": %d, \"checked\": %lld, \"mismatches\": %lld, \"first_bad\": [%d, %d, %g, %g], "
         "\"ms\": %.4f, \"input_Gsamples_per_s\": %.1f, \"GB_per_s\": %.1f, \"status\": \"%s\"}\n",
         G, DEC, acc_bad, acc_first[0], acc_first[1], acc_first[2], acc_first[3], S, n_in, chunks, checked, bad, first_bad_s, first_bad_k, gb, wb, ms, ms > 0 ? (double)S * n_in / ms / 1e6 : 0.0,
         ms > 0 ? (double)S * n_in * bps / ms / 1e6 : 0.0, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
  return bad ? 3 : 0;
}
