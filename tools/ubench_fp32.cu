// FP32-pipe micro-benchmark for B200 (sm_100a): measures the sustained FFMA /
// FFMA2 issue rate that bounds the PSS correlator and the decimator.  The result
// is the measured denominator for the "executed-flop" roofline fraction
// ("the builder must measure it with an FFMA micro-benchmark").
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ubench_fp32 ubench_fp32.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__constant__ float2 c_coef[64];

constexpr int NCH = 16;   // independent accumulator chains per thread

// mode 0: scalar FFMA, register operands
__global__ void __launch_bounds__(256) k_ffma_reg(float* out, int iters, float a, float b) {
  float acc[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < NCH; i++) acc[i] = fmaf(acc[i], a, b);
    }
  }
  float s = 0; 
#pragma unroll
  for (int i = 0; i < NCH; i++) s += acc[i];
  if (s == 12345.678f) out[0] = s;
}

// mode 1: scalar FFMA, multiplier from the constant bank (the correlator's form)
__global__ void __launch_bounds__(256) k_ffma_const(float* out, int iters, float x0) {
  float acc[NCH];
  float x[4];
#pragma unroll
  for (int i = 0; i < 4; i++) x[i] = x0 + threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < NCH; i++) acc[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < NCH; i++) acc[i] = fmaf(c_coef[(u * NCH + i) & 63].x, x[i & 3], acc[i]);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s += acc[i];
  if (s == 12345.678f) out[0] = s;
}

// mode 2: packed FFMA2, register operands
__global__ void __launch_bounds__(256) k_ffma2_reg(float* out, int iters, float a, float b) {
  float2 acc[NCH];
  float2 x[4];
#pragma unroll
  for (int i = 0; i < 4; i++) x[i] = make_float2(a + threadIdx.x * 1e-3f + i, b + i);
#pragma unroll
  for (int i = 0; i < NCH; i++) acc[i] = make_float2(i, -i);
  float2 cc[4] = {make_float2(a, b), make_float2(b, a), make_float2(a * 0.5f, b), make_float2(b, a * 0.5f)};
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < NCH; i++) acc[i] = __ffma2_rn(cc[(u + i) & 3], x[i & 3], acc[i]);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s += acc[i].x + acc[i].y;
  if (s == 12345.678f) out[0] = s;
}

// mode 3: packed FFMA2, multiplier pair from the constant bank
__global__ void __launch_bounds__(256) k_ffma2_const(float* out, int iters, float x0) {
  float2 acc[NCH];
  float2 x[4];
#pragma unroll
  for (int i = 0; i < 4; i++) x[i] = make_float2(x0 + threadIdx.x * 1e-3f + i, x0 - i);
#pragma unroll
  for (int i = 0; i < NCH; i++) acc[i] = make_float2(i, -i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < NCH; i++) acc[i] = __ffma2_rn(c_coef[(u * NCH + i) & 63], x[i & 3], acc[i]);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s += acc[i].x + acc[i].y;
  if (s == 12345.678f) out[0] = s;
}

// mode 4: the correlator's inner-loop mix: per folded tap 1 FADD2 + 4 FFMA2 (const coef), 4 outputs/thread
__global__ void __launch_bounds__(256) k_mix(float* out, int iters, float x0) {
  float2 acc[16];
  float2 xa[4], xb[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { xa[i] = make_float2(x0 + threadIdx.x * 1e-3f + i, x0 - i); xb[i] = make_float2(x0 * i, x0 + i); }
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = make_float2(i, -i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int o = 0; o < 4; o++) {
        float2 s = __fadd2_rn(xa[o], xb[(o + u) & 3]);
        acc[4 * o + 0] = __ffma2_rn(c_coef[4 * u + 0], s, acc[4 * o + 0]);
        acc[4 * o + 1] = __ffma2_rn(c_coef[4 * u + 1], s, acc[4 * o + 1]);
        acc[4 * o + 2] = __ffma2_rn(c_coef[4 * u + 2], s, acc[4 * o + 2]);
        acc[4 * o + 3] = __ffma2_rn(c_coef[4 * u + 3], s, acc[4 * o + 3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) xa[i].x += 1e-9f;
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i].x + acc[i].y;
  if (s == 12345.678f) out[0] = s;
}

// mode 5: same mix with scalar FFMA/FADD (2 FADD + 8 FFMA per tap)
__global__ void __launch_bounds__(256) k_mix_scalar(float* out, int iters, float x0) {
  float acc[32];
  float2 xa[4], xb[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { xa[i] = make_float2(x0 + threadIdx.x * 1e-3f + i, x0 - i); xb[i] = make_float2(x0 * i, x0 + i); }
#pragma unroll
  for (int i = 0; i < 32; i++) acc[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int o = 0; o < 4; o++) {
        float sr = xa[o].x + xb[(o + u) & 3].x, si = xa[o].y + xb[(o + u) & 3].y;
        acc[8 * o + 0] = fmaf(c_coef[4 * u + 0].x, sr, acc[8 * o + 0]);
        acc[8 * o + 1] = fmaf(c_coef[4 * u + 0].y, si, acc[8 * o + 1]);
        acc[8 * o + 2] = fmaf(c_coef[4 * u + 1].x, sr, acc[8 * o + 2]);
        acc[8 * o + 3] = fmaf(c_coef[4 * u + 1].y, si, acc[8 * o + 3]);
        acc[8 * o + 4] = fmaf(c_coef[4 * u + 2].x, sr, acc[8 * o + 4]);
        acc[8 * o + 5] = fmaf(c_coef[4 * u + 2].y, si, acc[8 * o + 5]);
        acc[8 * o + 6] = fmaf(c_coef[4 * u + 3].x, sr, acc[8 * o + 6]);
        acc[8 * o + 7] = fmaf(c_coef[4 * u + 3].y, si, acc[8 * o + 7]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) xa[i].x += 1e-9f;
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) s += acc[i];
  if (s == 12345.678f) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; i++) launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  float2 h[64]; for (int i = 0; i < 64; i++) h[i] = make_float2(1.0f / (i + 3), -1.0f / (i + 5));
  CK(cudaMemcpyToSymbol(c_coef, h, sizeof(h)));
  float* d; CK(cudaMalloc(&d, 4));
  const int iters = 4096;
  for (int bps = 1; bps <= 4; bps *= 2) {
    int grid = sms * bps * 4, thr = 256;
    double lanes = (double)grid * thr;
    struct { const char* name; double fma_per_thread; double ms; } r[6];
    r[0] = {"ffma_reg", (double)iters * 8 * NCH, time_ms([&] { k_ffma_reg<<<grid, thr>>>(d, iters, 1.0001f, 1e-5f); }, 5)};
    r[1] = {"ffma_const", (double)iters * 8 * NCH, time_ms([&] { k_ffma_const<<<grid, thr>>>(d, iters, 0.5f); }, 5)};
    r[2] = {"ffma2_reg", (double)iters * 8 * NCH * 2, time_ms([&] { k_ffma2_reg<<<grid, thr>>>(d, iters, 1.0001f, 1e-5f); }, 5)};
    r[3] = {"ffma2_const", (double)iters * 8 * NCH * 2, time_ms([&] { k_ffma2_const<<<grid, thr>>>(d, iters, 0.5f); }, 5)};
    // mix: per (u,o): 2 add-lanes + 8 fma-lanes = 10 FP32 lane-ops
    r[4] = {"mix_ffma2", (double)iters * 8 * 4 * 10, time_ms([&] { k_mix<<<grid, thr>>>(d, iters, 0.5f); }, 5)};
    r[5] = {"mix_scalar", (double)iters * 8 * 4 * 10, time_ms([&] { k_mix_scalar<<<grid, thr>>>(d, iters, 0.5f); }, 5)};
    for (int i = 0; i < 6; i++) {
      double ops = r[i].fma_per_thread * lanes;  // FP32 lane-ops (an FMA counts 1 op = 2 flop)
      printf("{\"bench\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"Tlaneops_per_s\": %.3f, \"TFLOPs_if_all_fma\": %.3f, \"laneops_per_clk_per_sm_at_max\": %.2f}\n",
             r[i].name, bps * 4, r[i].ms, ops / r[i].ms * 1e-9, 2 * ops / r[i].ms * 1e-9,
             ops / (r[i].ms * 1e-3) / sms / (p.clockRate * 1e3));
    }
  }
  // long sustained run (≈2 s) to see the clock under load
  {
    int grid = sms * 8, thr = 256; int it2 = iters * 16;
    double ms = time_ms([&] { k_ffma2_const<<<grid, thr>>>(d, it2, 0.5f); }, 20);
    double ops = (double)it2 * 8 * NCH * 2 * grid * thr;
    printf("{\"bench\": \"ffma2_const_sustained\", \"ms\": %.4f, \"Tlaneops_per_s\": %.3f, \"TFLOPs\": %.3f}\n", ms, ops / ms * 1e-9, 2 * ops / ms * 1e-9);
    ms = time_ms([&] { k_ffma_const<<<grid, thr>>>(d, it2, 0.5f); }, 20);
    ops = (double)it2 * 8 * NCH * grid * thr;
    printf("{\"bench\": \"ffma_const_sustained\", \"ms\": %.4f, \"Tlaneops_per_s\": %.3f, \"TFLOPs\": %.3f}\n", ms, ops / ms * 1e-9, 2 * ops / ms * 1e-9);
  }
  return 0;
}
