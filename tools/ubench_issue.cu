// Can other instructions issue in the shadow of FFMA2 (rt = 2 cycles/SMSP)?  Per 16 independent
// FFMA2 the loop adds K instructions of one kind (FSEL / IADD3 / SHFL / LDS.64) on independent
// registers and reports cycles per 16-FFMA2 group per SMSP (32 = FMA-pipe bound, no interference).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_issue ubench_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int K, int WSHARED>
__global__ void __launch_bounds__(128) k(float* out, int iters, float x0, int sel) {
  __shared__ float2 sm[1024];
  float2 acc[16], w[16];
  float e[8];
  int ia[8];
  for (int i = threadIdx.x; i < 1024; i += 128) sm[i] = make_float2(i, -i);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; i++) { acc[i] = make_float2(i, -i); w[i] = make_float2(x0 + i, x0 - i); }
#pragma unroll
  for (int i = 0; i < 8; i++) { e[i] = x0 * i; ia[i] = threadIdx.x + i; }
  const float c = x0 * 1e-3f;
  float cv[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { cv[i] = c * (i + 1); asm volatile("" : "+f"(cv[i])); }
  const bool up = (threadIdx.x & sel) != 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 4; rep++) {
#pragma unroll
      for (int o = 0; o < 16; o++) {
        // WSHARED: the 16 FFMA2 of a group share the data operand (reuse cache) and take different
        // scalar taps -> 3 register reads each instead of 4
        if (WSHARED) acc[o] = __ffma2_rn(make_float2(cv[o], cv[o]), w[rep], acc[o]);
        else acc[o] = __ffma2_rn(make_float2(c, c), w[o], acc[o]);
        if (o < K) {
          const int j = (o + rep) & 7;
          if (KIND == 1) e[j] = up ? e[j] : e[(j + 1) & 7];                       // FSEL
          if (KIND == 2) ia[j] = ia[j] + ia[(j + 3) & 7];                          // IADD3
          if (KIND == 3) e[j] = __shfl_xor_sync(0xffffffffu, e[j], 1 + (o & 7));   // SHFL
          if (KIND == 4) { float2 v = sm[(ia[j] + o * 32) & 1023]; e[j] += v.x; }   // LDS.64 (+FADD)
        }
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i].x + acc[i].y;
#pragma unroll
  for (int i = 0; i < 8; i++) s += e[i] + ia[i];
  if (s == 12345.678f) out[0] = s;
}

template <int KIND, int K, int WSHARED>
void run(const char* name, float* d, int sms, double clk_hz, int ctas_per_sm) {
  const int iters = 4000;
  k<KIND, K, WSHARED><<<sms * ctas_per_sm, 128>>>(d, 10, 0.5f, 8);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<KIND, K, WSHARED><<<sms * ctas_per_sm, 128>>>(d, iters, 0.5f, 8);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // per SMSP: ctas_per_sm warps, each iters*4 groups of 16 FFMA2
  double groups = (double)ctas_per_sm * iters * 4;
  printf("{\"wshared\": %d, \"kind\": \"%s\", \"extra_per_16_ffma2\": %d, \"warps_per_smsp\": %d, \"ms\": %.3f, \"cycles_per_group_per_smsp\": %.1f}\n",
         WSHARED, name, K, ctas_per_sm, ms, ms * 1e-3 * clk_hz / groups);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* d; cudaMalloc(&d, 4);
  const double clk = p.clockRate * 1e3;
  for (int w : {4}) {
#define ALLK(WS) \
    run<0, 0, WS>("none", d, p.multiProcessorCount, clk, w); \
    run<1, 8, WS>("fsel", d, p.multiProcessorCount, clk, w); \
    run<2, 4, WS>("iadd3", d, p.multiProcessorCount, clk, w); \
    run<2, 8, WS>("iadd3", d, p.multiProcessorCount, clk, w); \
    run<2, 16, WS>("iadd3", d, p.multiProcessorCount, clk, w); \
    run<3, 8, WS>("shfl", d, p.multiProcessorCount, clk, w); \
    run<4, 4, WS>("lds64", d, p.multiProcessorCount, clk, w); \
    run<4, 8, WS>("lds64", d, p.multiProcessorCount, clk, w);
    ALLK(0)
    ALLK(1)
  }
  return 0;
}
