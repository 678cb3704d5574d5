// LDGSTS (cp.async) issue-rate micro-benchmark: 4/8/16-byte async copies global(L2-resident)->shared,
// and their effect on concurrent LDS+FFMA2 work.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int BYTES>
__device__ __forceinline__ void cpa(void* s, const void* g) {
  unsigned d = (unsigned)__cvta_generic_to_shared(s);
  if (BYTES == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(g) : "memory");
  if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(g) : "memory");
  if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(g) : "memory");
}

// each CTA (128 threads) copies `iters` x 64 ops per thread; buffer small enough to stay in L2
template <int BYTES, bool SCATTER>
__global__ void __launch_bounds__(128) k_copy(const char* __restrict__ g, int iters, size_t span) {
  extern __shared__ __align__(16) char sm[];
  const int t = threadIdx.x;
  const char* src = g + ((size_t)blockIdx.x * 65536) % span;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 64; ++u) {
      // contiguous per warp in global; in shared either contiguous or scattered over 16 rows (stride 545*8+..)
      int idx = u * 128 + t;
      char* dst = SCATTER ? sm + ((idx & 15) * 545 + (idx >> 4)) * BYTES : sm + idx * BYTES;
      cpa<BYTES>(dst, src + (size_t)idx * BYTES);
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();
  }
}

int main() {
  char* g; size_t span = 64u << 20; CK(cudaMalloc(&g, span + (1 << 20))); CK(cudaMemset(g, 1, span + (1 << 20)));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 50, grid = 148 * 3;
#define RUN(B, S) { \
    size_t smem = (size_t)16 * 545 * B + 1024; if (smem < (size_t)64 * 128 * B) smem = (size_t)64 * 128 * B; \
    CK(cudaFuncSetAttribute(k_copy<B, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_copy<B, S><<<grid, 128, smem>>>(g, 2, span); CK(cudaDeviceSynchronize()); \
    cudaEventRecord(e0); k_copy<B, S><<<grid, 128, smem>>>(g, iters, span); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); \
    float ms; cudaEventElapsedTime(&ms, e0, e1); \
    double ops = (double)grid * iters * 64 * 4;  /* warp-level LDGSTS */ \
    double bytes = ops * 32 * B; \
    printf("{\"bytes_per_lane\": %d, \"scatter\": %d, \"ms\": %.3f, \"TBps\": %.3f, \"cycles_per_warp_op_per_sm\": %.2f}\n", B, (int)S, ms, bytes / ms * 1e-9, ms * 1e-3 * 1.965e9 / (ops / 148)); }
  RUN(4, false) RUN(8, false) RUN(16, false) RUN(4, true) RUN(8, true) RUN(16, true)
  return 0;
}
