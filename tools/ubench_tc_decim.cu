// Prototype, not product: the D = 16 decimator as a tensor-core GEMM (tcgen05, kind::tf32, 3xTF32),
// to measure what DESIGN.md section 10 only estimates.  Nothing in the library uses this.
//
//   Z[b][q] = sum_p X[b][p] * T[q][p]          X: input blocks of 16 complex samples (K = 32 interleaved
//   y[k]    = sum_q Z[k - q][q],  q = 0..32       re/im), T[q][p] = taps[16 q + 15 - p] (N = 66 -> 80 columns)
//
// i.e. y[k] = sum_j taps[j] x[16 k + 15 - j]: the polyphase filter with blocks aligned so that every position
// uses the same block offset.  One CTA = 128 threads; a tile is 128 consecutive blocks (M = 128) and yields the
// 96 outputs whose 33 blocks of history lie inside it.  A is split x = hi + lo (hi = TF32 truncation, exact),
// B likewise; D += hi*hi + lo*hi + hi*lo.  Checked against a double-precision evaluation on the host.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tc_decim tools/ubench_tc_decim.cu
//   ./tools/ubench_tc_decim [n_tiles]
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kM = 128, kK = 32, kN = 80, kQ = 33, kOutPerTile = kM - 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 in [0,14),
// LBO >> 4 in [16,30) (unused for swizzled K-major), SBO >> 4 in [32,46) = 1024 B between 8-row groups,
// version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// cute::UMMA::InstrDescriptor for kind::tf32, F32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
}

// element (row r, 16-byte chunk c) of a [rows][128 B] K-major tile in 128-byte swizzle
__device__ __forceinline__ int sw_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }

__global__ void __launch_bounds__(128, 2)
tc_decim_kernel(const float *__restrict__ x /* [n_blocks][32] */, const float *__restrict__ tmat /* [80][32] */,
                float2 *__restrict__ y, int n_tiles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *a_hi = smem, *a_lo = smem + 16384, *b_hi = smem + 32768, *b_lo = smem + 32768 + 10240;
  float *zs = reinterpret_cast<float *>(smem + 32768 + 20480);       // [128][81]
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;

  // B = T split into TF32 hi / lo, written once in the swizzled layout
  for (int i = tid; i < kN * 8; i += 128) {
    const int n = i >> 3, c = i & 7;
    float4 v = *reinterpret_cast<const float4 *>(tmat + n * kK + c * 4), h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
    *reinterpret_cast<float4 *>(b_hi + sw_off(n, c)) = h;
    *reinterpret_cast<float4 *>(b_lo + sw_off(n, c)) = l;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc();
  uint32_t parity = 0;

  float4 nxt[8];                                                       // the next tile's row, loaded one tile ahead
  if (blockIdx.x < n_tiles) {
    const float *row = x + ((size_t)blockIdx.x * kOutPerTile + tid) * kK;
#pragma unroll
    for (int c = 0; c < 8; ++c) nxt[c] = *reinterpret_cast<const float4 *>(row + c * 4);
  }
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- A: row tid of the tile, split and stored swizzled ----------------------------------------------
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4 v = nxt[c], h, l;
      h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
      h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
      h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
      h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
      *reinterpret_cast<float4 *>(a_hi + sw_off(tid, c)) = h;
      *reinterpret_cast<float4 *>(a_lo + sw_off(tid, c)) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> tensor core reads
    __syncthreads();
    if (tile + (int)gridDim.x < n_tiles) {
      const float *row = x + ((size_t)(tile + gridDim.x) * kOutPerTile + tid) * kK;
#pragma unroll
      for (int c = 0; c < 8; ++c) nxt[c] = *reinterpret_cast<const float4 *>(row + c * 4);
    }
    // ---- 12 MMAs by one thread ------------------------------------------------------------------------------
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t acc = 0;
#pragma unroll
      for (int k = 0; k < kK / 8; ++k) {
        const uint64_t ah = make_desc(smem_u32(a_hi) + k * 32), al = make_desc(smem_u32(a_lo) + k * 32);
        const uint64_t bh = make_desc(smem_u32(b_hi) + k * 32), bl = make_desc(smem_u32(b_lo) + k * 32);
        const uint64_t aa[3] = {ah, al, ah}, bb[3] = {bh, bh, bl};
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
              "l"(aa[t]), "l"(bb[t]), "r"(idesc), "r"(acc)
              : "memory");
          acc = 1;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // ---- wait for the accumulator, TMEM -> registers -> shared memory ---------------------------------------
    {
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(&mbar)), "r"(parity)
            : "memory");
      }
      parity ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float *zrow = zs + tid * 81;
#pragma unroll
    for (int c0 = 0; c0 < kN; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) zrow[c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- diagonal sum: output k of the tile = row 32 + k ------------------------------------------------------
    if (tid < kOutPerTile) {
      const int rr = 32 + tid;
      float re = 0.f, im = 0.f;
#pragma unroll
      for (int q = 0; q < kQ; ++q) {
        re += zs[(rr - q) * 81 + 2 * q];
        im += zs[(rr - q) * 81 + 2 * q + 1];
      }
      y[(size_t)tile * kOutPerTile + tid] = make_float2(re, im);
    }
    __syncthreads();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

static double izero(double v) {
  double sum = 1, u = 1, h = v / 2; int n = 1;
  do { double t = h / n; n++; t *= t; u *= t; sum += u; } while (u >= 1e-21 * sum);
  return sum;
}

int main(int argc, char **argv) {
  const int n_tiles = argc > 1 ? atoi(argv[1]) : 296 * 64;
  const size_t n_blocks = (size_t)n_tiles * kOutPerTile + 32;
  // 525 Kaiser taps of rational_resampler_ccc(1, 16) (same design as ltb_tables.cpp, in double)
  const int ntaps = 525, M2 = 262;
  std::vector<double> taps(528, 0.0);
  {
    const double beta = 7.0, tw = 0.1 / 16, mid = 0.5 / 16 - tw / 2, fw = 2 * M_PI * mid;
    double g = 0;
    for (int n = -M2; n <= M2; n++) {
      const double t = 2.0 * (n + M2) / (ntaps - 1) - 1, w = izero(beta * sqrt(1 - t * t)) / izero(beta);
      taps[n + M2] = (n == 0 ? fw / M_PI : sin(n * fw) / (n * M_PI)) * w;
      g += taps[n + M2];
    }
    for (auto &t : taps) t /= g;
  }
  std::vector<float> tmat(kN * kK, 0.f);
  for (int q = 0; q < kQ; q++)
    for (int p = 0; p < 16; p++)
      for (int c = 0; c < 2; c++) tmat[(2 * q + c) * kK + 2 * p + c] = (float)taps[16 * q + 15 - p];
  std::vector<float> x(n_blocks * kK);
  srand(1);
  for (auto &v : x) v = (float)rand() / RAND_MAX - 0.5f;
  float *d_x, *d_t; float2 *d_y;
  cudaMalloc(&d_x, x.size() * 4); cudaMalloc(&d_t, tmat.size() * 4); cudaMalloc(&d_y, (size_t)n_tiles * kOutPerTile * 8);
  cudaMemcpy(d_x, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_t, tmat.data(), tmat.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 32768 + 20480 + 128 * 81 * 4;
  cudaFuncSetAttribute(tc_decim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = n_tiles < 296 ? n_tiles : 296;
  tc_decim_kernel<<<grid, 128, smem>>>(d_x, d_t, d_y, n_tiles);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; i++) tc_decim_kernel<<<grid, 128, smem>>>(d_x, d_t, d_y, n_tiles);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  std::vector<float2> y((size_t)n_tiles * kOutPerTile);
  cudaMemcpy(y.data(), d_y, y.size() * 8, cudaMemcpyDeviceToHost);
  // check a sample of outputs against double precision: y[k] = sum_j taps[j] x[16 (k + 32) + 15 - j]
  double max_err = 0, max_ref = 0;
  for (size_t k = 0; k < y.size(); k += 997) {
    double re = 0, im = 0;
    for (int q = 0; q < kQ; q++)
      for (int p = 0; p < 16; p++) {
        const double t = (double)(float)taps[16 * q + 15 - p];
        const size_t b = k + 32 - q;
        re += t * x[b * kK + 2 * p]; im += t * x[b * kK + 2 * p + 1];
      }
    max_err = fmax(max_err, fmax(fabs(re - y[k].x), fabs(im - y[k].y)));
    max_ref = fmax(max_ref, fmax(fabs(re), fabs(im)));
  }
  const double in_samples = (double)n_tiles * kOutPerTile * 16;
  printf("{\"tiles\": %d, \"ms\": %.4f, \"input_Gsamples_per_s\": %.1f, \"GB_per_s_fc32\": %.1f, \"max_abs_err\": %.3e, "
         "\"max_ref\": %.3e, \"rel_err\": %.3e}\n",
         n_tiles, ms, in_samples / ms / 1e6, in_samples * 8 / ms / 1e6, max_err, max_ref, max_err / max_ref);
  return 0;
}
