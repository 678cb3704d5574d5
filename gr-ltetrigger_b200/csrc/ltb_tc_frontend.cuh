// LTB_FRONTEND_TC_INT: the D = 16 decimating front end for integer input (sc16) on the 5th-generation
// tensor cores, in exact integer arithmetic.  Included by ltb_api.cu and by tools/ubench_tc_i8.cu.
//
// Arithmetic (what the oracle's ORC_FRONT_TCINT restates with int64 on the CPU):
//   T[j] = rint(taps[j] * 2^27)                        525 integers, |T| < 2^23, three balanced base-256 digits
//   A[k] = sum_j T[j] * x[16 k - j]                    exact (|A| < 2^44), per component, x = the int16 samples
//   y[k] = float32(A[k]) * 2^-42                       ONE rounding per output (2^-27 taps, 2^-15 sc16 scale)
// Nothing else rounds, so the result does not depend on accumulation order: the tensor core's int32
// accumulators, the digit recombination and the diagonal sums below are all exact.
//
// GEMM form.  A row of the A operand is 256 consecutive samples of one component of one stream (16 outputs),
// as the 512 bytes they occupy (lo byte XOR 0x80 -> signed "lo - 128", hi byte signed: x = 256 hi + lo' + 128;
// the + 128 adds the constant 128 * sum(T) to every output).  Sixteen k-steps of 32 bytes (16 samples) each:
//   D[row][4 u + v] += sum_{p', byte} A[row][s][2 p' + byte] * B0[4 (u - s) + v][2 p' + byte]      u - s = 0..33
//   B0[4 d + v][2 p' + 0] = digit_v(T[16 d - p'])      B0[4 d + v][2 p' + 1] = digit_{v-1}(T[16 d - p'])
// i.e. every k-step multiplies by the SAME small tap matrix B0 (136 rows of 32 bytes) and accumulates into the
// accumulator tile at a column offset that advances by four columns per k-step (a banded Toeplitz product
// without materialising the band).  u = 16 q + r: column group u of row b holds the contribution of row b to
// output r of row b + q; the epilogue recombines the four weights v (value = sum_v 256^v D[.][4u+v], int64),
// adds the q = 1..3 contributions from the rows above (warp shuffles, a small shared-memory exchange at the warp
// boundary), converts once and stores.  An M = 128 tile is 64 stream rows: lanes 0..63 hold the real parts,
// lanes 64..127 the imaginary parts, so one accumulator buffer is 208 TMEM columns and two fit (MMA of tile
// t+1 runs while the epilogue drains tile t).  Rows 0..2 of a tile only feed the rows below them (halo): a tile
// yields 61 rows = 976 outputs.
//
// Pipeline (one 320-thread CTA per SM, persistent over a contiguous run of tiles):
//   warp 4      TMA producer: one cp.async.bulk.tensor.3d box (256 B x 64 rows) per stage -> raw ring (4 stages)
//   warps 6..9  transform: raw interleaved int16 I/Q -> two planar rows (re, im) with the lo bytes flipped,
//               written in the UMMA K-major SWIZZLE_128B layout (3 stages of 16 kB)
//   warp 5      MMA issuer: 4 x tcgen05.mma.kind::i8 (M = 128, N = 144, K = 32) per stage, accumulators in TMEM
//   warps 0..3  epilogue: tcgen05.ld -> int64 recombination -> diagonal sum -> float -> y_ring (coalesced)
// All hand-offs are mbarriers; every wait has a watchdog (a protocol error traps instead of hanging).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ltb {

constexpr int kTcRowSamples = 256;                 // input samples per A row (16 outputs)
constexpr int kTcTileRows = 64;                    // stream rows per tile (x 2 components = M 128)
constexpr int kTcHalo = 3;                         // rows a tile re-reads from the tile above
constexpr int kTcUseful = kTcTileRows - kTcHalo;   // 61
constexpr int kTcAtomSamples = 64;                 // samples per row and pipeline stage (128 B per component)
constexpr int kTcRawStages = 4, kTcAStages = 3;
constexpr int kTcRawBytes = kTcTileRows * kTcAtomSamples * 4;   // 16384
constexpr int kTcABytes = 128 * 128;                            // 16384
constexpr int kTcBRows = 208;                      // accumulator columns of one tile (49 u's x 4, padded to 16)
constexpr int kTcBTileBytes = kTcBRows * 128;      // 26624: four 32-byte k-slices per row
constexpr int kTcThreads = 320;
constexpr int kTcTailSamples = kTcHalo * kTcRowSamples;   // 768 raw samples of history per stream
constexpr int kTcTapShift = 27;
constexpr int kTcStagePitch = 17;                  // floats per (row, component) in the output staging

__host__ __device__ constexpr int tc_btiles(int G) { return (G + 3) / 4; }
__host__ __device__ constexpr int tc_n_step(int G) { return (4 * (34 + G - 1) + 15) / 16 * 16; }   // N of a k-step's MMA
__host__ __device__ constexpr size_t tc_smem_bytes(int G) {
  return 1024 /*alignment slack*/ + (size_t)tc_btiles(G) * kTcBTileBytes + (size_t)kTcAStages * kTcABytes +
         (size_t)kTcRawStages * kTcRawBytes + 2 * kTcTileRows * kTcStagePitch * 4 + 2 * 3 * 33 * 8 + 256;
}

struct TcParams {
  const void *in;              // [n_streams] rows of interleaved int16 I/Q
  long long stride_bytes;
  int n_in;                    // new input samples per stream (multiple of 128)
  int n_streams;
  const short2 *tail;          // [n_streams][768]: the 768 samples before this chunk (zeros at stream start)
  float2 *y_ring;
  long long n_base;
  unsigned cap_mask;
  int cap;
  int tiles_per_stream, total_tiles;
  const int8_t *btab;          // [tc_btiles(G)][208][128]: tap tables, k-slice i = table for k-step s with s % G == i
  long long c_const;           // 128 * sum_j T[j]
  int *err;                    // device flag: 0 ok, else the code of the watchdog that fired
  int *dbg_acc;                // null, or [128][208] int32: the raw accumulator tile of tile 0 (tools/ubench_tc_i8)
};

// ---- PTX helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: ~2 s of spinning means a protocol error -> record the code and trap (no hang)
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity, int *err, int code) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000LL) {
      atomicExch(err, code);
      __threadfence_system();
      asm volatile("trap;");
    }
  }
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (8-row groups 1024 B apart), as tools/ubench_tc_decim.cu
__device__ __forceinline__ uint64_t tc_make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::i8: D = S32 (2 << 4), A and B signed 8 bit (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t tc_idesc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// byte offset of (row r, 16-byte chunk c) in a [rows][128 B] K-major tile with the 128-byte swizzle
__device__ __forceinline__ int tc_sw_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }

__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tc_tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// value of one column group: R0 + 256 R1 + 65536 R2 + 2^24 R3, exact in int64
__device__ __forceinline__ long long tc_combine(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  long long p = (long long)(int)r3;
  p = p * 256 + (long long)(int)r2;
  p = p * 256 + (long long)(int)r1;
  p = p * 256 + (long long)(int)r0;
  return p;
}
__device__ __forceinline__ long long tc_shfl_up(long long v, int delta) {
  const int lo = __shfl_up_sync(0xffffffffu, (int)(v & 0xffffffffLL), delta);
  const int hi = __shfl_up_sync(0xffffffffu, (int)(v >> 32), delta);
  return ((long long)hi << 32) | (unsigned)lo;
}

// ---- the kernel ----------------------------------------------------------------------------------------------
// G: k-steps that share one accumulator column offset (the offset advances by 4 G columns); 1 is the cheapest,
// larger values are kept for hardware whose tcgen05.mma wants a coarser column alignment of D.
template <int G>
__global__ void __launch_bounds__(kTcThreads, 1)
decimate_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams P) {
  constexpr int NB = tc_btiles(G), NSTEP = tc_n_step(G);
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *s_b = smem;                                            // [NB][208][128]
  unsigned char *s_a = s_b + (size_t)NB * kTcBTileBytes;                // [kTcAStages][128][128]
  unsigned char *s_raw = s_a + (size_t)kTcAStages * kTcABytes;          // [kTcRawStages][64][256]
  float *s_stage = reinterpret_cast<float *>(s_raw + (size_t)kTcRawStages * kTcRawBytes);   // [2][64][17]
  long long *s_xchg = reinterpret_cast<long long *>(s_stage + 2 * kTcTileRows * kTcStagePitch);   // [2][3][33]
  __shared__ __align__(8) unsigned long long bars[2 * kTcRawStages + 2 * kTcAStages + 4];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar0 = tc_smem_u32(bars);
  auto raw_full = [&](int i) { return bar0 + 8u * i; };
  auto raw_empty = [&](int i) { return bar0 + 8u * (kTcRawStages + i); };
  auto a_full = [&](int i) { return bar0 + 8u * (2 * kTcRawStages + i); };
  auto a_empty = [&](int i) { return bar0 + 8u * (2 * kTcRawStages + kTcAStages + i); };
  auto acc_full = [&](int i) { return bar0 + 8u * (2 * kTcRawStages + 2 * kTcAStages + i); };
  auto acc_empty = [&](int i) { return bar0 + 8u * (2 * kTcRawStages + 2 * kTcAStages + 2 + i); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t_begin = (int)((long long)P.total_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)P.total_tiles * (blockIdx.x + 1) / gridDim.x);
  const int full_rows = P.n_in / kTcRowSamples;             // rows the tensor map covers
  const int m_out = P.n_in / 16;

  // tap tables -> shared memory (generic proxy writes, made visible to the tensor core by the fence below)
  for (int i = tid; i < NB * kTcBTileBytes / 16; i += kTcThreads)
    reinterpret_cast<uint4 *>(s_b)[i] = reinterpret_cast<const uint4 *>(P.btab)[i];
  if (tid == 0) {
    for (int i = 0; i < kTcRawStages; ++i) { tc_mbar_init(raw_full(i), 1); tc_mbar_init(raw_empty(i), 4); }
    for (int i = 0; i < kTcAStages; ++i) { tc_mbar_init(a_full(i), 4); tc_mbar_init(a_empty(i), 1); }
    for (int i = 0; i < 2; ++i) { tc_mbar_init(acc_full(i), 1); tc_mbar_init(acc_empty(i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  if (warp == 4) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int stream = t / P.tiles_per_stream, ti = t - stream * P.tiles_per_stream;
        const int row0 = ti * kTcUseful - kTcHalo;
#ifdef LTB_TC_NO_TMA
        const bool any = false;                                          // bisection build: every row takes the patch path
#else
        const bool any = row0 + kTcTileRows > 0 && row0 < full_rows;     // rows out of range are zero-filled
#endif
        for (int a = 0; a < 4; ++a, ++it) {
          const int rs = it % kTcRawStages;
          tc_mbar_wait(raw_empty(rs), ((it / kTcRawStages) & 1) ^ 1, P.err, 1);
          if (any) {
            tc_mbar_expect_tx(raw_full(rs), kTcRawBytes);
            tc_tma_load_3d(tc_smem_u32(s_raw + (size_t)rs * kTcRawBytes), &tmap, a * 256, row0, stream, raw_full(rs));
          } else {
            tc_mbar_arrive(raw_full(rs));
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    int it = 0, tl = 0;
    for (int t = t_begin; t < t_end; ++t, ++tl) {
      const int buf = tl & 1;
      tc_mbar_wait(acc_empty(buf), ((tl >> 1) & 1) ^ 1, P.err, 2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int a = 0; a < 4; ++a, ++it) {
        const int as = it % kTcAStages;
        tc_mbar_wait(a_full(as), (it / kTcAStages) & 1, P.err, 3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t a_base = tc_smem_u32(s_a + (size_t)as * kTcABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int s = 4 * a + k, i = s % G, grp = s / G;
            const bool first = s == 0;
            const uint64_t adesc = tc_make_desc(a_base + 32 * k);
            const uint64_t bdesc = tc_make_desc(tc_smem_u32(s_b + (size_t)(i >> 2) * kTcBTileBytes) + 32 * (i & 3));
            const uint32_t d = tmem + buf * 256 + (first ? 0 : 4 * G * grp);
            tc_mma_i8(d, adesc, bdesc, first ? tc_idesc(kTcBRows) : tc_idesc(NSTEP), first ? 0u : 1u);
          }
          tc_commit(a_empty(as));
          if (a == 3) tc_commit(acc_full(buf));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 6) {
    // ===== transform: raw interleaved int16 -> planar re / im rows, lo bytes flipped, swizzled K-major =====
    const int tt = tid - 6 * 32;                                         // 0..127
    int it = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int stream = t / P.tiles_per_stream, ti = t - stream * P.tiles_per_stream;
      const int row0 = ti * kTcUseful - kTcHalo;
#ifdef LTB_TC_NO_TMA
      const bool patch = true;
#else
      const bool patch = row0 < 0 || row0 + kTcTileRows > full_rows;     // some rows are not plain tensor rows
#endif
      for (int a = 0; a < 4; ++a, ++it) {
        const int rs = it % kTcRawStages, as = it % kTcAStages;
        tc_mbar_wait(raw_full(rs), (it / kTcRawStages) & 1, P.err, 4);
        tc_mbar_wait(a_empty(as), ((it / kTcAStages) & 1) ^ 1, P.err, 5);
        const unsigned char *raw = s_raw + (size_t)rs * kTcRawBytes;
        unsigned char *dst = s_a + (size_t)as * kTcABytes;
#pragma unroll 4
        for (int j = 0; j < kTcTileRows * 16 / 128; ++j) {               // 64 rows x 16 four-sample items, 128 threads
          const int item = j * 128 + tt, row = item >> 4, c = item & 15;
          uint4 w = *reinterpret_cast<const uint4 *>(raw + row * 256 + c * 16);
          if (patch) {
            const int ra = row0 + row;
            if (ra < 0) {                                                // history: the carried tail (rows -3..-1)
              w = *reinterpret_cast<const uint4 *>(P.tail + (size_t)stream * kTcTailSamples + (ra + kTcHalo) * kTcRowSamples +
                                                   a * kTcAtomSamples + c * 4);
#ifdef LTB_TC_NO_TMA
            } else {
#else
            } else if (ra >= full_rows) {                                // the partial last row, then nothing
#endif
              const int n0 = ra * kTcRowSamples + a * kTcAtomSamples + c * 4;
              const unsigned *src = reinterpret_cast<const unsigned *>((const char *)P.in + (long long)stream * P.stride_bytes);
              w.x = n0 + 0 < P.n_in ? src[n0 + 0] : 0u;
              w.y = n0 + 1 < P.n_in ? src[n0 + 1] : 0u;
              w.z = n0 + 2 < P.n_in ? src[n0 + 2] : 0u;
              w.w = n0 + 3 < P.n_in ? src[n0 + 3] : 0u;
            }
          }
          uint2 re, im;
          re.x = __byte_perm(w.x, w.y, 0x5410) ^ 0x00800080u;
          re.y = __byte_perm(w.z, w.w, 0x5410) ^ 0x00800080u;
          im.x = __byte_perm(w.x, w.y, 0x7632) ^ 0x00800080u;
          im.y = __byte_perm(w.z, w.w, 0x7632) ^ 0x00800080u;
          const int off = (c & 1) * 8;
          *reinterpret_cast<uint2 *>(dst + tc_sw_off(row, c >> 1) + off) = re;
          *reinterpret_cast<uint2 *>(dst + tc_sw_off(64 + row, c >> 1) + off) = im;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) { tc_mbar_arrive(a_full(as)); tc_mbar_arrive(raw_empty(rs)); }
      }
    }
  } else {
    // ===== epilogue (warps 0..3 = TMEM lane quarters; warps 0,1: re rows 0..63, warps 2,3: im rows 0..63) =====
    const int comp = warp >> 1, row = (warp & 1) * 32 + lane;
    int tl = 0;
    for (int t = t_begin; t < t_end; ++t, ++tl) {
      const int stream = t / P.tiles_per_stream, ti = t - stream * P.tiles_per_stream;
      const int row0 = ti * kTcUseful - kTcHalo;
      const int buf = tl & 1;
      tc_mbar_wait(acc_full(buf), (tl >> 1) & 1, P.err, 6);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem + buf * 256 + ((uint32_t)(warp * 32) << 16);
      long long p[49];
#pragma unroll
      for (int c0 = 0; c0 < 6; ++c0) {                                   // columns 0..191: u = 0..47
        uint32_t r[32];
        tc_ld32(taddr + 32 * c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int u = 0; u < 8; ++u) p[8 * c0 + u] = tc_combine(r[4 * u], r[4 * u + 1], r[4 * u + 2], r[4 * u + 3]);
        if (P.dbg_acc && t == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) P.dbg_acc[(warp * 32 + lane) * kTcBRows + 32 * c0 + i] = (int)r[i];
        }
      }
      {
        uint32_t r[4];                                                   // columns 192..195: u = 48
        tc_ld4(taddr + 192, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        p[48] = tc_combine(r[0], r[1], r[2], r[3]);
        if (P.dbg_acc && t == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) P.dbg_acc[(warp * 32 + lane) * kTcBRows + 192 + i] = (int)r[i];
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(acc_empty(buf));                     // the MMA warp may refill this buffer

      // contributions to the rows below: rows 29..31 of the upper warp hand theirs over in shared memory
      if ((warp & 1) == 0 && lane >= 29) {
        long long *x = s_xchg + (comp * 3 + (lane - 29)) * 33;
#pragma unroll
        for (int u = 16; u < 49; ++u) x[u - 16] = p[u];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float *stg = s_stage + (comp * kTcTileRows + row) * kTcStagePitch;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        long long acc = p[r];
#pragma unroll
        for (int q = 1; q <= 3; ++q) {
          if (q == 3 && r != 0) continue;
          const int u = 16 * q + r;
          long long v = tc_shfl_up(p[u], q);
          if (lane < q) v = s_xchg[(comp * 3 + (3 + lane - q)) * 33 + (u - 16)];   // row 32 + lane - q of the upper warp
          acc += v;
        }
        acc += P.c_const;
        stg[r] = __fmul_rn(__ll2float_rn(acc), 2.2737367544323206e-13f);   // 2^-42
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // coalesced store: 16 consecutive float2 per row
      {
        float2 *yr = P.y_ring + (size_t)stream * P.cap;
        const int rr = tid & 15;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rw = (tid >> 4) + 8 * j;                             // tile row 0..63
          const long long k = (long long)(row0 + rw) * 16 + rr;
          if (rw >= kTcHalo && k < m_out) {
            const float2 v = make_float2(s_stage[(0 * kTcTileRows + rw) * kTcStagePitch + rr],
                                         s_stage[(1 * kTcTileRows + rw) * kTcStagePitch + rr]);
            yr[(unsigned)((P.n_base + k) & P.cap_mask)] = v;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// keep the last 768 raw samples of each stream for the next call (the three halo rows of its first tile)
__global__ void __launch_bounds__(256) tc_tail_kernel(const void *__restrict__ in, long long stride_bytes, int n_in,
                                                      const short2 *__restrict__ tail_old, short2 *__restrict__ tail_new) {
  const int stream = blockIdx.x;
  const short2 *src = reinterpret_cast<const short2 *>((const char *)in + (long long)stream * stride_bytes);
  for (int i = threadIdx.x; i < kTcTailSamples; i += blockDim.x) {
    const int idx = n_in - kTcTailSamples + i;
    tail_new[(size_t)stream * kTcTailSamples + i] = idx >= 0 ? src[idx] : tail_old[(size_t)stream * kTcTailSamples + kTcTailSamples + idx];
  }
}

}  // namespace ltb
