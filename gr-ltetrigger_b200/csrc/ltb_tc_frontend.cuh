// LTB_FRONTEND_TC_INT: the decimating front end on the 5th-generation tensor cores, in exact integer arithmetic, for
// sc16 / sc8 input and for fc32 input taken as 23-bit fixed point, at D = 2 ... 32 (TcGeom below lists the pairs).
// Included by ltb_api.cu and by tools/ubench_tc_i8.cu.  Described for D = 16; "geometry" below has the other rates.
//
// Arithmetic (what the oracle's ORC_FRONT_TCINT restates with int64 on the CPU):
//   T[j] = rint(taps[j] * 2^27)                        525 integers, |T| < 2^23, three balanced base-256 digits
//   A[k] = sum_j T[j] * x[16 k - j]                    exact (|A| < 2^44), per component, x = the integer samples
//   y[k] = float32(A[k]) * 2^-42  (sc8: 2^-34)         ONE rounding per output (2^-27 taps, 2^-15 / 2^-7 input scale)
// Nothing else rounds, so the result does not depend on accumulation order: the tensor core's int32
// accumulators, the digit recombination and the diagonal sums below are all exact.
// fc32: x = q(sample) - 2^22 with q() the two-FFMA fixed-point step at tc_split below (1 <= q < 2^23), the product of the
// lowest sample byte with the lowest tap digit is not formed, y[k] = float32(A[k] / 256) * float32(F / (2^22 - 1) * 2^-19).
//
// GEMM form (sc16).  A row of the A operand is 256 consecutive samples of one component of one stream (16
// outputs), as the 512 bytes they occupy (lo byte XOR 0x80 -> signed "lo - 128", hi byte signed:
// x = 256 hi + lo' + 128; the + 128 adds the constant 128 * sum(T) to every output).  Sixteen k-steps of 32
// bytes (16 samples) each:
//   D[row][4 u + v] += sum_{p', byte} A[row][s][2 p' + byte] * B0[4 (u - s) + v][2 p' + byte]      u - s = 0..33
//   B0[4 d + v][2 p' + 0] = digit_v(T[16 d - p'])      B0[4 d + v][2 p' + 1] = digit_{v-1}(T[16 d - p'])
// i.e. every k-step multiplies by the SAME small tap matrix B0 (136 rows of 32 bytes) and accumulates into the
// accumulator tile at a column offset that advances by four columns per k-step (a banded Toeplitz product
// without materialising the band).  sc8: a row is 256 bytes per component, eight k-steps of 32 samples,
// u - 2 s = 0..34, B0[4 d + v][p'] = digit_v(T[16 d - p']) (v = 3 stays zero), the offset advances by eight
// columns per k-step, the samples are signed bytes as they are (no flip, no constant).
// u = 16 q + r: column group u of row b holds the contribution of row b to output r of row b + q; the
// epilogue recombines the four weights v (value = sum_v 256^v D[.][4u+v], int64), adds the q = 1..3
// contributions from the rows above (warp shuffles, a small shared-memory exchange at the warp boundary),
// converts once and stores.  An M = 128 tile is 64 stream rows: lanes 0..63 hold the real parts, lanes
// 64..127 the imaginary parts, so one accumulator buffer is 208 TMEM columns and two fit (MMA of tile t+1 runs
// while the epilogue drains tile t).  Rows 0..2 of a tile only feed the rows below them (halo): a tile yields
// 61 rows = 976 outputs.
//
// Pipeline (one 576-thread CTA per SM, persistent over a contiguous run of tiles):
//   warp 8        TMA producer: one cp.async.bulk.tensor.3d box (256 B x 64 rows) per stage -> raw ring (6 stages)
//   warps 10..17  transform: raw interleaved I/Q -> two planar rows (re, im), written in the UMMA K-major
//                 SWIZZLE_128B layout (3 stages of 16 kB), in two groups of four warps that take alternate stages
//   warp 9        MMA issuer: 4 x tcgen05.mma.kind::i8 (M = 128, N = 144, K = 32) per stage, accumulators in TMEM
//   warps 0..7    epilogue: tcgen05.ld -> int64 recombination -> diagonal sum -> float -> y_ring (coalesced);
//                 warp w reads TMEM lane quarter w % 4 and owns outputs r = 8 (w / 4) .. 8 (w / 4) + 7 of its rows
//                 (column groups u = r, 16 + r, 32 + r: every other 32-column chunk)
// All hand-offs are mbarriers; every wait has a watchdog (a protocol error traps instead of hanging).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/ltetrigger_b200.h"

namespace ltb {

constexpr int kTcRowSamples = 256;                 // input samples per A row (16 outputs) at D = 16; in general 16 D
constexpr int kTcTileRows = 64;                    // stream rows per tile (x 2 components = M 128)
constexpr int kTcHalo = 3;                         // rows a tile re-reads from the tile above
constexpr int kTcUseful = kTcTileRows - kTcHalo;   // 61
#ifndef LTB_TC_RAW_STAGES
#define LTB_TC_RAW_STAGES 6
#endif
#ifndef LTB_TC_A_STAGES
#define LTB_TC_A_STAGES 3      // measured: three planar stages beat four by 13 % on sc16 (profiles/tc_sweep_r02.log)
#endif
#ifndef LTB_TC_BISECT
#define LTB_TC_BISECT 0     // profiling builds only (tools/ubench_tc_i8): 1 no transform work, 2 no MMAs, 4 no epilogue work, 8 no proxy fence,
                            // 16 epilogue stops after its TMEM loads, 32 epilogue without its TMEM loads
#endif
#ifndef LTB_TC_LB_THREADS
#define LTB_TC_LB_THREADS kTcThreads               // a larger value here only lowers the register budget ptxas may use
#endif
constexpr int kTcRawStages = LTB_TC_RAW_STAGES, kTcAStages = LTB_TC_A_STAGES;
constexpr int kTcRawBytes = kTcTileRows * 256;     // 16384: 256 raw bytes per row and stage (64 sc16 / 128 sc8 samples)
constexpr int kTcABytes = 128 * 128;               // 16384
constexpr int kTcBRows = 208;                      // accumulator columns of one tile (49 u's x 4, padded to 16)
constexpr int kTcBTileBytes = kTcBRows * 128;      // 26624 (the first 32 bytes of each 128-byte row are used)
constexpr int kTcEpiWarps = 8, kTcXformWarps = 8;  // warps 0..7; 8 producer, 9 MMA; 10..17
#ifndef LTB_TC_XFORM_GROUPS
#define LTB_TC_XFORM_GROUPS 2
#endif
constexpr int kTcXformGroups = LTB_TC_XFORM_GROUPS;                 // groups of transform warps that take alternate stages
constexpr int kTcXformGroupWarps = kTcXformWarps / kTcXformGroups;
constexpr int kTcThreads = 32 * (kTcEpiWarps + 2 + kTcXformWarps);   // 576
constexpr int kTcTailSamples = kTcHalo * 16 * 32;         // raw samples of history per stream at D = 32 (the largest rate built)
constexpr int kTcTapShift = 27;                    // at D = 16; tc_tap_shift(D) in general

// ---- geometry of the (format, decimation) variants ------------------------------------------------------------
// A row is always 16 outputs = 16 D input samples of one component; a k-step is 32 bytes of it = SPK samples
// (sc16 16, sc8 32, fc32 8), i.e. SPK / D outputs: the accumulator column offset of k-step s is 4 floor(s SPK / D) and the
// taps it meets are T[D d - phase - p] with phase = (s SPK) mod D -- one tap table per distinct phase (one when D
// divides SPK; two for fc32 at D = 16; three at D = 12), side by side in the 128-byte rows of the table tile.
__host__ __device__ constexpr int tc_gcd(int a, int b) { return b == 0 ? a : tc_gcd(b, a % b); }
__host__ __device__ constexpr int tc_ntaps(int D) {           // gr-filter's default design (ltb_tables.cpp make_decim_taps)
  return (int)((7.0 / 0.1102 + 8.7) / (22.0 * 0.1 / D)) + 1 - ((int)((7.0 / 0.1102 + 8.7) / (22.0 * 0.1 / D)) & 1);
}
__host__ __device__ constexpr int tc_log2(int D) { return D <= 1 ? 0 : 1 + tc_log2(D / 2); }
__host__ __device__ constexpr int tc_tap_shift(int D) { return 23 + tc_log2(D); }    // max |T| = 0.9 / 0.6 x 2^23: three digits
template <int FMT, int D>
struct TcGeom {
  static constexpr int BPS = FMT == LTB_FMT_FC32 ? 8 : FMT == LTB_FMT_SC16 ? 4 : 2;   // bytes per complex input sample
  static constexpr int ROW = 16 * D;                           // samples per A row
  static constexpr int ROW_BYTES = ROW * BPS;
  static constexpr int SPT = ROW_BYTES / 256;                  // pipeline stages (128-byte A atoms) per tile
  static constexpr int SPK = 64 / BPS;                         // samples per k-step
  static constexpr int KSTEPS = ROW / SPK;                     // = 4 SPT
  static constexpr int G = tc_gcd(SPK, D);
  static constexpr int NPH = D / G;                            // distinct phases (tap tables)
  static constexpr int NTAPS = tc_ntaps(D);
  static constexpr int ND = (NTAPS - 1 + (D - G) + SPK - 1) / D + 1;   // tap-table row groups a k-step can meet
  static constexpr int NSTEP = (4 * ND + 15) / 16 * 16;        // N of a k-step's MMA
  static constexpr int TAIL = kTcHalo * ROW;                   // raw samples of history per stream
  static constexpr bool ok = ROW_BYTES % 256 == 0 && NPH <= 4 && 4 * (((KSTEPS - 1) * SPK) / D) + NSTEP <= kTcBRows &&
                             (ROW - 1 + NTAPS - 1) / D <= 48 && NTAPS <= 3 * ROW + D;   // a row reaches output 48 of its tile at most
};
__host__ inline bool tc_supported(int fmt, int D) {
  switch (fmt * 100 + D) {
    case 2: case 4: case 8: case 12: case 16: case 24: case 32: return true;               // fc32
    case 104: case 108: case 112: case 116: case 124: case 132: return true;               // sc16
    case 208: case 216: case 224: case 232: return true;                                   // sc8
  }
  return false;
}
constexpr int kTcStagePitch = 17;                  // floats per (row, component) in the output staging

__host__ __device__ constexpr int tc_sample_bytes(int fmt) { return fmt == LTB_FMT_FC32 ? 8 : fmt == LTB_FMT_SC16 ? 4 : 2; }
// fc32 input: fixed point with 23 bits.  q(x) = bits(fma(fma.sat(x, 0.5 / full_scale, 0.5), 2^23 - 2, 2^23 + 1)) & 0x7fffff
// is an integer 1 .. 2^23 - 1 whose distance from 2^22 is x / full_scale * (2^22 - 1) rounded (two roundings, both
// restated by the oracle); values beyond +-full_scale saturate.
constexpr float kTcQMul = 8388606.0f, kTcQAdd = 8388609.0f;
constexpr int kTcQMid = 1 << 22;
__host__ __device__ constexpr size_t tc_smem_bytes() {
  return (size_t)kTcBTileBytes + (size_t)kTcAStages * kTcABytes + (size_t)kTcRawStages * kTcRawBytes +
         2 * kTcTileRows * kTcStagePitch * 4 + 2 * 2 * 3 * 17 * 8;
}

struct TcParams {
  const void *in;              // [n_streams] rows of interleaved integer I/Q
  long long stride_bytes;
  int n_in;                    // new input samples per stream (multiple of 128)
  int n_streams;
  const void *tail;            // [n_streams][768] raw samples before this chunk (zeros at stream start)
  float2 *y_ring;
  long long n_base;
  unsigned cap_mask;
  int cap;
  int tiles_per_stream, total_tiles;
  const int8_t *btab;          // [208][128]: tap table in its shared-memory image (ltb_tables.cpp make_tc_btab)
  long long c_const;           // sc16: 128 * sum_j T[j]; sc8: 0; fc32: -2^14 * sum_j T[j]
  float q_inv;                 // fc32: 0.5 / full_scale
  float out_scale;             // fc32: float(full_scale / (2^22 - 1) * 2^-(shift - 8)); sc16 2^-(shift + 15), sc8 2^-(shift + 7)
  int m_out;                   // outputs per stream of this call: n_in / D
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
// bounded wait: ~2 s without progress means a protocol error -> record the code and trap (no hang).
// BACKOFF: sleep between polls (roles that wait long: their polling would take issue slots from the others)
template <int BACKOFF>
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity, int *err, int code) {
  uint32_t done = 0;
  long long t0 = 0;
  for (int spins = 0;; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (BACKOFF) __nanosleep(BACKOFF);
    if ((spins & 1023) == 1023) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) {
        atomicExch(err, code);
        __threadfence_system();
        asm volatile("trap;");
      }
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
// kind::i8: D = S32 (2 << 4), A signed (1 << 7) or unsigned (0) 8 bit, B signed 8 bit (1 << 10), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t tc_idesc(int n, bool a_signed = true) {
  return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// byte offset of (row r, 16-byte chunk c) in a [rows][128 B] K-major tile with the 128-byte swizzle
__host__ __device__ constexpr int tc_sw_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }

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
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
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
// value of one column group: R0 + 256 R1 + 65536 R2 + 2^24 R3, exact in int64 (three IMAD.WIDE)
__device__ __forceinline__ long long tc_combine(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  return (long long)(int)r0 + (long long)(int)r1 * 256 + (long long)(int)r2 * 65536 + (long long)(int)r3 * 16777216;
}
__device__ __forceinline__ long long tc_shfl_up(long long v, int delta) {
  const int lo = __shfl_up_sync(0xffffffffu, (int)(v & 0xffffffffLL), delta);
  const int hi = __shfl_up_sync(0xffffffffu, (int)(v >> 32), delta);
  return ((long long)hi << 32) | (unsigned)lo;
}

// 16 raw bytes -> the planar (re, im) halves of an A row, 8 bytes each.  sc16: four samples, lo bytes flipped
// (XOR 0x80: unsigned lo -> signed lo - 128); sc8: eight samples, signed bytes as they are.
template <int FMT>
__device__ __forceinline__ void tc_split(const uint4 w, uint2 &re, uint2 &im, const float q_inv) {
  if (FMT == LTB_FMT_FC32) {
    // two samples (re0, im0, re1, im1) -> their 23-bit fixed-point words; the top byte of each word (the float's
    // exponent bits, 0x4b) meets a zero row of the tap table, so it is not masked off
    float u;
    asm("fma.rn.sat.f32 %0, %1, %2, 0f3F000000;" : "=f"(u) : "f"(__uint_as_float(w.x)), "f"(q_inv));
    re.x = __float_as_uint(__fmaf_rn(u, kTcQMul, kTcQAdd));
    asm("fma.rn.sat.f32 %0, %1, %2, 0f3F000000;" : "=f"(u) : "f"(__uint_as_float(w.z)), "f"(q_inv));
    re.y = __float_as_uint(__fmaf_rn(u, kTcQMul, kTcQAdd));
    asm("fma.rn.sat.f32 %0, %1, %2, 0f3F000000;" : "=f"(u) : "f"(__uint_as_float(w.y)), "f"(q_inv));
    im.x = __float_as_uint(__fmaf_rn(u, kTcQMul, kTcQAdd));
    asm("fma.rn.sat.f32 %0, %1, %2, 0f3F000000;" : "=f"(u) : "f"(__uint_as_float(w.w)), "f"(q_inv));
    im.y = __float_as_uint(__fmaf_rn(u, kTcQMul, kTcQAdd));
  } else if (FMT == LTB_FMT_SC16) {
    re.x = __byte_perm(w.x, w.y, 0x5410) ^ 0x00800080u;
    re.y = __byte_perm(w.z, w.w, 0x5410) ^ 0x00800080u;
    im.x = __byte_perm(w.x, w.y, 0x7632) ^ 0x00800080u;
    im.y = __byte_perm(w.z, w.w, 0x7632) ^ 0x00800080u;
  } else {
    re.x = __byte_perm(w.x, w.y, 0x6420);
    re.y = __byte_perm(w.z, w.w, 0x6420);
    im.x = __byte_perm(w.x, w.y, 0x7531);
    im.y = __byte_perm(w.z, w.w, 0x7531);
  }
}

// ---- epilogue role (warps 0..7 of both kernels): warp w reads TMEM lane quarter w % 4 (w % 4 = 0,1: re rows 0..63; 2,3: im
// rows 0..63) and owns the outputs r = 8 h .. 8 h + 7 (h = w / 4) of its 32 rows: column groups u = 8 (2 c + h) .. + 7, c = 0..2
// SYSTOLIC picks how the q = 0..3 contributions of the rows above are summed -- running sums that move down a row between
// the column groups (80 registers, the accumulator buffer is released after the last sum) or all groups recombined first
// (96 registers, released right after the loads); measured: the first is 4 % faster on fc32, whose tiles are long, the
// second 6 % on sc16 / sc8, where the MMA warp waits for the buffer (profiles/tc_epilogue_r02.log)
template <bool SYSTOLIC>
__device__ __forceinline__ void tc_epilogue_role(const TcParams &P, const uint32_t tmem, const uint32_t acc_full0,
                                                 const uint32_t acc_empty0, float *s_stage, long long *s_xchg,
                                                 const int t_begin, const int t_end) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_out = P.m_out;
  int stream = t_begin / P.tiles_per_stream, ti = t_begin - stream * P.tiles_per_stream;
  auto next_tile = [&]() { if (++ti == P.tiles_per_stream) { ti = 0; ++stream; } };
  auto acc_full = [&](int i) { return acc_full0 + 8u * i; };
  auto acc_empty = [&](int i) { return acc_empty0 + 8u * i; };
  {
    const int quarter = warp & 3, half = warp >> 2;
    const int comp = quarter >> 1, row = (quarter & 1) * 32 + lane;
    int tl = 0;
    for (int t = t_begin; t < t_end; ++t, ++tl, next_tile()) {
      const int row0 = ti * kTcUseful - kTcHalo;
      const int buf = tl & 1;
      tc_mbar_wait<64>(acc_full(buf), (tl >> 1) & 1, P.err, 6);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (LTB_TC_BISECT & 4) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(acc_empty(buf));
        continue;
      }
      const uint32_t taddr = tmem + buf * 256 + ((uint32_t)(quarter * 32) << 16);
      if constexpr (SYSTOLIC) {
        // Output r of row b is p0_b + p1_(b-1) + p2_(b-2) (+ p48_(b-3) for r = 0), p_q = the recombined column group
        // u = 16 q + r.  The groups are taken q = 2, 1, 0 and the running sums move down one row (one lane) between them, so
        // only eight 64-bit sums and one 16-column piece of the accumulator are live at a time.  What crosses from the
        // upper quarter's rows 29..31 into the lower quarter's rows 0..2 goes through shared memory and is added after the
        // barrier; the top quarter's rows 0..2 are the tile's halo (their outputs are not stored).
        long long acc[8];
        long long *xq = s_xchg + (size_t)((comp * 2 + half) * 3) * 17;
        const bool pub = (quarter & 1) == 0 && lane >= 29;
        long long *xp = xq + (lane - 29) * 17;                              // this row's slot when pub
#pragma unroll
        for (int q = 2; q >= 0; --q) {
          const int c0 = 2 * q + half;                                     // 32-column chunk, loaded as two halves
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            uint32_t r[16];
            if (LTB_TC_BISECT & 32) {
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = (uint32_t)(lane * 3 + i + tl);
            } else {
              tc_ld16(taddr + 32 * c0 + 16 * h2, r);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
#ifdef LTB_TC_DBG_ACC
            if (P.dbg_acc && t == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) P.dbg_acc[(quarter * 32 + lane) * kTcBRows + 32 * c0 + 16 * h2 + i] = (int)r[i];
            }
#endif
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int jj = 4 * h2 + j;
              const long long v = tc_combine(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
              if (q == 2) {
                acc[jj] = v;
                if (pub) xp[8 + jj] = v;
              } else {
                if (q == 1 && pub) xp[jj] = v;
                const long long up = tc_shfl_up(acc[jj], 1);
                acc[jj] = (lane == 0 ? 0LL : up) + v;
              }
            }
          }
        }
        if (half == 0) {
          uint32_t r[4];                                                   // columns 192..195: u = 48 (q = 3, r = 0)
          tc_ld4(taddr + 192, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const long long p48 = tc_combine(r[0], r[1], r[2], r[3]);
#ifdef LTB_TC_DBG_ACC
          if (P.dbg_acc && t == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) P.dbg_acc[(quarter * 32 + lane) * kTcBRows + 192 + i] = (int)r[i];
          }
#endif
          if (pub) xp[16] = p48;
          const long long up = tc_shfl_up(p48, 3);
          acc[0] += lane < 3 ? 0LL : up;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(acc_empty(buf));                     // the MMA warp may refill this buffer
        if (LTB_TC_BISECT & 16) continue;

        asm volatile("bar.sync 1, 256;" ::: "memory");
        if ((quarter & 1) == 1 && lane < 3) {                              // rows 32..34: what rows 29..31 contribute
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (lane == 0) acc[j] += xq[2 * 17 + j] + xq[1 * 17 + 8 + j];  // p1 of row 31 + p2 of row 30
            else if (lane == 1) acc[j] += xq[2 * 17 + 8 + j];              // p2 of row 31
          }
          if (half == 0) acc[0] += xq[lane * 17 + 16];                     // p48 of row 29 + lane
        }
        float *stg = s_stage + (comp * kTcTileRows + row) * kTcStagePitch + 8 * half;
#pragma unroll
        for (int j = 0; j < 8; ++j) stg[j] = __fmul_rn(__ll2float_rn(acc[j] + P.c_const), P.out_scale);
      } else {
        long long p[3][8];                                                 // p[q][j]: u = 16 q + 8 half + j
        long long p48 = 0;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int c0 = 2 * q + half;                                     // 32-column chunk, loaded as two halves
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            uint32_t r[16];
            if (LTB_TC_BISECT & 32) {
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = (uint32_t)(lane * 3 + i + tl);
            } else {
              tc_ld16(taddr + 32 * c0 + 16 * h2, r);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) p[q][4 * h2 + j] = tc_combine(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
#ifdef LTB_TC_DBG_ACC
            if (P.dbg_acc && t == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) P.dbg_acc[(quarter * 32 + lane) * kTcBRows + 32 * c0 + 16 * h2 + i] = (int)r[i];
            }
#endif
          }
        }
        if (half == 0) {
          uint32_t r[4];                                                   // columns 192..195: u = 48 (q = 3, r = 0)
          tc_ld4(taddr + 192, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          p48 = tc_combine(r[0], r[1], r[2], r[3]);
#ifdef LTB_TC_DBG_ACC
          if (P.dbg_acc && t == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) P.dbg_acc[(quarter * 32 + lane) * kTcBRows + 192 + i] = (int)r[i];
          }
#endif
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(acc_empty(buf));                     // the MMA warp may refill this buffer
        if (LTB_TC_BISECT & 16) continue;

        // contributions to the rows below: rows 29..31 of the upper quarter hand theirs over in shared memory
        long long *xq = s_xchg + (size_t)((comp * 2 + half) * 3) * 17;
        if ((quarter & 1) == 0 && lane >= 29) {
          long long *x = xq + (lane - 29) * 17;
#pragma unroll
          for (int j = 0; j < 8; ++j) { x[j] = p[1][j]; x[8 + j] = p[2][j]; }
          x[16] = p48;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float *stg = s_stage + (comp * kTcTileRows + row) * kTcStagePitch + 8 * half;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          long long acc = p[0][j];
#pragma unroll
          for (int q = 1; q <= 3; ++q) {
            if (q == 3 && j != 0) continue;
            long long v = tc_shfl_up(q == 3 ? p48 : p[q][j], q);
            if (lane < q) v = xq[(3 + lane - q) * 17 + (q == 3 ? 16 : 8 * (q - 1) + j)];   // row 32 + lane - q of the upper quarter
            if (q == 3 && half != 0) v = 0;                                // u = 48 only feeds output r = 0
            acc += v;
          }
          acc += P.c_const;
          stg[j] = __fmul_rn(__ll2float_rn(acc), P.out_scale);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // coalesced store: 16 consecutive float2 per row; ring offsets in 32 bits (the ring is at most 2^31 entries, so the
      // low word of the absolute index decides the slot)
      {
        float2 *yr = P.y_ring + (size_t)stream * P.cap;
        const int rr = tid & 15, rw0 = tid >> 4;
        const int k0 = (row0 + rw0) * 16 + rr;                           // output index of this thread's first row
        const unsigned slot0 = (unsigned)P.n_base + (unsigned)k0;
        const float *sg = s_stage + rw0 * kTcStagePitch + rr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rw = rw0 + 16 * j;                                   // tile row 0..63
          if (rw >= kTcHalo && k0 + 256 * j < m_out) {
            const float2 v = make_float2(sg[16 * j * kTcStagePitch], sg[(kTcTileRows + 16 * j) * kTcStagePitch]);
            yr[(slot0 + 256u * j) & P.cap_mask] = v;
          }
        }
      }
    }
    }
}

// ---- the kernel ----------------------------------------------------------------------------------------------
template <int FMT, int D>
__global__ void __launch_bounds__(LTB_TC_LB_THREADS, 1)
decimate_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams P) {
  typedef TcGeom<FMT, D> GEO;
  static_assert(GEO::ok, "unsupported (format, decimation) for the tensor-core front end");
  constexpr int BPS = GEO::BPS;                             // bytes per complex input sample
  constexpr int SPT = GEO::SPT;                             // pipeline stages (128-byte A atoms) per tile
  constexpr int ITEM = 16 / BPS;                            // samples per 16-byte transform item
  constexpr int STAGE_SAMPLES = 256 / BPS;                  // samples per row and stage
  constexpr int ROW = GEO::ROW;
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  unsigned char *s_b = tc_smem;                                         // [208][128]
  unsigned char *s_a = s_b + kTcBTileBytes;                             // [kTcAStages][128][128]
  unsigned char *s_raw = s_a + (size_t)kTcAStages * kTcABytes;          // [kTcRawStages][64][256]
  float *s_stage = reinterpret_cast<float *>(s_raw + (size_t)kTcRawStages * kTcRawBytes);   // [2][64][17]
  long long *s_xchg = reinterpret_cast<long long *>(s_stage + 2 * kTcTileRows * kTcStagePitch);   // [2 comp][2 half][3][17]
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
  const int full_rows = P.n_in / ROW;                       // rows the tensor map covers
  // every role walks the same tiles; (stream, tile-in-stream) advances incrementally (one division up front)
  int stream = t_begin / P.tiles_per_stream, ti = t_begin - stream * P.tiles_per_stream;
  auto next_tile = [&]() { if (++ti == P.tiles_per_stream) { ti = 0; ++stream; } };

  // tap table -> shared memory (generic proxy writes, made visible to the tensor core by the fence below)
  for (int i = tid; i < kTcBTileBytes / 16; i += kTcThreads)
    reinterpret_cast<uint4 *>(s_b)[i] = reinterpret_cast<const uint4 *>(P.btab)[i];
  if (tid == 0) {
    for (int i = 0; i < kTcRawStages; ++i) { tc_mbar_init(raw_full(i), 1); tc_mbar_init(raw_empty(i), kTcXformGroupWarps); }
    for (int i = 0; i < kTcAStages; ++i) { tc_mbar_init(a_full(i), kTcXformGroupWarps); tc_mbar_init(a_empty(i), 1); }
    for (int i = 0; i < 2; ++i) { tc_mbar_init(acc_full(i), 1); tc_mbar_init(acc_empty(i), kTcEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  if (warp == 8) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, next_tile()) {
        const int row0 = ti * kTcUseful - kTcHalo;
#ifdef LTB_TC_NO_TMA
        const bool any = false;                                          // bisection build: every row takes the patch path
#else
        const bool any = row0 + kTcTileRows > 0 && row0 < full_rows;     // rows out of range are zero-filled
#endif
        for (int a = 0; a < SPT; ++a, ++it) {
          const int rs = it % kTcRawStages;
          tc_mbar_wait<0>(raw_empty(rs), ((it / kTcRawStages) & 1) ^ 1, P.err, 1);
          if (any) {
            tc_mbar_expect_tx(raw_full(rs), kTcRawBytes);
            tc_tma_load_3d(tc_smem_u32(s_raw + (size_t)rs * kTcRawBytes), &tmap, a * 256, row0, stream, raw_full(rs));
          } else {
            tc_mbar_arrive(raw_full(rs));
          }
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer =====
    int it = 0, tl = 0;
    for (int t = t_begin; t < t_end; ++t, ++tl) {
      const int buf = tl & 1;
      tc_mbar_wait<20>(acc_empty(buf), ((tl >> 1) & 1) ^ 1, P.err, 2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int a = 0; a < SPT; ++a, ++it) {
        const int as = it % kTcAStages;
        tc_mbar_wait<20>(a_full(as), (it / kTcAStages) & 1, P.err, 3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t a_base = tc_smem_u32(s_a + (size_t)as * kTcABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int s = 4 * a + k;                                     // k-step of the row
            const bool first = s == 0;
            const uint64_t adesc = tc_make_desc(a_base + 32 * k);
            if (LTB_TC_BISECT & 2) continue;
            const int off = s * GEO::SPK;                                // first sample of the k-step in its row
            const int u0 = off / D, ph = (off - u0 * D) / GEO::G;        // accumulator column group, tap table
            const uint32_t d = tmem + buf * 256 + 4 * u0;
            const uint64_t bd = tc_make_desc(tc_smem_u32(s_b) + 32 * ph);
            constexpr bool a_signed = FMT != LTB_FMT_FC32;
            tc_mma_i8(d, adesc, bd, first ? tc_idesc(kTcBRows, a_signed) : tc_idesc(GEO::NSTEP, a_signed), first ? 0u : 1u);
          }
          tc_commit(a_empty(as));
          if (a == SPT - 1) tc_commit(acc_full(buf));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 10) {
    // ===== transform: raw interleaved I/Q -> planar re / im rows in the swizzled K-major layout =====
    // two groups of four warps take alternate stages, so two stages are in flight: one group's chain of waits,
    // loads, stores and the proxy fence (~600 cycles) overlaps the other's
    const int grp = (warp - 10) / kTcXformGroupWarps;
    const int tt = tid - 10 * 32 - grp * 32 * kTcXformGroupWarps;        // 0..127
    constexpr int NJ = kTcTileRows * 16 / (32 * kTcXformGroupWarps);     // 64 rows x 16 items per stage / threads of a group
    constexpr int ROWSTEP = 32 * kTcXformGroupWarps / 16;                // rows between a thread's items
    const int r_t = tt >> 4, c = tt & 15;                                // this thread's first row, its 16-byte column
    // item j: row r_t + ROWSTEP j; ROWSTEP is a multiple of 8, so (row & 7) and the swizzled chunk stay fixed
    const int src_off = r_t * 256 + c * 16;
    const int dst_off = tc_sw_off(r_t, c >> 1) + (c & 1) * 8;
    static_assert(ROWSTEP % 8 == 0, "swizzle phase must not change between a thread's items");
    static_assert(kTcXformGroups == 1 || kTcXformGroups == 2, "stages are dealt to the groups by the parity of their running index");
    int it0 = 0;
    for (int t = t_begin; t < t_end; ++t, next_tile(), it0 += SPT) {
      const int row0 = ti * kTcUseful - kTcHalo;
#ifdef LTB_TC_NO_TMA
      const bool patch = true;
#else
      const bool patch = row0 < 0 || row0 + kTcTileRows > full_rows;     // some rows are not plain tensor rows
#endif
      for (int a = 0; a < SPT; ++a) {
        const int it = it0 + a;
        if (kTcXformGroups == 2 && (it & 1) != grp) continue;
        const int rs = it % kTcRawStages, as = it % kTcAStages;
        tc_mbar_wait<0>(raw_full(rs), (it / kTcRawStages) & 1, P.err, 4);
        tc_mbar_wait<0>(a_empty(as), ((it / kTcAStages) & 1) ^ 1, P.err, 5);
        const unsigned char *raw = s_raw + (size_t)rs * kTcRawBytes + src_off;
        unsigned char *dst = s_a + (size_t)as * kTcABytes + dst_off;
        if (LTB_TC_BISECT & 1) {
          __syncwarp();
          if (lane == 0) { tc_mbar_arrive(a_full(as)); tc_mbar_arrive(raw_empty(rs)); }
          continue;
        }
        uint4 w[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) w[j] = *reinterpret_cast<const uint4 *>(raw + j * ROWSTEP * 256);
        if (patch) {
          // first / last tile of a stream's chunk: rows before it come from the carried tail, the partial last
          // row straight from global memory, rows past the end are zeros
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int ra = row0 + r_t + ROWSTEP * j;
            const int n0 = ra * ROW + a * STAGE_SAMPLES + c * ITEM;                // first sample of the item
            if (ra < 0) {
              w[j] = *reinterpret_cast<const uint4 *>((const char *)P.tail + ((size_t)stream * GEO::TAIL + (n0 + GEO::TAIL)) * BPS);
#ifdef LTB_TC_NO_TMA
            } else {
#else
            } else if (ra >= full_rows) {
#endif
              const char *src = (const char *)P.in + (long long)stream * P.stride_bytes;
              unsigned v[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {                                        // 4 bytes at a time
                const int n1 = n0 + q * 4 / BPS;                                   // sc16: one sample, sc8: two, fc32: half
                v[q] = n1 < P.n_in ? *reinterpret_cast<const unsigned *>(src + (size_t)n0 * BPS + 4 * q) : 0u;
              }
              w[j] = make_uint4(v[0], v[1], v[2], v[3]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          uint2 re, im;
          tc_split<FMT>(w[j], re, im, P.q_inv);
          *reinterpret_cast<uint2 *>(dst + j * ROWSTEP * 128) = re;                // rows 0..63: real parts
          *reinterpret_cast<uint2 *>(dst + 64 * 128 + j * ROWSTEP * 128) = im;     // rows 64..127: imaginary parts
        }
        if (!(LTB_TC_BISECT & 8)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) { tc_mbar_arrive(a_full(as)); tc_mbar_arrive(raw_empty(rs)); }
      }
    }
  } else {
    tc_epilogue_role<FMT == LTB_FMT_FC32>(P, tmem, acc_full(0), acc_empty(0), s_stage, s_xchg, t_begin, t_end);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// keep the last three rows (48 D raw samples) of each stream for the next call (the halo rows of its first tile)
template <int FMT, int D>
__global__ void __launch_bounds__(256) tc_tail_kernel(const void *__restrict__ in, long long stride_bytes, int n_in,
                                                      const void *__restrict__ tail_old, void *__restrict__ tail_new) {
  typedef typename std::conditional<FMT == LTB_FMT_FC32, unsigned long long,
                                    typename std::conditional<FMT == LTB_FMT_SC16, unsigned, unsigned short>::type>::type raw_t;   // one complex sample
  constexpr int TAIL = TcGeom<FMT, D>::TAIL;
  const int stream = blockIdx.x;
  const raw_t *src = reinterpret_cast<const raw_t *>((const char *)in + (long long)stream * stride_bytes);
  const raw_t *told = reinterpret_cast<const raw_t *>(tail_old) + (size_t)stream * TAIL;
  raw_t *tnew = reinterpret_cast<raw_t *>(tail_new) + (size_t)stream * TAIL;
  for (int i = threadIdx.x; i < TAIL; i += blockDim.x) {
    const int idx = n_in - TAIL + i;
    tnew[i] = idx >= 0 ? src[idx] : told[TAIL + idx];
  }
}

}  // namespace ltb
