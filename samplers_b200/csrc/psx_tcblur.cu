// psx_tcblur.cu -- K1 of the separable blur on the 5th-generation tensor cores (tcgen05 + TMEM), one launch.
//
// Replaces, for A = V H (separable blur, 256 x 256 planes), the reference lines K1 stands for
// (samplers/networks/base.py:41-43, samplers/inverse_problem.py:17-21, samplers/noise.py:77-79 / 121-123 and
// their autograd mirror samplers/samplers/dps.py:102-103,117-120):
//     cot = (w / sa) * A^T (y - A x0),   x0 = (x_t - s1 eps) / sa,   |r|^2 partial sums.
//
// Formulation.  Each 1-D pass is a banded-Toeplitz product evaluated in K-steps of 16 inputs: the 16 inputs
// k = 16 s .. 16 s + 15 (zero-padded coordinates, PAD = 24) reach the outputs n = 16 s - 48 .. 16 s + 15 only, so
// one K-step is ONE 64-wide UMMA window  D[:, 16 s - 48 .. 16 s + 16) += A[:, 16 s .. 16 s + 16) Bw^T  with the
// SAME 64 x 16 Toeplitz block Bw[n'][k'] = tap[k' - n' + 24 - lo] for every step (clipped at the borders by moving
// the start of the B descriptor and shrinking N).  The image is always the A operand (M = 128 TMEM lanes), the
// accumulators are cleared by one UMMA against a zero block, and the separable factors are applied in the order
//     P1  V   (columns pass)   A = x_t - s1 eps       MN-major, lanes = image columns, K = rows
//     P2  H   (rows pass)      A = V x0               MN-major, lanes = image rows,    K = columns (+ halo)
//     P3  H^T (rows pass)      A = r = y - H V x0     K-major,  lanes = image rows,    K = columns (+ halo)
//     P4  V^T (columns pass)   A = H^T r              MN-major, lanes = image columns, K = rows
// (V and H commute).  The thread that reads lane m of an accumulator always writes 16-byte pieces of the next
// operand, x_t / eps stream in row by row behind the first pass's K loop and cot leaves the last accumulator as
// full 128-byte lines: no transposes, no staging buffers.
//
// Precision.  Operands are fp16 pairs x = hi + lo (22 significant bits); a product is the three UMMAs
// A_hi B_hi + A_lo B_hi + A_hi B_lo accumulated in fp32 in TMEM: 4e-7 relative (profiles/r02_umma_probe.txt).
// fp16 has five exponent bits, so every operand is kept near 1 by a power of two: the taps are scaled into
// [512, 1024), x0 enters as x_t - s1 eps (the 1 / sa is applied to the accumulator), and the residual -- which is
// O(1 / sa): hundreds at the first timesteps, the noise level at the last -- is scaled by 32 * 2^floor(log2 sa).
// Range: |x_t - s1 eps| < 6e4 and |r| < 2e3 / sa.
//
// Work split.  One thread-block CLUSTER of two CTAs per plane; CTA `rank` owns the 128 image columns
// [128 rank, 128 rank + 128) and all 256 rows.  The column passes are local; each row pass needs 24 halo columns
// from the neighbour: 3 contiguous K blocks per row tile, moved by ONE bulk copy (cp.async.bulk shared::cta ->
// shared::cluster) straight into this CTA's operand buffer, completing on this CTA's mbarrier; two more mbarriers
// per exchange that the CTAs arrive on remotely say "your slot is free" and "your copy has been received".
//
// Pipeline.  One operand buffer (176 KB) is rewritten in place A1 -> A2 -> A3 -> A4; the layouts are chosen so that
// the half of the next operand that is complete first never overlaps what the running pass still reads:
//     column-pass operand (A1, A4)  [38 K blocks of 8 padded rows][hi | lo][16 column groups][128 B]      152 KB
//     row-pass operand    (A2, A3)  [2 row tiles of 128][22 K blocks of 8 padded columns][hi | lo][16][128 B]
// and every hand-over is an mbarrier of its own: x_t / eps arrive in 8 chunks of 32 rows (P1 starts on the first),
// P1 commits after the K-step that completes rows 0..127, the row passes run per 128-row tile and commit per 64
// output columns, P4 commits per 64 output rows.  Warps 0-15 load / convert / run the epilogues (each task is
// split over all 16 warps), warp 16 allocates TMEM and issues the UMMAs from one lane, warp 17 issues the
// halo copies, warp 18 sends the "slot is free" signals.
#include <cuda_fp16.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "psx_common.cuh"
#include "psx_tc.cuh"

namespace psx {

using namespace tc;

#ifdef PSX_TRACE
// phase timeline (tools/micro/tc_trace_main.cu): [CTA][role: 0 = epilogue warp 0, 1 = UMMA thread][slot] = %globaltimer
// role 0 = epilogue warp 0, 1 = UMMA thread, 2 = halo-copy thread, 3 = slot-free signal thread
__device__ long long psx_trace_tc[1024 * 4 * 32];
#define PSX_TCTICK(role, slot)                                                             \
  if (warp == ((role) ? kTcWarps + (role) - 1 : 0) && blockIdx.x < 1024 && ((role) ? true : lane == 0)) { \
    long long t_;                                                                          \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                 \
    psx_trace_tc[(blockIdx.x * 4 + (role)) * 32 + (slot)] = t_;                            \
  }
#define PSX_TCTICK_W0(slot)  /* epilogue warp 0 writing into role 3's slots 8.. */                         \
  if (warp == 0 && lane == 0 && blockIdx.x < 1024) {                                          \
    long long t_;                                                                          \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                 \
    psx_trace_tc[(blockIdx.x * 4 + 3) * 32 + (slot)] = t_;                                 \
  }
#define PSX_TCCLOCK(slot)                                                      \
  if (tid == 0 && blockIdx.x < 1024) psx_trace_tc[(blockIdx.x * 4) * 32 + (slot)] = clock64();
#else
#define PSX_TCTICK(role, slot)
#define PSX_TCTICK_W0(slot)
#define PSX_TCCLOCK(slot)
#endif

namespace {

constexpr int kTcN = 256;                  // plane side
constexpr int kTcPad = 24;                 // zero padding / halo width = largest supported tap offset
constexpr int kTcWarps = 16;               // loader / epilogue warps
constexpr int kTcEpi = 32 * kTcWarps;      // loader / epilogue threads
constexpr int kTcThreads = kTcEpi + 96;    // + the UMMA warp, the halo-copy warp, the slot-free signal warp
#ifndef PSX_TC_AHEAD
#define PSX_TC_AHEAD 2
#endif
constexpr int kTcAhead = PSX_TC_AHEAD;     // chunks of the next plane requested ahead of E4 (16 registers each)
constexpr int kHaloBytes = (kTcPad / 8) * 4096;  // 24 halo columns of one row tile: 3 K blocks

constexpr int kKc = 4096;                  // one K block: 8 K values x 128 M values, [hi 2 KB | lo 2 KB]
constexpr int kLo = 2048;                  // offset of the lo part inside a K block
constexpr int kCpKc = (kTcN + 2 * kTcPad) / 8;        // 38 K blocks of the column-pass operand
constexpr int kRpKc = (128 + 2 * kTcPad) / 8;         // 22 K blocks of one row tile of the row-pass operand
constexpr int kRpTile = kRpKc * kKc;                  // 88 KB
constexpr int kOpBytes = 2 * kRpTile;                 // 176 KB (the column-pass operand, 152 KB, fits inside)
constexpr int kBBytes = 4096;              // Toeplitz block: [2 K core columns][hi: 8 N groups | lo: 8 N groups][128 B]
constexpr int kZBytes = 8192;              // zero block (clears up to 256 accumulator columns)
constexpr int kTcSmem = kOpBytes + kBBytes + kZBytes + 1024;
constexpr int kCpSteps = kCpKc / 2;        // 19 K-steps of a column pass (the first and the last are all padding)
constexpr int kRpSteps = kRpKc / 2;        // 11 K-steps of a row pass
static_assert(kCpKc * kKc <= kOpBytes && kTcSmem <= 227 * 1024, "shared memory");

// mbarriers
enum {
  kBLd = 0,      // [8]    chunk c of x_t / eps is in A1                       (16 warps)
  kBImg = 8,     //        Toeplitz block has landed                            (bulk copy)
  kBD1 = 9,      // [2]    P1: rows of tile m are complete                      (commit)
  kBE1 = 11,     // [2]    A2 tile m written locally                            (16 warps)
  kBH1 = 13,     // [2]    A2 tile m: the neighbour's halo columns have landed  (bulk copy from the neighbour)
  kBF1 = 15,     // [2]    the NEIGHBOUR's halo slot of tile m is free (A1 consumed there)   (1 remote thread)
  kBD2 = 17,     // [2][2] P2 tile m: output columns of half h are complete     (commit)
  kBE2 = 21,     // [2]    A3 tile m written locally                            (2 x 16 warps)
  kBH2 = 23,     // [2]    A3 tile m: the neighbour's halo columns have landed  (bulk copy from the neighbour)
  kBF2 = 25,     // [2]    the NEIGHBOUR's halo slot of tile m is free (its P2 is done)      (1 remote thread)
  kBD3 = 27,     // [2]    P3 tile m complete                                   (commit)
  kBE3 = 29,     // [2]    A4 rows of tile m written                            (16 warps)
  kBD4 = 31,     // [4]    P4: output rows of quarter q are complete            (commit)
  kBS1 = 35,     // [2]    the K blocks of A2 tile m that the neighbour needs are written   (4 warps)
  kBS2 = 37,     // [2]    the K blocks of A3 tile m that the neighbour needs are written   (8 warps)
  kBA1 = 39,     // [2]    the neighbour has received this CTA's A2 halo of tile m (source blocks may be rewritten)
  kBA2 = 41,     // [2]    ... A3 halo of tile m                                        (1 remote thread each)
  kBCount = 43
};

// One K-step of a banded pass on NOUT outputs: the window of the 16 inputs 16 S .. 16 S + 15, three split terms.
// Everything but the operand base is a compile-time constant (the issuing thread shares its scheduler with four
// epilogue warps: every instruction it does not execute is tensor-pipe time won).  `a_lo32` / `b_lo32` are the low
// words of the descriptors of the operand base / the Toeplitz block, `hi32` their common high word.
template <int NOUT, int S, int A_MN>
__device__ __forceinline__ void issue_kstep(uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32, uint32_t hi32) {
  constexpr int w0 = 16 * S - 48;
  constexpr int n_lo = w0 > 0 ? w0 : 0;
  constexpr int n_hi = w0 + 64 < NOUT ? w0 + 64 : NOUT;
  constexpr uint32_t id = idesc_f16(128, n_hi - n_lo, A_MN, 0);
  const uint64_t up = (uint64_t)hi32 << 32;
  const uint64_t dAh = up | (a_lo32 + (uint32_t)((2 * S * kKc) >> 4)), dAl = up | (a_lo32 + (uint32_t)((2 * S * kKc + kLo) >> 4));
  const uint64_t dBh = up | (b_lo32 + (uint32_t)((((n_lo - w0) >> 3) * 128) >> 4));
  const uint64_t dBl = up | (b_lo32 + (uint32_t)((((n_lo - w0) >> 3) * 128 + 1024) >> 4));
  umma_f16_acc(d_tmem + n_lo, dAh, dBh, id);
  umma_f16_acc(d_tmem + n_lo, dAl, dBh, id);
  umma_f16_acc(d_tmem + n_lo, dAh, dBl, id);
}
template <int S0, int S1, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (S0 < S1) {
    f(std::integral_constant<int, S0>{});
    static_for<S0 + 1, S1>(f);
  }
}
// D[:, 0 .. n) = 0
__device__ __forceinline__ void issue_clear(uint32_t d_tmem, int n, uint32_t z_addr) {
  umma_f16(d_tmem, smem_desc(z_addr, 2048, 128), smem_desc(z_addr, 4096, 128), idesc_f16(128, n, 0, 0), 0);
}

// 8 fp32 values (times `sc`) -> one hi and one lo 16-byte operand piece
__device__ __forceinline__ void split8(const uint32_t* v, float sc, uint4& hi, uint4& lo) {
  split2(__uint_as_float(v[0]) * sc, __uint_as_float(v[1]) * sc, hi.x, lo.x);
  split2(__uint_as_float(v[2]) * sc, __uint_as_float(v[3]) * sc, hi.y, lo.y);
  split2(__uint_as_float(v[4]) * sc, __uint_as_float(v[5]) * sc, hi.z, lo.z);
  split2(__uint_as_float(v[6]) * sc, __uint_as_float(v[7]) * sc, hi.w, lo.w);
}
__device__ __forceinline__ void st_piece(uint8_t* d, const uint4& hi, const uint4& lo) {
  *reinterpret_cast<uint4*>(d) = hi;
  *reinterpret_cast<uint4*>(d + kLo) = lo;
}
__device__ __forceinline__ void zero_fill(uint8_t* base, int bytes, int tid) {
  for (int i = tid; i < bytes / 16; i += kTcEpi) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
}
// ld_nc8 that stays where it is written (the requests for the next plane must leave BEFORE the waits of E4)
__device__ __forceinline__ void ld_nc8_v(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}
// all lanes have fenced their shared-memory writes; one arrival per warp
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// LOOP = false: one cluster pair per plane (grid = 2 x planes; all of them resident: up to 74 planes), straight-line
// code.  LOOP = true: one resident wave of cluster pairs, each looping over its planes.
//
// MEAN = true (LOOP = false only): the grid carries EXTRA cluster pairs beyond the planes -- the SMs a small batch
// leaves idle (config 2: 96 of 148 SMs hold planes) -- and those CTAs stream x_t and eps once more (second readers:
// L2 hits) and write the bridge mean
//     mean = c_ell x_t + c_s x0            (bridge_kernels.py:41, the roundings of K2: tweedie(), mul, mul, add)
// so that K2 reads ONE array instead of x_t and eps (psx_dps_post_mean): 4 B per element leave the HBM-bound K2 and
// are produced under K1's span by SMs that had nothing to do.  (Writing the mean from the loader warps, which hold
// x_t and eps in registers, was measured first: the stores share the SM <-> L2 path that bounds the load phase --
// K1 20.4 -> 23.7 us for K2 15.1 -> 12.9 us, a net loss; profiles/r02d_inloader_mean_bench_line.json.)
template <bool LOOP, bool MEAN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
    blur_k1_tc(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ y,
               float* __restrict__ cot, float* __restrict__ err_part, const uint8_t* __restrict__ bimg, float inv_scale,
               int C, int64_t obs_repeat, int pp, float sa, float s1, float coef, const float* __restrict__ dsc,
               int planes, float* __restrict__ mean_out, float c_ell, float c_s, int mean_lag_ns,
               const float* __restrict__ zn, float sd) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* op = smem;
  uint8_t* bsm = smem + kOpBytes;
  uint8_t* zsm = bsm + kBBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(zsm + kZBytes);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + kBCount);
  float* red = reinterpret_cast<float*>(bars + kBCount + 1);  // [kTcWarps]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (MEAN && (int)(blockIdx.x >> 1) >= planes) {
    // ================================================================ bridge-mean role (extra cluster pairs; no barriers,
    // no TMEM: both CTAs of such a pair take this branch and leave)
    if (dsc != nullptr) {
      sa = __ldg(dsc);
      s1 = __ldg(dsc + 1);
      c_ell = __ldg(dsc + 3);
      c_s = __ldg(dsc + 4);
      sd = __ldg(dsc + 5);
    }
    const TweedieC tw = make_tc(s1, sa);
    const bool has_z = zn != nullptr;  // the step's noise is known: the stored array is mean + std z (K2's roundings)
    // Tiles of 4096 elements: x_t and eps arrive by bulk copies, the mean leaves through the LSU -- the load / store
    // path of one SM streams ~50 GB/s, the copy engine more than twice that, so the 12 B per element are split 8 / 4
    // between them.  (Measured: everything through the LSU, or the mean stored by bulk copies whose smem reads the
    // producer has to await: 14 us per 725 KB either way.)  full[s]: both arrays of the stage have landed (bytes);
    // done[s]: the 16 warps hold the stage's values in registers.  Warp 18's lane 0 is the producer.
    constexpr int kMT = 4096, kMS = 3, kMA = 3;              // tile elements, stages (3 x 48 KB), arrays per stage
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* done = full + kMS;
    float* stage = reinterpret_cast<float*>(smem + 1024);    // [kMS][x: kMT | eps: kMT | z: kMT]
    static_assert(1024 + kMS * kMA * kMT * 4 <= kTcSmem, "mean-role stages");
    const int64_t tiles = (int64_t)planes * (kTcN * kTcN / kMT);
    const int64_t E = (int64_t)gridDim.x - 2 * planes, e = (int64_t)blockIdx.x - 2 * planes;
    const int nt = (int)((tiles - e + E - 1) / E);           // this CTA's tiles: e, e + E, ...
    if (tid == 0) {
      for (int i = 0; i < kMS; ++i) {
        mbar_init(full + i, 1);
        mbar_init(done + i, kTcWarps);
      }
      fence_mbar_init();
    }
    __syncthreads();
    if (warp == kTcWarps + 2) {
      if (lane == 0) {
        // Tile order: 16-row band t of every plane before band t + 1 of any -- the order in which the plane CTAs
        // stream their chunks in, so that the two readers of a line ask for it at about the same time (one DRAM
        // fetch) instead of competing for different lines.
        auto tile_off = [&](int k) -> int64_t {
          const int g = (int)e + k * (int)E;  // < 16 x 74 tiles
          return (int64_t)((g % planes) * (kTcN * kTcN / kMT) + g / planes) * kMT;
        };
        auto request = [&](int k) {
          const int st_ = k % kMS;
          const int64_t off = tile_off(k);
          mbar_expect_tx(full + st_, (has_z ? 3 : 2) * kMT * 4);
          bulk_g2s(stage + st_ * kMA * kMT, x + off, kMT * 4, full + st_);
          bulk_g2s(stage + st_ * kMA * kMT + kMT, eps + off, kMT * 4, full + st_);
          if (has_z) bulk_g2s(stage + st_ * kMA * kMT + 2 * kMT, zn + off, kMT * 4, full + st_);
        };
        // Start BEHIND most of the plane CTAs' load phase (~8 us; default delay 5 us since the z stage made this role 14.5 us long): that phase is bound by the SM <-> L2 path and gives way to
        // any other traffic (K1 +1.0 us when both start together), the ~10 us of UMMA chain after it leave the memory
        // system idle, and by then x_t / eps are L2 hits.  Measured at L = 16 (profiles/r02i_mean_ab.txt): lag 0 ->
        // K1 21.6 us, 6-10 us -> 20.9-21.1 us (classic 20.7), from 11 us on the mean CTAs (10.5 us) end after the planes.
        if (mean_lag_ns > 0) {
          uint64_t t0, t1;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
          for (int poll = 0; poll < 8192; ++poll) {  // bounded: a timer that does not advance must not hang the launch
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 >= (uint64_t)mean_lag_ns) break;
          }
        }
        for (int k = 0; k < kMS && k < nt; ++k) request(k);
        for (int k = 0; k + kMS < nt; ++k) {
          mbar_wait(done + k % kMS, (uint32_t)(k / kMS) & 1u);  // the 16 warps have read the stage out
          request(k + kMS);
        }
      }
    } else if (warp < kTcWarps) {
      for (int k = 0; k < nt; ++k) {
        const int st_ = k % kMS;
        const float4* xs = reinterpret_cast<const float4*>(stage + st_ * kMA * kMT);
        const float4* es = xs + kMT / 4;
        const float4* zs = es + kMT / 4;
        const int g = (int)e + k * (int)E;
        float* mp = mean_out + (int64_t)((g % planes) * (kTcN * kTcN / kMT) + g / planes) * kMT;
        mbar_wait(full + st_, (uint32_t)(k / kMS) & 1u);
        float4 xv[kMT / 4 / kTcEpi], ev[kMT / 4 / kTcEpi], zv[kMT / 4 / kTcEpi];
#pragma unroll
        for (int u = 0; u < kMT / 4 / kTcEpi; ++u) {
          xv[u] = xs[tid + u * kTcEpi];
          ev[u] = es[tid + u * kTcEpi];
          zv[u] = has_z ? zs[tid + u * kTcEpi] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        warp_arrive(done + st_, lane);  // the values are in registers: the stage may be refilled
#pragma unroll
        for (int u = 0; u < kMT / 4 / kTcEpi; ++u) {
          float4 m;
          m.x = __fadd_rn(__fmul_rn(c_ell, xv[u].x), __fmul_rn(c_s, tweedie(xv[u].x, ev[u].x, tw)));
          m.y = __fadd_rn(__fmul_rn(c_ell, xv[u].y), __fmul_rn(c_s, tweedie(xv[u].y, ev[u].y, tw)));
          m.z = __fadd_rn(__fmul_rn(c_ell, xv[u].z), __fmul_rn(c_s, tweedie(xv[u].z, ev[u].z, tw)));
          m.w = __fadd_rn(__fmul_rn(c_ell, xv[u].w), __fmul_rn(c_s, tweedie(xv[u].w, ev[u].w, tw)));
          if (has_z) {  // bridge_kernels.py:59: mean + std * z, one rounding per operation
            m.x = __fadd_rn(m.x, __fmul_rn(sd, zv[u].x));
            m.y = __fadd_rn(m.y, __fmul_rn(sd, zv[u].y));
            m.z = __fadd_rn(m.z, __fmul_rn(sd, zv[u].z));
            m.w = __fadd_rn(m.w, __fmul_rn(sd, zv[u].w));
          }
          st_stream4(mp + 4 * (tid + u * kTcEpi), m);
        }
      }
    }
    return;
  }
  const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;
  // persistent: cluster c works on planes c, c + (clusters in the grid), ...; every mbarrier completes exactly one
  // phase per plane, so the parity every wait uses is the low bit of the plane iteration `it`
  const int plane0 = blockIdx.x >> 1, pstride = gridDim.x >> 1;
  const int j0 = (int)rank * 128;
  PSX_TCTICK(0, 0)
  PSX_TCTICK(1, 0)
  PSX_TCCLOCK(30)

  float4 xe0, xe1, ee0, ee1;  // LOOP = false: chunk 0 of the plane, requested before the set-up
  if (warp == kTcWarps) {
    if (lane == 0) {
      mbar_init(bars + kBImg, 1);
      fence_mbar_init();
      mbar_expect_tx(bars + kBImg, kBBytes);
      bulk_g2s(bsm, bimg, kBBytes, bars + kBImg);
    }
    __syncwarp();
    tmem_alloc<512>(tslot);
  } else if (warp == kTcWarps + 1) {
    // one barrier per lane (two rounds); arrival counts by barrier index
    for (int i = lane; i < kBCount; i += 32) {
      if (i == kBImg) continue;
      int cnt = 1;  // commits, bulk copies, single remote threads
      if (i < kBLd + 8 || (i >= kBE1 && i < kBE1 + 2) || (i >= kBE3 && i < kBE3 + 2)) cnt = kTcWarps;
      if (i >= kBE2 && i < kBE2 + 2) cnt = 2 * kTcWarps;
      if (i >= kBS1 && i < kBS1 + 2) cnt = 4;
      if (i >= kBS2 && i < kBS2 + 2) cnt = 8;
      mbar_init(bars + i, cnt);
    }
    fence_mbar_init();
    __syncwarp();
    if (lane < 4)  // armed here: the neighbour's bulk copies only add the bytes
      mbar_expect_tx(bars + (lane < 2 ? kBH1 + lane : kBH2 + lane - 2), kHaloBytes);
  } else if (warp < kTcWarps) {
    reinterpret_cast<uint4*>(zsm)[tid] = make_uint4(0, 0, 0, 0);  // 512 x 16 B
    fence_async_smem();
    // Chunk 0 of the (first) plane is requested HERE: its latency then runs under the set-up of the other warps (TMEM
    // allocation, barrier initialisation, Toeplitz block).  For that, nothing the loader warps execute between here and
    // the chunk loop may wait for a load in flight: the proxy fence is above, they skip the tcgen05 fences around the
    // block barrier (their first tcgen05 operation, the TMEM read of E1, follows an mbarrier wait and its own
    // fence::after_thread_sync), and they arrive RELAXED on the cluster barrier (the release belongs to the warps that
    // initialised the mbarriers).  Measured: set-up 0.92 -> 0.53 us, all chunks in 9.15 -> 7.7 us, CTA 19.7 -> 17.9 us.
    // (One chunk only, and only without the plane loop: more early registers spill across the role branches.)
    if (!LOOP) {
      const int r_ = lane & 7, cg_ = 4 * (warp & 3) + (lane >> 3), kb_ = warp >> 2;
      const int64_t off_ = plane0 * (int64_t)(kTcN * kTcN) + (int64_t)(8 * kb_ + r_) * kTcN + j0 + 8 * cg_;
      ld_nc8_v(x + off_, xe0, xe1);
      ld_nc8_v(eps + off_, ee0, ee1);
    }
  }
  if (warp >= kTcWarps) tc_fence_before();
  __syncthreads();
  if (warp >= kTcWarps) {
    tc_fence_after();
    cluster_arrive_release();  // #0: this CTA's barriers are initialised
  } else {
    cluster_arrive_relaxed();
  }
  const uint32_t tb = *tslot;
  PSX_TCTICK(0, 1)
  PSX_TCTICK(1, 1)

  if (warp == kTcWarps) {
    // ================================================================================== UMMA issue (one lane)
    cluster_wait_acquire();  // #0
    if (elect_one()) {
      const uint32_t op_s = smem_u32(op), z_s = smem_u32(zsm);
      const uint64_t da = smem_desc(op_s, kKc, 128), db = smem_desc(smem_u32(bsm), 2048, 128);
      const uint32_t hi32 = (uint32_t)(da >> 32), a0 = (uint32_t)da, b0 = (uint32_t)db;  // same SBO / version word
      const uint32_t cbar = mapa(smem_u32(bars), peer);
      mbar_spin(bars + kBImg, 0);
      uint32_t ph = 0;
#pragma unroll 1
      for (int plane = plane0, it_ = 0; LOOP ? plane < planes : it_ < 1; plane += pstride, ++it_, ph ^= 1u) {
      // ---- P1: V on the own columns, K = rows, streamed in behind the load.  (The accumulator is cleared after the
      // first chunk has been seen: from the second plane on that also says that E3 / E4 of the previous plane have
      // read their accumulators.)
      static_for<1, kCpSteps - 1>([&](auto S) {
        constexpr int s = decltype(S)::value;
        constexpr int last = 16 * s - 9 < kTcN - 1 ? 16 * s - 9 : kTcN - 1;  // last image row of this K-step
        constexpr int prev = s == 1 ? -1 : (16 * (s - 1) - 9 < kTcN - 1 ? 16 * (s - 1) - 9 : kTcN - 1);
        if constexpr ((last >> 5) != (prev >> 5) || s == 1) {
          mbar_spin(bars + kBLd + (last >> 5), ph);
          fence_async_smem();  // the loaders' generic-proxy stores, acquired above -> visible to the UMMA's async proxy
          tc_fence_after();
        }
        if constexpr (s == 1) issue_clear(tb, 256, z_s);
        issue_kstep<256, s, 1>(tb, a0, b0, hi32);
        if constexpr (s == 10) umma_commit(bars + kBD1 + 0);  // outputs n <= 132 are final
      });
      umma_commit(bars + kBD1 + 1);
      PSX_TCTICK(1, 2)
      // ---- P2: H per row tile, K = columns.  The K-steps that need only own columns go first, the two that touch
      // the neighbour's halo last (rank 0: K blocks 19..21 = steps 9, 10, step 0 is border padding; rank 1: K blocks
      // 0..2 = steps 0, 1, step 10 is border padding).  The half of the outputs that is complete first ("th 0") is
      // columns 0..63 on rank 0 (after step 6) and columns 64..127 on rank 1 (after its own steps 2..9).
#pragma unroll 1
      for (int m = 0; m < 2; ++m) {
        mbar_spin(bars + kBE1 + m, ph);
        fence_async_smem();
        tc_fence_after();
        PSX_TCTICK(1, 15 + m)
        const uint32_t d = tb + 256 + 128 * m, am = a0 + (uint32_t)((m * kRpTile) >> 4);
        issue_clear(d, 128, z_s);
        if (rank == 0) {
          static_for<1, 9>([&](auto S) {
            constexpr int s = decltype(S)::value;
            issue_kstep<128, s, 1>(d, am, b0, hi32);
            if constexpr (s == 6) umma_commit(bars + kBD2 + 2 * m);  // outputs n <= 68 are final
          });
        } else {
          static_for<2, 10>([&](auto S) { issue_kstep<128, decltype(S)::value, 1>(d, am, b0, hi32); });
          umma_commit(bars + kBD2 + 2 * m);  // outputs n >= 59 are final
        }
        mbar_spin(bars + kBH1 + m, ph);
        if (LOOP && plane + pstride < planes) mbar_expect_tx(bars + kBH1 + m, kHaloBytes);  // armed for the next plane
        mbar_arrive_remote_relaxed(cbar + 8 * (kBA1 + m));  // the neighbour's copy has been read out of its operand
        tc_fence_after();
        PSX_TCTICK(1, 3 + 2 * m)
        if (rank == 0) {
          issue_kstep<128, 9, 1>(d, am, b0, hi32);
          issue_kstep<128, 10, 1>(d, am, b0, hi32);
        } else {
          issue_kstep<128, 0, 1>(d, am, b0, hi32);
          issue_kstep<128, 1, 1>(d, am, b0, hi32);
        }
        umma_commit(bars + kBD2 + 2 * m + 1);
        PSX_TCTICK(1, 4 + 2 * m)
      }
      // ---- P3: H^T per row tile, the same order
#pragma unroll 1
      for (int m = 0; m < 2; ++m) {
        mbar_spin(bars + kBE2 + m, ph);
        fence_async_smem();
        tc_fence_after();
        PSX_TCTICK(1, 17 + m)
        const uint32_t d = tb + 128 * m, am = a0 + (uint32_t)((m * kRpTile) >> 4);
        issue_clear(d, 128, z_s);
        if (rank == 0) {
          static_for<1, 9>([&](auto S) { issue_kstep<128, decltype(S)::value, 0>(d, am, b0, hi32); });
        } else {
          static_for<2, 10>([&](auto S) { issue_kstep<128, decltype(S)::value, 0>(d, am, b0, hi32); });
        }
        mbar_spin(bars + kBH2 + m, ph);
        if (LOOP && plane + pstride < planes) mbar_expect_tx(bars + kBH2 + m, kHaloBytes);
        mbar_arrive_remote_relaxed(cbar + 8 * (kBA2 + m));
        tc_fence_after();
        PSX_TCTICK(1, 7 + 2 * m)
        if (rank == 0) {
          issue_kstep<128, 9, 0>(d, am, b0, hi32);
          issue_kstep<128, 10, 0>(d, am, b0, hi32);
        } else {
          issue_kstep<128, 0, 0>(d, am, b0, hi32);
          issue_kstep<128, 1, 0>(d, am, b0, hi32);
        }
        umma_commit(bars + kBD3 + m);
        PSX_TCTICK(1, 8 + 2 * m)
      }
      // ---- P4: V^T, K = rows; K-steps 1..7 read rows < 104 (tile 0) and write output rows < 128
      mbar_spin(bars + kBE3 + 0, ph);
      fence_async_smem();
      tc_fence_after();
      PSX_TCTICK(1, 11)
      issue_clear(tb + 256, 128, z_s);
      static_for<1, 8>([&](auto S) {
        constexpr int s = decltype(S)::value;
        issue_kstep<256, s, 1>(tb + 256, a0, b0, hi32);
        if constexpr (s == 6) umma_commit(bars + kBD4 + 0);
      });
      PSX_TCTICK(1, 12)
      mbar_spin(bars + kBE3 + 1, ph);
      fence_async_smem();
      tc_fence_after();
      PSX_TCTICK(1, 13)
      issue_clear(tb + 384, 128, z_s);
      static_for<8, kCpSteps - 1>([&](auto S) {
        constexpr int s = decltype(S)::value;
        issue_kstep<256, s, 1>(tb + 256, a0, b0, hi32);
        if constexpr (s == 10) umma_commit(bars + kBD4 + 1);
        if constexpr (s == 14) umma_commit(bars + kBD4 + 2);
      });
      umma_commit(bars + kBD4 + 3);
      PSX_TCTICK(1, 14)
      }  // planes
    }
    __syncwarp();
  } else if (warp == kTcWarps + 1) {
    // ================================================================================== halo copies to the neighbour
    // The 24 own columns next to the neighbour are 3 contiguous K blocks of a row tile, here and there: one bulk copy
    // through distributed shared memory per tile and exchange, completing on the neighbour's H barrier.
    cluster_wait_acquire();  // #0: the neighbour's barriers exist
    if (elect_one()) {
      const uint32_t src = (rank == 0 ? 16 : 3) * kKc, dst = (rank == 0 ? 0 : kRpKc - 3) * kKc;
      const uint32_t cop = mapa(smem_u32(op), peer), cbar = mapa(smem_u32(bars), peer);
      uint32_t ph = 0;
#pragma unroll 1
      for (int plane = plane0, it_ = 0; LOOP ? plane < planes : it_ < 1; plane += pstride, ++it_, ph ^= 1u) {
        for (int m = 0; m < 2; ++m) {
          mbar_spin(bars + kBS1 + m, ph);
          fence_async_smem();  // the source blocks were written through the generic proxy, the bulk copy reads through the async one
          PSX_TCTICK(2, 5 + m)
          mbar_spin(bars + kBF1 + m, ph);
          PSX_TCTICK(2, 1 + m)
          bulk_s2peer(cop + m * kRpTile + dst, op + m * kRpTile + src, kHaloBytes, cbar + 8 * (kBH1 + m));
        }
        for (int m = 0; m < 2; ++m) {
          mbar_spin(bars + kBS2 + m, ph);
          fence_async_smem();
          PSX_TCTICK(2, 7 + m)
          mbar_spin(bars + kBF2 + m, ph);
          PSX_TCTICK(2, 3 + m)
          bulk_s2peer(cop + m * kRpTile + dst, op + m * kRpTile + src, kHaloBytes, cbar + 8 * (kBH2 + m));
        }
      }
    }
    __syncwarp();
  } else if (warp == kTcWarps + 2) {
    // ================================================================================== "slot is free" signals
    cluster_wait_acquire();  // #0
    if (elect_one()) {
      const uint32_t f1 = mapa(smem_u32(bars + kBF1), peer), f2 = mapa(smem_u32(bars + kBF2), peer);
      uint32_t ph = 0;
#pragma unroll 1
      for (int plane = plane0, it_ = 0; LOOP ? plane < planes : it_ < 1; plane += pstride, ++it_, ph ^= 1u) {
        // this CTA's halo slot of tile m overlaps A1 rows that P1 reads up to the commit of tile m
        mbar_spin(bars + kBD1 + 0, ph);
        mbar_arrive_remote_relaxed(f1);
        PSX_TCTICK(3, 1)
        mbar_spin(bars + kBD1 + 1, ph);
        mbar_arrive_remote_relaxed(f1 + 8);
        PSX_TCTICK(3, 2)
        // ... and holds the A2 halo until P2 of tile m is done
        mbar_spin(bars + kBD2 + 1, ph);
        mbar_arrive_remote_relaxed(f2);
        PSX_TCTICK(3, 3)
        mbar_spin(bars + kBD2 + 3, ph);
        mbar_arrive_remote_relaxed(f2 + 8);
        PSX_TCTICK(3, 4)
      }
    }
    __syncwarp();
  } else {
    // ================================================================================== load / epilogues
    // warp (q, cq): TMEM lanes 32 q .. 32 q + 31, a cq-th share of the accumulator columns of every task
    const int q = warp & 3, cq = warp >> 2;
    const uint32_t tlane = tb + ((uint32_t)(32 * q) << 16);
    const int ml = 32 * q + lane;  // this thread's TMEM lane
    // chunk c = rows 32 c .. 32 c + 31 = K blocks 3 + 4 c ..; warp = (K block kb, column-group quad), lane = (row r,
    // column group cl): 32 bytes of one row per lane, 8 rows x 128 B per warp instruction
    const int ld_r = lane & 7, ld_cg = 4 * (warp & 3) + (lane >> 3), ld_kb = warp >> 2;
    const int64_t ld_off = (int64_t)(8 * ld_kb + ld_r) * kTcN + j0 + 8 * ld_cg;
    constexpr int kAhead = LOOP ? kTcAhead : 4;
    float4 xa[4], xb[4], ea[4], eb[4];
    // the first chunks of the first plane (those of every further plane are requested ahead of E4 of the plane before)
    if (!LOOP) {
      xa[0] = xe0; xb[0] = xe1; ea[0] = ee0; eb[0] = ee1;
    }
#pragma unroll
    for (int u = LOOP ? 0 : 1; u < kAhead; ++u) {
      ld_nc8_v(x + plane0 * (int64_t)(kTcN * kTcN) + ld_off + u * 32 * kTcN, xa[u], xb[u]);
      ld_nc8_v(eps + plane0 * (int64_t)(kTcN * kTcN) + ld_off + u * 32 * kTcN, ea[u], eb[u]);
    }
    uint32_t ph = 0;
#pragma unroll 1
    for (int plane = plane0, it_ = 0; LOOP ? plane < planes : it_ < 1; plane += pstride, ++it_, ph ^= 1u) {
    const float* xp = x + plane * (int64_t)(kTcN * kTcN) + ld_off;
    const float* ep = eps + plane * (int64_t)(kTcN * kTcN) + ld_off;
    // chunks 0 .. kTcAhead - 1 were requested ahead of E4 of the plane before (register budget); the rest now, before
    // anything that waits (the step scalars are first used by the conversions below)
#pragma unroll
    for (int u = kAhead; u < 4; ++u) {
      ld_nc8_v(xp + u * 32 * kTcN, xa[u], xb[u]);
      ld_nc8_v(ep + u * 32 * kTcN, ea[u], eb[u]);
    }
    // graph replay: the step scalars are a device row, requested behind the chunks (re-read per plane: three words
    // that hit L1 / L2, instead of three registers -- or stack slots, whose store would wait for the load -- kept
    // alive across the whole loop)
    if (dsc != nullptr) {
      sa = __ldg(dsc);
      s1 = __ldg(dsc + 1);
    }
    const int64_t l = plane / C, ch = plane % C;
    const float* yplane = y + ((l / obs_repeat) * C + ch) * (int64_t)(kTcN * kTcN) + j0;
    // the observation is read much later: pull this CTA's half plane towards L2 now
    for (int i = tid; i < 256 * 4; i += kTcEpi)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(yplane + (i >> 2) * kTcN + (i & 3) * 32));

    // ------------------------------------------------------------------------------------ load: A1 = x_t - s1 eps
    {
      uint8_t* d0 = op + (3 + ld_kb) * kKc + ld_cg * 128 + ld_r * 16;
      // rows above / below the image: K blocks 0..2 and 35..37 of A1 (covered by the arrival on chunk 0)
      zero_fill(op, 3 * kKc, tid);
      zero_fill(op + 35 * kKc, 3 * kKc, tid);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int u = c & 3;
        uint4 hi, lo;
        split2(fmaf(-s1, ea[u].x, xa[u].x), fmaf(-s1, ea[u].y, xa[u].y), hi.x, lo.x);
        split2(fmaf(-s1, ea[u].z, xa[u].z), fmaf(-s1, ea[u].w, xa[u].w), hi.y, lo.y);
        split2(fmaf(-s1, eb[u].x, xb[u].x), fmaf(-s1, eb[u].y, xb[u].y), hi.z, lo.z);
        split2(fmaf(-s1, eb[u].z, xb[u].z), fmaf(-s1, eb[u].w, xb[u].w), hi.w, lo.w);
        if (c + 4 < 8) {
          ld_nc8(xp + (c + 4) * 32 * kTcN, xa[u], xb[u]);
          ld_nc8(ep + (c + 4) * 32 * kTcN, ea[u], eb[u]);
        }
        st_piece(d0 + c * 4 * kKc, hi, lo);
        // no proxy fence here: it would wait for this thread's loads in flight and serialise the prefetch; the UMMA
        // thread fences after it has acquired the chunk's barrier (measured: 0.6 us of the load phase)
        warp_arrive(bars + kBLd + c, lane);
      }
    }
    PSX_TCTICK(0, 2)
    // residual scale 32 * 2^floor(log2 sa) (sa is a positive normal number: sqrt of a clipped alpha-bar)
    const float rscale = __int_as_float((__float_as_int(sa) & 0x7f800000) + (5 << 23));
    if (plane == plane0) cluster_wait_acquire();  // #0

    // the observation values of the first E2 task: requested now, needed ~2 us from here (nothing in between waits
    // for this thread's loads: the proxy fences are on the consumers' side)
    const int hfirst = rank == 0 ? 0 : 1;
    const float* yrow = yplane + (int64_t)ml * kTcN + 16 * cq;
    float4 ya, yb, yc, yd;
    ld_nc8(yrow + 64 * hfirst, ya, yb);
    ld_nc8(yrow + 64 * hfirst + 8, yc, yd);

    // ------------------------------------------------------------------------------------ E1: V x0 -> A2 (MN-major)
    // lane = own column jl, accumulator columns = image rows; tile m = rows 128 m .. 128 m + 127
    {
      const int jl = ml;
      const bool halo_warp = rank == 0 ? q == 3 : q == 0;  // owns the columns the neighbour needs
#pragma unroll 1
      for (int m = 0; m < 2; ++m) {
        mbar_wait(bars + kBD1 + m, ph);
        tc_fence_after();
        PSX_TCTICK(0, 3 + 2 * m)
        // TMEM is read 8 columns at a time, the next chunk in flight while this one is split and stored (TMEM read,
        // ALU and shared-memory stores are three different pipes: a task costs the longest, not the sum)
        uint32_t va[8], vb[8];
        const uint32_t tsrc = tlane + 128 * m + 32 * cq;
        tmem_ld8(tsrc, va);
        // border side of the row-pass operand: K blocks 0..2 (rank 0) / 19..21 (rank 1) of tile m
        zero_fill(op + m * kRpTile + (rank == 0 ? 0 : kRpKc - 3) * kKc, 3 * kKc, tid);
        uint8_t* d = op + m * kRpTile + (3 + (jl >> 3)) * kKc + (4 * cq) * 128 + (jl & 7) * 16;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          tmem_ld_wait();
          if (g < 3) tmem_ld8(tsrc + 8 * (g + 1), (g & 1) ? va : vb);
          uint4 hi, lo;
          split8((g & 1) ? vb : va, inv_scale, hi, lo);
          st_piece(d + g * 128, hi, lo);
        }
        tc_fence_before();
        warp_arrive(bars + kBE1 + m, lane);  // (proxy fence: on the consumers' side, see the chunk loop)
        if (halo_warp && lane == 0) mbar_arrive(bars + kBS1 + m);
        PSX_TCTICK(0, 4 + 2 * m)
      }
    }

    // ------------------------------------------------------------------------------------ E2: r = y - H V x0 -> A3 (K-major)
    // lane = image row il of tile m, accumulator columns = own columns; task (m, h): columns 64 h + 16 cq .. + 15
    float acc = 0.f;
    {
      // task t = (tile m, th): th 0 = the half of the columns whose outputs are complete first (see P2)
      const float sc = inv_scale / sa;
#pragma unroll 1
      for (int t = 0; t < 4; ++t) {
        const int m = t >> 1, th = t & 1, h = th ^ hfirst, n0 = 64 * h + 16 * cq;
        const float yy[16] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w, yc.x, yc.y, yc.z, yc.w, yd.x, yd.y, yd.z, yd.w};
        if (t < 3) {  // the next task's observation values: in flight across this task (no fence in it waits for them)
          const float* yn = yrow + (int64_t)(128 * ((t + 1) >> 1)) * kTcN + 64 * (((t + 1) & 1) ^ hfirst);
          ld_nc8(yn, ya, yb);
          ld_nc8(yn + 8, yc, yd);
        }
        mbar_wait(bars + kBD2 + t, ph);
        tc_fence_after();
        PSX_TCTICK(0, 7 + 2 * t)
        uint32_t va[8], vb[8];
        const uint32_t tsrc = tlane + 256 + 128 * m + n0;
        tmem_ld8(tsrc, va);
        uint8_t* d = op + m * kRpTile + (3 + (n0 >> 3)) * kKc + (ml >> 3) * 128 + (ml & 7) * 16;
        // the columns next to the neighbour are in the late half on both ranks
        const bool halo_warp = th == 1 && (rank == 0 ? cq >= 2 : cq <= 1);
        if (halo_warp) mbar_wait_cluster(bars + kBA1 + m, ph);  // (the copy left long ago; this only makes it formal)
        tmem_ld_wait();
        tmem_ld8(tsrc + 8, vb);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g == 1) tmem_ld_wait();
          uint32_t rr[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float rv = yy[8 * g + e] - __uint_as_float(g ? vb[e] : va[e]) * sc;
            acc = fmaf(rv, rv, acc);
            rr[e] = __float_as_uint(rv);
          }
          uint4 hi, lo;
          split8(rr, rscale, hi, lo);
          st_piece(d + g * kKc, hi, lo);
        }
        tc_fence_before();
        warp_arrive(bars + kBE2 + m, lane);
        if (halo_warp && lane == 0) mbar_arrive(bars + kBS2 + m);
        PSX_TCTICK(0, 8 + 2 * t)
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;

    // ------------------------------------------------------------------------------------ E3: H^T r -> A4 (MN-major)
    // lane = image row il of tile m, accumulator columns = own columns 32 cq .. 32 cq + 31 (= M of the next pass)
#pragma unroll 1
    for (int m = 0; m < 2; ++m) {
      mbar_wait(bars + kBD3 + m, ph);
      tc_fence_after();
      PSX_TCTICK(0, 15 + 2 * m)
      uint32_t va[8], vb[8];
      const uint32_t tsrc = tlane + 128 * m + 32 * cq;
      tmem_ld8(tsrc, va);
      // rows above (tile 0) / below (tile 1) the image: K blocks 0..2 / 35..37 of A4
      zero_fill(op + (m == 0 ? 0 : 35) * kKc, 3 * kKc, tid);
      mbar_wait_cluster(bars + kBA2 + m, ph);  // this CTA's A3 halo copy of tile m has left its source blocks
      const int kp = 128 * m + ml + kTcPad;
      uint8_t* d = op + (kp >> 3) * kKc + (4 * cq) * 128 + (kp & 7) * 16;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        tmem_ld_wait();
        if (g < 3) tmem_ld8(tsrc + 8 * (g + 1), (g & 1) ? va : vb);
        uint4 hi, lo;
        split8((g & 1) ? vb : va, inv_scale, hi, lo);
        st_piece(d + g * 128, hi, lo);
      }
      tc_fence_before();
      warp_arrive(bars + kBE3 + m, lane);
      PSX_TCTICK(0, 16 + 2 * m)
    }

    // |r|^2 of this CTA's half plane -> its partial-sum slots (the epilogue warps only)
    asm volatile("bar.sync 1, %0;" ::"n"(kTcEpi) : "memory");
    if (tid < pp / 2) {
      float tot = 0.f;
      if (tid == 0)
        for (int i = 0; i < kTcWarps; ++i) tot += red[i];
      err_part[plane * pp + rank * (pp / 2) + tid] = tot;
    }

    // the next plane's first four chunks: requested here, in flight across E4 (registers only; A1 itself is written
    // after E4 has seen P4's last commit, i.e. after the tensor cores have read A4 out of the operand buffer)
    if (LOOP && plane + pstride < planes) {
      const float* xn = x + (plane + pstride) * (int64_t)(kTcN * kTcN) + ld_off;
      const float* en = eps + (plane + pstride) * (int64_t)(kTcN * kTcN) + ld_off;
#pragma unroll
      for (int u = 0; u < kAhead; ++u) {
        ld_nc8_v(xn + u * 32 * kTcN, xa[u], xb[u]);
        ld_nc8_v(en + u * 32 * kTcN, ea[u], eb[u]);
      }
    }

    // ------------------------------------------------------------------------------------ E4: cot
    // lane = own column, accumulator columns = image rows: every store instruction writes one 128-byte line
    {
      step_scalars_coef(dsc, coef);  // (first needed here: not kept alive across the plane)
      const float sc = inv_scale * coef / rscale;
      float* cp = cot + plane * (int64_t)(kTcN * kTcN) + j0 + ml;
#pragma unroll 1
      for (int qq = 0; qq < 4; ++qq) {
        mbar_wait(bars + kBD4 + qq, ph);
        tc_fence_after();
        PSX_TCTICK(0, 19 + 2 * qq)
        const int i0 = 64 * qq + 16 * cq;
        uint32_t va[8], vb[8];
        tmem_ld8(tlane + 256 + i0, va);
        tmem_ld_wait();
        tmem_ld8(tlane + 256 + i0 + 8, vb);
#pragma unroll
        for (int e = 0; e < 8; ++e) __stcs(cp + (i0 + e) * kTcN, __uint_as_float(va[e]) * sc);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e) __stcs(cp + (i0 + 8 + e) * kTcN, __uint_as_float(vb[e]) * sc);
        PSX_TCTICK(0, 20 + 2 * qq)
      }
    }
    }  // planes
  }
  // Every bulk copy into this CTA and every remote arrival on its barriers is awaited by one of its own waits above
  // (H1 / H2 before P2 / P3, F1 / F2 before the copies, A1 / A2 before the copied blocks are rewritten), and the A
  // acknowledgements also say that this CTA's own outgoing copies have been read out of its shared memory: nothing
  // of or for the neighbour can still be in flight here.
  tc_fence_before();
  __syncthreads();
  PSX_TCTICK(0, 27)
  PSX_TCCLOCK(31)
  if (warp == kTcWarps) tmem_dealloc<512>(tb);
}

// ------------------------------------------------------------------------------------------ host: Toeplitz block image
// Bw[n'][k'] = scale * tap[k' - n' + PAD - lo] for the pass whose 1-D operation is out[p] = sum_i w[i] in[p + lo + i]
// (psx::Taps), as fp16 hi / lo in the shared-memory image [K core column (2)][hi: 8 N groups | lo: 8 N groups][8 x 8].
bool taps_extent(const Taps& t, int& lo_off, int& hi_off) {
  int a = -1, b = -1;
  for (int i = 0; i < t.k; ++i)
    if (t.ww[i].x != 0.f) {
      if (a < 0) a = i;
      b = i;
    }
  if (a < 0) return false;
  lo_off = t.lo + a;
  hi_off = t.lo + b;
  return true;
}

void build_image(const Taps& t, float scale, std::vector<uint8_t>& img) {
  img.assign(kBBytes, 0);
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 16; ++k) {
      const int ti = k - n + kTcPad - t.lo;
      const float w = (ti >= 0 && ti < t.k) ? t.ww[ti].x * scale : 0.f;
      const __half h = __float2half_rn(w);
      const __half l = __float2half_rn(w - __half2float(h));
      const size_t core = (size_t)(k / 8) * 2048 + (size_t)(n / 8) * 128 + (size_t)(n % 8) * 16 + (size_t)(k % 8) * 2;
      std::memcpy(&img[core], &h, 2);
      std::memcpy(&img[core + 1024], &l, 2);
    }
}

}  // namespace

// Decides whether the tensor-core K1 applies to this operator and uploads its Toeplitz block.
void tcblur_plan(psx_op* op) {
  op->tc_pad = 0;
  op->d_tc_img = nullptr;
  if (op->H != kTcN || op->W != kTcN) return;
  const Taps* ts[4] = {&op->fh, &op->fv, &op->av, &op->ah};
  int need = 0;
  float mx = 0.f;
  for (const Taps* t : ts) {
    int a, b;
    if (!taps_extent(*t, a, b)) return;
    need = std::max(need, std::max(-a, b));
    for (int i = 0; i < t->k; ++i) mx = std::fmax(mx, std::fabs(t->ww[i].x));
  }
  if (need > kTcPad) return;  // the 64-wide window of a K-step covers tap offsets up to +-24 (61-tap sigma = 3: 19)
  if (!(mx > 0.f) || !std::isfinite(mx)) return;
  int e = 0;
  std::frexp(mx, &e);                           // mx = f * 2^e, f in [0.5, 1)
  const float scale = std::ldexp(1.f, 10 - e);  // largest tap -> [512, 1024)
  std::vector<uint8_t> img[4];
  for (int p = 0; p < 4; ++p) build_image(*ts[p], scale, img[p]);
  for (int p = 1; p < 4; ++p)
    if (img[p] != img[0]) return;  // one resident block: symmetric taps, the same for rows and columns
  void* d = nullptr;
  if (cudaMalloc(&d, img[0].size()) != cudaSuccess ||
      cudaMemcpy(d, img[0].data(), img[0].size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    if (d) cudaFree(d);
    return;
  }
  op->d_tc_img = (uint8_t*)d;
  op->tc_pad = kTcPad;
  op->tc_inv_scale = 1.f / scale;
}

void tcblur_release(psx_op* op) {
  if (op->d_tc_img) cudaFree(op->d_tc_img);
  op->d_tc_img = nullptr;
  op->tc_pad = 0;
}

bool tcblur_available(const psx_op* op) { return op->tc_pad != 0 && op->err_parts % (2 * op->C) == 0; }

// Cluster pairs of blur_k1_tc that are resident at once on the current device (74 on a 148-SM B200), per device.
static int tc_resident_clusters() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * 148);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = kTcSmem;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, blur_k1_tc<true, false>, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = sm_count() / 2;
    }
    cached[dev] = n;
  }
  return cached[dev];
}

// The bridge-mean role needs idle SMs: the pairs the planes leave free must stream 12 B per element (x_t and eps in,
// the mean out) within K1's span -- measured ~50 GB/s per SM against ~20 us: at most 2.2 planes per free pair.
bool tcblur_mean_fits(const psx_op* op, int64_t L) {
  const int64_t planes = L * op->C, resident = tc_resident_clusters();
  return planes < resident && 10 * planes <= 22 * (resident - planes);
}

int launch_pre_sepblur_tc(const psx_op* op, const float* x, const float* eps, const float* y, int64_t L,
                          int64_t obs_repeat, float sa, float s1, float w, const float* dsc, float* cot,
                          float* err_part, float* mean_out, float c_ell, float c_s, const float* zn, float sd,
                          cudaStream_t st) {
  static bool attr[64] = {};  // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!attr[dev]) {
    const void* fns[3] = {(const void*)blur_k1_tc<false, false>, (const void*)blur_k1_tc<true, false>,
                          (const void*)blur_k1_tc<false, true>};
    for (const void* f : fns)
      if (int rc = check_cuda(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem),
                              "blur_k1_tc attribute"))
        return rc;
    attr[dev] = true;
  }
  const int64_t planes = L * op->C;
  PSX_REQUIRE(planes <= INT32_MAX / 2, "blur_k1_tc: too many planes");
  const float coef = (float)((double)w / (double)sa);
  // A cluster pair per plane, in as many waves as it takes.  PSX_TC_PERSIST=1 selects, beyond one resident wave (74
  // pairs on a 148-SM part), ONE wave of persistent pairs that loop over their planes: it was 3 % faster at L = 64
  // until chunk 0 moved ahead of the set-up, which the loop kernel cannot afford in registers (62.9 against 65.1 us).
  const int64_t resident = tc_resident_clusters();
  const bool loop = planes > resident && env_opts().tc_persist;
  PSX_REQUIRE(!mean_out || tcblur_mean_fits(op, L), "blur_k1_tc: no idle SMs for the bridge-mean role at this batch");
  // with the mean: every pair the planes leave free carries the bridge-mean role
  const unsigned grid = (unsigned)((loop ? resident : (mean_out ? resident : planes)) * 2);
#define PSX_TC_LAUNCH(LOOP_, MEAN_)                                                                                  \
  blur_k1_tc<LOOP_, MEAN_><<<grid, kTcThreads, kTcSmem, st>>>(x, eps, y, cot, err_part, op->d_tc_img, op->tc_inv_scale, \
                                                              op->C, obs_repeat, op->err_parts / op->C, sa, s1, coef, \
                                                              dsc, (int)planes, mean_out, c_ell, c_s, env_opts().mean_lag_ns, zn, sd)
  if (mean_out) PSX_TC_LAUNCH(false, true);
  else if (loop) PSX_TC_LAUNCH(true, false);
  else PSX_TC_LAUNCH(false, false);
#undef PSX_TC_LAUNCH
  return check_cuda(cudaGetLastError(), "blur_k1_tc launch");
}

}  // namespace psx
