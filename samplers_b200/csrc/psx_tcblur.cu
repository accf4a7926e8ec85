// psx_tcblur.cu -- K1 of the separable blur on the 5th-generation tensor cores (tcgen05 + TMEM), one launch.
//
// Replaces, for A = V H (separable blur, 256 x 256 planes), the reference lines K1 stands for
// (samplers/networks/base.py:41-43, samplers/inverse_problem.py:17-21, samplers/noise.py:77-79 / 121-123 and
// their autograd mirror samplers/samplers/dps.py:102-103,117-120):
//     cot = (w / sa) * H^T V^T (y - V H x0),   x0 = (x_t - s1 eps) / sa,   |r|^2 partial sums.
//
// Formulation.  Each of the four 1-D passes is a banded-Toeplitz matrix product, evaluated as N-tiles of 64 outputs
// whose K window is the 64 + 2 PAD inputs the band touches.  The Toeplitz block B[n][k] = tap[(k - PAD) - n - lo] is
// the same for every tile (the operands are zero-padded by PAD, so the image border needs no special case) and is
// always the B operand of the UMMA; the image data is always the A operand (M = 128 TMEM lanes):
//     P1  H1 [i , j'] = sum_j  X [i , j ] Bh [j', j ]      A = X   K-major   (lanes = image rows)
//     P2  AXt[j', i'] = sum_i  H1[i , j'] Bv [i', i ]      A = H1  MN-major  (lanes = image columns)
//     P3  H2t[j', i ] = sum_i' R [i', j'] Bvt[i , i']      A = Rt  K-major   (lanes = image columns)
//     P4  cot[i , j ] = sum_j' H2[i , j'] Bht[j , j']      A = H2  MN-major  (lanes = image rows)
// so that the thread that reads row m of an accumulator from TMEM always writes 16-byte pieces of the next operand
// (its values run along K for a K-major operand and along M for an MN-major one): no transposes, and eight
// consecutive lanes fill one contiguous 128-byte core matrix (conflict-free shared-memory stores).
//
// Precision.  Operands are fp16 pairs x = hi + lo (22 significant bits); a product is the three UMMAs
// A_hi B_hi + A_lo B_hi + A_hi B_lo accumulated in fp32 in TMEM: measured 4e-7 relative (profiles/r02_umma_probe.txt).
// The taps are scaled by a power of two into fp16's normal range, x0 enters as x_t - s1 eps (the 1/sa is applied to the
// accumulator) and the residual is scaled by 64, so that fp16's range is not an issue for |x_t - s1 eps| < 6e4.
//
// Work split.  One thread-block CLUSTER of two CTAs per plane; CTA `rank` owns the 128 image columns
// [128 rank, 128 rank + 128) and all 256 rows.  The column passes P2 / P3 are local to a column range; P1 reads its
// PAD halo columns of x_t / eps straight from global memory; only P4 needs the neighbour's H2 halo columns, which the
// neighbour's P3 epilogue writes into this CTA's operand buffer through distributed shared memory (one exchange,
// two cluster barriers, both split into arrive / wait so that nobody blocks on them in the common case).
//
// Shared memory: one operand buffer (A1 -> A2 -> A3 -> A4 in place, 176 KB for PAD = 24) + the Toeplitz block
// (28 KB); TMEM: 2 x 256 columns, alternating between passes.  Warps 0-15 load / convert / run the epilogues,
// warp 16 allocates TMEM, fetches the Toeplitz block with a bulk copy and issues the UMMAs from one lane.
#include <cuda_fp16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "psx_common.cuh"
#include "psx_tc.cuh"

namespace psx {

using namespace tc;

#ifdef PSX_TRACE
// phase timeline (tools/micro/tc_trace_main.cu): [CTA][role: 0 = epilogue warp 0, 1 = UMMA warp][slot] = %globaltimer
__device__ long long psx_trace_tc[1024 * 2 * 16];
#define PSX_TCTICK(role, slot)                                                             \
  if (lane == 0 && warp == ((role) ? 16 : 0) && blockIdx.x < 1024) {                        \
    long long t_;                                                                          \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                 \
    psx_trace_tc[(blockIdx.x * 2 + (role)) * 16 + (slot)] = t_;                            \
  }
#else
#define PSX_TCTICK(role, slot)
#endif

namespace {

constexpr int kTcN = 256;        // plane side
constexpr int kTcNT = 64;        // outputs per N-tile
constexpr int kTcWarps = 16;               // loader / epilogue warps
constexpr int kTcEpi = 32 * kTcWarps;      // loader / epilogue threads
constexpr int kTcThreads = kTcEpi + 32;    // + the UMMA warp
constexpr float kTcRScale = 64.f;

template <int PAD>
struct TcGeo {
  static constexpr int KW = kTcNT + 2 * PAD;        // K window of one N-tile
  static constexpr int KS = KW / 16;                // UMMA k-steps per tile
  static constexpr int KC_ROW = (128 + 2 * PAD) / 8;  // K core columns of A1 / A4 (own columns + halo)
  static constexpr int KC_COL = (256 + 2 * PAD) / 8;  // K core columns of A2 / A3 (all rows + zero pad)
  // operand layout: [K core column][part: hi, lo][M group][128-byte core matrix]
  static constexpr int LBO_ROW = 2 * 32 * 128;      // A1 / A4: 256 rows = 32 M groups
  static constexpr int LBO_COL = 2 * 16 * 128;      // A2 / A3: 128 columns = 16 M groups
  static constexpr int LO_ROW = 32 * 128, LO_COL = 16 * 128;  // offset of the lo part
  static constexpr int OP_BYTES = KC_ROW * LBO_ROW;
  static constexpr int COL_BYTES = KC_COL * LBO_COL;
  static constexpr int HALO_BYTES = (PAD / 8) * LBO_ROW;  // the K core columns of A4 the neighbour writes
  static constexpr int B_BYTES = (KW / 8) * 2048;   // [K core column][16 N groups: 8 hi + 8 lo][128]
  static constexpr int SMEM = OP_BYTES + B_BYTES + 256;
  static_assert(PAD % 8 == 0 && KW % 16 == 0, "PAD must be a multiple of 8");
  static_assert(OP_BYTES - COL_BYTES == HALO_BYTES, "A2 / A3 must fit beside the halo slot");
  static_assert(SMEM <= 227 * 1024, "shared memory");
};

// three UMMAs per k-step: D (+)= Ah Bh + Al Bh + Ah Bl
__device__ __forceinline__ void issue_tile(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo_off, uint32_t lbo_a,
                                           uint64_t dBh, uint64_t dBl, uint32_t idesc, int ksteps) {
  const uint64_t dAh = smem_desc(a_hi, lbo_a, 128);
  const uint64_t dAl = dAh + (uint64_t)(a_lo_off >> 4);
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint64_t aa = (uint64_t)((2u * lbo_a * ks) >> 4), ab = (uint64_t)((2u * 2048u * ks) >> 4);
    umma_f16(d_tmem, dAh + aa, dBh + ab, idesc, ks > 0);
    umma_f16(d_tmem, dAl + aa, dBh + ab, idesc, 1);
    umma_f16(d_tmem, dAh + aa, dBl + ab, idesc, 1);
  }
}

// 32 fp32 accumulator values (times `sc`) -> four hi and four lo 16-byte operand pieces
__device__ __forceinline__ void split32(const uint32_t (&v)[32], float sc, uint4 (&hi)[4], uint4 (&lo)[4]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    split2(__uint_as_float(v[8 * g + 0]) * sc, __uint_as_float(v[8 * g + 1]) * sc, hi[g].x, lo[g].x);
    split2(__uint_as_float(v[8 * g + 2]) * sc, __uint_as_float(v[8 * g + 3]) * sc, hi[g].y, lo[g].y);
    split2(__uint_as_float(v[8 * g + 4]) * sc, __uint_as_float(v[8 * g + 5]) * sc, hi[g].z, lo[g].z);
    split2(__uint_as_float(v[8 * g + 6]) * sc, __uint_as_float(v[8 * g + 7]) * sc, hi[g].w, lo[g].w);
  }
}

template <int PAD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
    blur_k1_tc(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ y,
               float* __restrict__ cot, float* __restrict__ err_part, const uint8_t* __restrict__ bimg, float4 bsc,
               int C, int64_t obs_repeat, int pp, float sa, float s1, float coef, const float* __restrict__ dsc) {
  using G = TcGeo<PAD>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* op = smem;
  uint8_t* bsm = smem + G::OP_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(bsm + G::B_BYTES);  // [4] operand of pass p is in shared memory
  uint64_t* done = full + 4;                                        // [4] UMMAs of pass p have completed
  uint64_t* bfull = full + 8;                                       // Toeplitz block has landed
  uint32_t* tslot = reinterpret_cast<uint32_t*>(full + 9);
  float* red = reinterpret_cast<float*>(full + 10);                 // [kTcWarps]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t plane = blockIdx.x >> 1;
  const int j0 = (int)rank * 128;
  const uint32_t colbase = rank ? G::HALO_BYTES : 0;  // A2 / A3 live beside the halo slot the neighbour writes
  step_scalars_k1(dsc, sa, s1, coef);
  PSX_TCTICK(0, 0)
  PSX_TCTICK(1, 0)

  if (warp == kTcWarps) {
    if (lane == 0) {
      for (int i = 0; i < 4; ++i) {
        mbar_init(full + i, kTcEpi);
        mbar_init(done + i, 1);
      }
      mbar_init(bfull, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<512>(tslot);
    if (lane == 0) {
      mbar_expect_tx(bfull, G::B_BYTES);
      bulk_g2s(bsm, bimg, G::B_BYTES, bfull);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tslot;
  PSX_TCTICK(0, 1)
  PSX_TCTICK(1, 1)

  if (warp == kTcWarps) {
    // ------------------------------------------------------------------------------------ UMMA issue
    const uint32_t op_s = smem_u32(op), b_s = smem_u32(bsm);
    const uint64_t dBh = smem_desc(b_s, 2048, 128), dBl = dBh + (uint64_t)(1024 >> 4);
    mbar_wait(bfull, 0);
    PSX_TCTICK(1, 2)
    mbar_wait(full + 0, 0);
    tc_fence_after();
    PSX_TCTICK(1, 3)
    if (lane == 0) {
      const uint32_t id = idesc_f16(128, kTcNT, 0, 0);
      for (int mt = 0; mt < 2; ++mt)
        for (int t = 0; t < 2; ++t)
          issue_tile(tb + (mt * 2 + t) * 64, op_s + t * 8 * G::LBO_ROW + mt * 16 * 128, G::LO_ROW, G::LBO_ROW, dBh, dBl, id,
                     G::KS);
      umma_commit(done + 0);
    }
    __syncwarp();
    PSX_TCTICK(1, 4)
    mbar_wait(done + 0, 0);
    PSX_TCTICK(1, 5)
    cluster_arrive_release();  // #1: this CTA has finished reading A1, its halo slot may be written
    mbar_wait(full + 1, 0);
    tc_fence_after();
    PSX_TCTICK(1, 6)
    if (lane == 0) {
      const uint32_t id = idesc_f16(128, kTcNT, 1, 0);
      for (int t = 0; t < 4; ++t)
        issue_tile(tb + 256 + t * 64, op_s + colbase + t * 8 * G::LBO_COL, G::LO_COL, G::LBO_COL, dBh, dBl, id, G::KS);
      umma_commit(done + 1);
    }
    __syncwarp();
    PSX_TCTICK(1, 7)
    mbar_wait(full + 2, 0);
    tc_fence_after();
    PSX_TCTICK(1, 8)
    if (lane == 0) {
      const uint32_t id = idesc_f16(128, kTcNT, 0, 0);
      for (int t = 0; t < 4; ++t)
        issue_tile(tb + t * 64, op_s + colbase + t * 8 * G::LBO_COL, G::LO_COL, G::LBO_COL, dBh, dBl, id, G::KS);
      umma_commit(done + 2);
    }
    __syncwarp();
    PSX_TCTICK(1, 9)
    cluster_wait_acquire();    // #1
    cluster_arrive_release();  // #2 (this warp writes nothing)
    mbar_wait(full + 3, 0);
    cluster_wait_acquire();    // #2: the neighbour's halo columns of H2 are in this CTA's A4
    fence_async_all();
    tc_fence_after();
    PSX_TCTICK(1, 10)
    if (lane == 0) {
      const uint32_t id = idesc_f16(128, kTcNT, 1, 0);
      for (int mt = 0; mt < 2; ++mt)
        for (int t = 0; t < 2; ++t)
          issue_tile(tb + 256 + (mt * 2 + t) * 64, op_s + t * 8 * G::LBO_ROW + mt * 16 * 128, G::LO_ROW, G::LBO_ROW, dBh,
                     dBl, id, G::KS);
      umma_commit(done + 3);
    }
    __syncwarp();
    PSX_TCTICK(1, 11)
  } else {
    // warp (q, cq): TMEM lanes 32 q .. 32 q + 31, accumulator columns 64 cq .. 64 cq + 63 (= one N-tile)
    const int q = warp & 3, cq = warp >> 2;
    // ------------------------------------------------------------------------------------ load: A1 = x_t - s1 eps
    {
      const float* xp = x + plane * (int64_t)(kTcN * kTcN);
      const float* ep = eps + plane * (int64_t)(kTcN * kTcN);
      const int r = lane & 7, c = lane >> 3;
      // item = (8-row group rg, four K core columns 4 qd .. 4 qd + 3); lane (r, c) converts 8 columns of one row
      constexpr int kItems = 32 * 6, kPer = kItems / kTcWarps, kBatch = 4;
      static_assert(kItems % kTcWarps == 0 && kPer % kBatch == 0, "load schedule");
#pragma unroll 1
      for (int b0 = 0; b0 < kPer; b0 += kBatch) {
        float4 xa[kBatch], xb[kBatch], ea[kBatch], eb[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int it = warp + (b0 + u) * kTcWarps, rg = it / 6, kc = (it % 6) * 4 + c;
          const int j = j0 - PAD + 8 * kc;
          const bool in = kc < G::KC_ROW && j >= 0 && j < kTcN;
          const int off = (8 * rg + r) * kTcN + (in ? j : j0);
          ld_nc8(xp + off, xa[u], xb[u]);  // one full 32-byte sector per lane and instruction
          ld_nc8(ep + off, ea[u], eb[u]);
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int it = warp + (b0 + u) * kTcWarps, rg = it / 6, kc = (it % 6) * 4 + c;
          const int j = j0 - PAD + 8 * kc;
          const bool in = j >= 0 && j < kTcN;
          uint4 hi, lo;
          split2(fmaf(-s1, ea[u].x, xa[u].x), fmaf(-s1, ea[u].y, xa[u].y), hi.x, lo.x);
          split2(fmaf(-s1, ea[u].z, xa[u].z), fmaf(-s1, ea[u].w, xa[u].w), hi.y, lo.y);
          split2(fmaf(-s1, eb[u].x, xb[u].x), fmaf(-s1, eb[u].y, xb[u].y), hi.z, lo.z);
          split2(fmaf(-s1, eb[u].z, xb[u].z), fmaf(-s1, eb[u].w, xb[u].w), hi.w, lo.w);
          if (!in) hi = lo = make_uint4(0, 0, 0, 0);
          if (kc < G::KC_ROW) {
            uint8_t* d = op + kc * G::LBO_ROW + rg * 128 + r * 16;
            *reinterpret_cast<uint4*>(d) = hi;
            *reinterpret_cast<uint4*>(d + G::LO_ROW) = lo;
          }
        }
      }
    }
    fence_async_smem();
    mbar_arrive(full + 0);
    PSX_TCTICK(0, 2)

    // ------------------------------------------------------------------------------------ E1: H1 -> A2 (MN-major)
    mbar_wait(done + 0, 0);
    tc_fence_after();
    PSX_TCTICK(0, 3)
    cluster_arrive_release();  // #1
    {
      const int qi = 128 * (cq >> 1) + 32 * q + lane + PAD;  // K index of this thread's image row
      uint8_t* d = op + colbase + (qi >> 3) * G::LBO_COL + (qi & 7) * 16 + (cq & 1) * 8 * 128;
      const uint32_t ta = tb + ((uint32_t)(32 * q) << 16) + 64 * cq;
      uint32_t v0[32], v1[32];
      tmem_ld32(ta, v0);
      tmem_ld_wait();
      tmem_ld32(ta + 32, v1);
      uint4 hi[4], lo[4];
      split32(v0, bsc.x, hi, lo);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        *reinterpret_cast<uint4*>(d + g * 128) = hi[g];
        *reinterpret_cast<uint4*>(d + G::LO_COL + g * 128) = lo[g];
      }
      tmem_ld_wait();
      split32(v1, bsc.x, hi, lo);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        *reinterpret_cast<uint4*>(d + (4 + g) * 128) = hi[g];
        *reinterpret_cast<uint4*>(d + G::LO_COL + (4 + g) * 128) = lo[g];
      }
      // zero rows above / below the image (K core columns [0, PAD/8) and [32 + PAD/8, KC_COL)); A3 reuses them
      for (int idx = tid; idx < 2 * (PAD / 8) * 256; idx += kTcEpi) {
        const int kz = idx >> 8, kc = kz < PAD / 8 ? kz : 32 + kz;
        reinterpret_cast<uint4*>(op + colbase + kc * G::LBO_COL)[idx & 255] = make_uint4(0, 0, 0, 0);
      }
    }
    tc_fence_before();
    fence_async_smem();
    mbar_arrive(full + 1);
    PSX_TCTICK(0, 4)

    // ------------------------------------------------------------------------------------ E2: r = y - AX -> A3 (K-major)
    float acc = 0.f;
    {
      const int jl = 32 * q + lane;
      const int64_t l = plane / C, ch = plane % C;
      const float* yp = y + ((l / obs_repeat) * C + ch) * (int64_t)(kTcN * kTcN) + (int64_t)(64 * cq) * kTcN + j0 + jl;
      uint8_t* d = op + colbase + (jl >> 3) * 128 + (jl & 7) * 16 + ((64 * cq + PAD) >> 3) * G::LBO_COL;
      const uint32_t ta = tb + ((uint32_t)(32 * q) << 16) + 256 + 64 * cq;
      const float sc = bsc.y / sa;
      float y0[32], y1[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) y0[e] = __ldg(yp + e * kTcN);  // in flight while P2 runs
      mbar_wait(done + 1, 0);
      tc_fence_after();
      PSX_TCTICK(0, 5)
#pragma unroll
      for (int e = 0; e < 32; ++e) y1[e] = __ldg(yp + (32 + e) * kTcN);
      uint32_t v[32];
      uint4 hi[4], lo[4];
      tmem_ld32(ta, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float rr = y0[e] - __uint_as_float(v[e]) * sc;
        acc = fmaf(rr, rr, acc);
        y0[e] = rr;
      }
      tmem_ld32(ta + 32, v);
      split32(reinterpret_cast<const uint32_t(&)[32]>(y0), kTcRScale, hi, lo);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        *reinterpret_cast<uint4*>(d + g * G::LBO_COL) = hi[g];
        *reinterpret_cast<uint4*>(d + g * G::LBO_COL + G::LO_COL) = lo[g];
      }
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float rr = y1[e] - __uint_as_float(v[e]) * sc;
        acc = fmaf(rr, rr, acc);
        y1[e] = rr;
      }
      split32(reinterpret_cast<const uint32_t(&)[32]>(y1), kTcRScale, hi, lo);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        *reinterpret_cast<uint4*>(d + (4 + g) * G::LBO_COL) = hi[g];
        *reinterpret_cast<uint4*>(d + (4 + g) * G::LBO_COL + G::LO_COL) = lo[g];
      }
    }
    tc_fence_before();
    fence_async_smem();
    mbar_arrive(full + 2);
    PSX_TCTICK(0, 6)
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;

    // ------------------------------------------------------------------------------------ E3: H2 -> A4 (MN-major) + halo
    mbar_wait(done + 2, 0);
    tc_fence_after();
    PSX_TCTICK(0, 7)
    cluster_wait_acquire();  // #1: the neighbour has finished reading its A1
    PSX_TCTICK(0, 8)
    {
      const int jl = 32 * q + lane, pk = jl + PAD;
      uint8_t* d = op + (pk >> 3) * G::LBO_ROW + (pk & 7) * 16 + 8 * cq * 128;
      const bool rem = rank == 0 ? jl >= 128 - PAD : jl < PAD;
      const int pkr = rank == 0 ? jl - 128 + PAD : jl + 128 + PAD;
      const uint32_t rd =
          mapa(smem_u32(op) + (uint32_t)((pkr >> 3) * G::LBO_ROW + (pkr & 7) * 16 + 8 * cq * 128), rank ^ 1u);
      const uint32_t ta = tb + ((uint32_t)(32 * q) << 16) + 64 * cq;
      uint32_t v0[32], v1[32];
      tmem_ld32(ta, v0);
      tmem_ld_wait();
      tmem_ld32(ta + 32, v1);
      uint4 hi[4], lo[4];
      split32(v0, bsc.z, hi, lo);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        *reinterpret_cast<uint4*>(d + g * 128) = hi[g];
        *reinterpret_cast<uint4*>(d + G::LO_ROW + g * 128) = lo[g];
      }
      if (rem) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          st_cluster_v4(rd + g * 128, hi[g]);
          st_cluster_v4(rd + G::LO_ROW + g * 128, lo[g]);
        }
      }
      tmem_ld_wait();
      split32(v1, bsc.z, hi, lo);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        *reinterpret_cast<uint4*>(d + (4 + g) * 128) = hi[g];
        *reinterpret_cast<uint4*>(d + G::LO_ROW + (4 + g) * 128) = lo[g];
      }
      if (rem) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          st_cluster_v4(rd + (4 + g) * 128, hi[g]);
          st_cluster_v4(rd + G::LO_ROW + (4 + g) * 128, lo[g]);
        }
      }
      // columns beyond the image border: zero K core columns on the outer side
      uint4* z = reinterpret_cast<uint4*>(op + (rank == 0 ? 0 : G::COL_BYTES));
      for (int idx = tid; idx < G::HALO_BYTES / 16; idx += kTcEpi) z[idx] = make_uint4(0, 0, 0, 0);
    }
    tc_fence_before();
    fence_async_all();
    cluster_arrive_release();  // #2: halo written
    mbar_arrive(full + 3);
    PSX_TCTICK(0, 9)

    // |r|^2 of this CTA's half plane -> its partial-sum slots (the epilogue warps only)
    asm volatile("bar.sync 1, %0;" ::"n"(kTcEpi) : "memory");
    if (tid < pp / 2) {
      float tot = 0.f;
      if (tid == 0)
        for (int i = 0; i < kTcWarps; ++i) tot += red[i];
      err_part[plane * pp + rank * (pp / 2) + tid] = tot;
    }
    cluster_wait_acquire();  // #2
    PSX_TCTICK(0, 10)

    // ------------------------------------------------------------------------------------ E4: cot
    mbar_wait(done + 3, 0);
    tc_fence_after();
    PSX_TCTICK(0, 11)
    {
      const uint32_t ta = tb + ((uint32_t)(32 * q) << 16) + 256 + 64 * cq;
      uint8_t* stage = op + warp * 8192;  // 32 rows x 64 columns fp32 per warp, 16-byte chunks XOR-swizzled by row
      const float sc = bsc.w * coef * (1.f / kTcRScale);
      uint32_t v0[32], v1[32];
      tmem_ld32(ta, v0);
      tmem_ld_wait();
      tmem_ld32(ta + 32, v1);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 o = make_float4(__uint_as_float(v0[4 * g]) * sc, __uint_as_float(v0[4 * g + 1]) * sc,
                                     __uint_as_float(v0[4 * g + 2]) * sc, __uint_as_float(v0[4 * g + 3]) * sc);
        *reinterpret_cast<float4*>(stage + lane * 256 + ((g ^ (lane & 7)) * 16)) = o;
      }
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 o = make_float4(__uint_as_float(v1[4 * g]) * sc, __uint_as_float(v1[4 * g + 1]) * sc,
                                     __uint_as_float(v1[4 * g + 2]) * sc, __uint_as_float(v1[4 * g + 3]) * sc);
        *reinterpret_cast<float4*>(stage + lane * 256 + (((8 + g) ^ (lane & 7)) * 16)) = o;
      }
      __syncwarp();
      // two rows (2 x 256 B) per warp instruction
      const int hr = lane >> 4, cc = lane & 15;
      float* cp = cot + plane * (int64_t)(kTcN * kTcN) + (int64_t)(128 * (cq >> 1) + 32 * q + hr) * kTcN + j0 +
                  64 * (cq & 1) + 4 * cc;
#pragma unroll 4
      for (int r2 = 0; r2 < 16; ++r2) {
        const int rr = 2 * r2 + hr;
        const float4 o = *reinterpret_cast<const float4*>(stage + rr * 256 + ((cc ^ (rr & 7)) * 16));
        st_stream4(cp + (2 * r2) * kTcN, o);
      }
    }
    PSX_TCTICK(0, 12)
  }
  tc_fence_before();
  __syncthreads();
  PSX_TCTICK(0, 13)
  if (warp == kTcWarps) tmem_dealloc<512>(tb);
}

// ------------------------------------------------------------------------------------------ host: Toeplitz block image
// B[n][k] = scale * tap[(k - PAD) - n - lo] for the pass whose 1-D operation is out[p] = sum_i w[i] in[p + lo + i]
// (psx::Taps), as fp16 hi / lo in the shared-memory image [K core column][16 N groups: hi 0-7, lo 8-15][8 x 8 core].
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

void build_image(const Taps& t, int pad, float scale, std::vector<uint8_t>& img) {
  const int KW = kTcNT + 2 * pad;
  img.assign((size_t)(KW / 8) * 2048, 0);
  for (int n = 0; n < kTcNT; ++n)
    for (int k = 0; k < KW; ++k) {
      const int ti = (k - pad) - n - t.lo;
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
  const int pad = (need + 7) & ~7;
  if (pad != 24) return;  // instantiated window: radius 17 .. 24 (the 61-tap sigma = 3 Gaussian prunes to 19)
  if (!(mx > 0.f) || !std::isfinite(mx)) return;
  int e = 0;
  std::frexp(mx, &e);                       // mx = f * 2^e, f in [0.5, 1)
  const float scale = std::ldexp(1.f, 10 - e);  // largest tap -> [512, 1024)
  std::vector<uint8_t> img[4];
  for (int p = 0; p < 4; ++p) build_image(*ts[p], pad, scale, img[p]);
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
  op->tc_pad = pad;
  op->tc_inv_scale = 1.f / scale;
}

void tcblur_release(psx_op* op) {
  if (op->d_tc_img) cudaFree(op->d_tc_img);
  op->d_tc_img = nullptr;
  op->tc_pad = 0;
}

bool tcblur_available(const psx_op* op) { return op->tc_pad != 0 && op->err_parts % (2 * op->C) == 0; }

int launch_pre_sepblur_tc(const psx_op* op, const float* x, const float* eps, const float* y, int64_t L,
                          int64_t obs_repeat, float sa, float s1, float w, const float* dsc, float* cot,
                          float* err_part, cudaStream_t st) {
  using G = TcGeo<24>;
  static bool attr = false;
  if (!attr) {
    if (int rc = check_cuda(cudaFuncSetAttribute(blur_k1_tc<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM),
                            "blur_k1_tc attribute"))
      return rc;
    attr = true;
  }
  const int64_t planes = L * op->C;
  const float coef = (float)((double)w / (double)sa);
  const float s = op->tc_inv_scale;
  blur_k1_tc<24><<<(unsigned)(planes * 2), kTcThreads, G::SMEM, st>>>(x, eps, y, cot, err_part, op->d_tc_img,
                                                                       make_float4(s, s, s, s), op->C, obs_repeat,
                                                                       op->err_parts / op->C, sa, s1, coef, dsc);
  return check_cuda(cudaGetLastError(), "blur_k1_tc launch");
}

}  // namespace psx
