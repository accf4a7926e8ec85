// psx_conv.cu -- shared-memory-tiled convolution kernels (sm_100a):
//   separable blur  A = V . H  (rows then columns), adjoint = H^T . V^T,
//   depthwise 2-D correlation with an arbitrary PSF as row or column segments (motion blur),
//   and K1 (psx_dps_pre) built from them.  Default K1 for a separable blur (W % 32 == 0, H % 16 == 0, both <= 512):
//     conv_rows_pipe<TWEEDIE> : x0 = (x_t - s1 eps)/sa on the fly, h1 = H x0              -> workspace
//     conv_cols16             : r = y - V h1, |r|^2 partials, h2 = V^T r (strip-local)     -> workspace, in place,
//                               row-pair interleaved
//     conv_rows_il            : cot = (w / sa) * H^T h2                                    -> d_cot
//   persistent CTAs fed by TMA (cp.async.bulk / cp.async.bulk.tensor.2d + mbarriers), packed FFMA2 register blocks
//   with the tap pairs in uniform registers; inside a stream capture (and from L = 32) the batch runs as two sample
//   groups on two streams (launch_pre_sepblur).  Other shapes / tap counts fall back to conv_cols_pipe + conv_rows_pipe
//   <COT> or the one-tile-per-CTA conv_rows / conv_cols; blur_k1_fused is the opt-in single-launch cluster variant.
// All passes are zero-padded "same" cross-correlations; halos are zero-filled in shared memory.
#include <cooperative_groups.h>
#include <cuda.h>

#include <cstdlib>

#include "psx_common.cuh"

namespace psx {

enum RowMode { ROWS_PLAIN = 0, ROWS_TWEEDIE = 1, ROWS_COT = 2 };

constexpr int kRowTH = 32;  // rows per tile: 16 row pairs, packed as float2 for FFMA2

// 8 outputs x 8 taps register block on PAIRS: acc[j] += (w_i, w_i) * win[i + j]   (64 FFMA2 = 128 FMA)
__device__ __forceinline__ void fma2_block(float2 (&acc)[8], const float2 (&win)[16], const float2* ww8) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 w = ww8[i];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __ffma2_rn(w, win[i + j], acc[j]);
  }
}

__device__ __forceinline__ float4 load_row4(const float* __restrict__ in, const float* __restrict__ eps,
                                            int64_t plane, int gr, int gc, int H, int W, bool vec_ok,
                                            bool tw, const TweedieC& tc) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (gr >= H) return v;
  const int64_t g = plane + (int64_t)gr * W + gc;
  if (vec_ok && gc >= 0 && gc + 3 < W) {
    v = ld_stream4(in + g);
    if (tw) {
      const float4 e = ld_stream4(eps + g);
      v.x = tweedie(v.x, e.x, tc); v.y = tweedie(v.y, e.y, tc);
      v.z = tweedie(v.z, e.z, tc); v.w = tweedie(v.w, e.w, tc);
    }
  } else {
    float t[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      t[c] = 0.f;
      if (gc + c >= 0 && gc + c < W) {
        t[c] = in[g + c];
        if (tw) t[c] = tweedie(t[c], eps[g + c], tc);
      }
    }
    v = make_float4(t[0], t[1], t[2], t[3]);
  }
  return v;
}

#ifdef PSX_TRACE
// %globaltimer phase trace of the three default-path kernels (tools/micro/k1_trace_main.cu): [kernel][cta][slot]
__device__ long long psx_trace3[3 * 1024 * 8];
#define PSX_TRK(kid, slot)                                                             \
  if (threadIdx.x == 0 && blockIdx.x < 1024) {                                         \
    long long t_;                                                                      \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                             \
    psx_trace3[((kid) * 1024 + blockIdx.x) * 8 + (slot)] = t_;                         \
  }
#else
#define PSX_TRK(kid, slot)
#endif

// ------------------------------------------------------------------------------------------ rows
// grid = (ceil(W/TW), ceil(H/32), planes_total); dynamic smem = 16 * pitch2 float2.
// The tile is stored ROW-PAIR INTERLEAVED: smem2[rp][s] = (row 2rp, row 2rp+1) at image column
// c0 + taps.lo + s, so that one FFMA2 advances two output rows with an aligned register pair for
// every tap.  pitch2 % 16 == 2 makes the LDS.128 of 8 consecutive row pairs hit 8 distinct 16 B banks.
template <int MODE>
__global__ void __launch_bounds__(kThreads)
conv_rows(const float* __restrict__ in, const float* __restrict__ eps, float* __restrict__ out, int H, int W,
          int TW, int pitch2, const __grid_constant__ Taps taps, float sa, float s1, float coef, const float* __restrict__ dsc) {
  extern __shared__ __align__(16) float2 smem2[];
  step_scalars_k1(dsc, sa, s1, coef);
  const TweedieC tc = make_tc(s1, sa);
  const int c0 = blockIdx.x * TW;
  const int r0 = blockIdx.y * kRowTH;
  const int64_t plane = (int64_t)blockIdx.z * H * W;
  const int in_w4 = (TW + taps.k) >> 2;
  const bool vec_ok = (W & 3) == 0;

  for (int idx = threadIdx.x; idx < (kRowTH / 2) * in_w4; idx += kThreads) {
    const int rp = idx / in_w4, c4 = idx - rp * in_w4;
    const int gc = c0 + taps.lo + 4 * c4, gr = r0 + 2 * rp;
    const float4 a = load_row4(in, eps, plane, gr, gc, H, W, vec_ok, MODE == ROWS_TWEEDIE, tc);
    const float4 b = load_row4(in, eps, plane, gr + 1, gc, H, W, vec_ok, MODE == ROWS_TWEEDIE, tc);
    float2* dst = smem2 + rp * pitch2 + 4 * c4;
    *reinterpret_cast<float4*>(dst) = make_float4(a.x, b.x, a.y, b.y);
    *reinterpret_cast<float4*>(dst + 2) = make_float4(a.z, b.z, a.w, b.w);
  }
  __syncthreads();

  const int ntask = (kRowTH / 2) * (TW >> 3);  // <= 512
  float2 res[2][8];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int q = threadIdx.x + t * kThreads;
    if (q < ntask) {
      const int rp = q & 15, g = q >> 4;
      const float2* row = smem2 + rp * pitch2 + 8 * g;
      float2 acc[8], win[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float4 v = *reinterpret_cast<const float4*>(row + 2 * m);
        win[2 * m] = make_float2(v.x, v.y); win[2 * m + 1] = make_float2(v.z, v.w);
      }
      for (int c = 0; c < taps.k; c += 8) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const float4 v = *reinterpret_cast<const float4*>(row + c + 8 + 2 * m);
          win[8 + 2 * m] = make_float2(v.x, v.y); win[9 + 2 * m] = make_float2(v.z, v.w);
        }
        fma2_block(acc, win, taps.ww + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) win[j] = win[j + 8];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) res[t][j] = acc[j];
    }
  }
  __syncthreads();  // input tile fully consumed: reuse it as the (interleaved) output stage

  const int opitch2 = TW + 2;  // TW % 32 == 0  =>  opitch2 % 16 == 2
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int q = threadIdx.x + t * kThreads;
    if (q < ntask) {
      const int rp = q & 15, g = q >> 4;
      float2* o = smem2 + rp * opitch2 + 8 * g;
#pragma unroll
      for (int m = 0; m < 4; ++m)
        *reinterpret_cast<float4*>(o + 2 * m) =
            make_float4(res[t][2 * m].x, res[t][2 * m].y, res[t][2 * m + 1].x, res[t][2 * m + 1].y);
    }
  }
  __syncthreads();

  const int tw4 = TW >> 2;
  for (int idx = threadIdx.x; idx < (kRowTH / 2) * tw4; idx += kThreads) {
    const int rp = idx / tw4, c4 = idx - rp * tw4;
    const int gc = c0 + 4 * c4;
    if (gc >= W) continue;
    const float2* src = smem2 + rp * opitch2 + 4 * c4;
    const float4 p0 = *reinterpret_cast<const float4*>(src), p1 = *reinterpret_cast<const float4*>(src + 2);
    float4 rows[2] = {make_float4(p0.x, p0.z, p1.x, p1.z), make_float4(p0.y, p0.w, p1.y, p1.w)};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gr = r0 + 2 * rp + h;
      if (gr >= H) continue;
      float4 v = rows[h];
      if (MODE == ROWS_COT) {
        v.x = __fmul_rn(coef, v.x); v.y = __fmul_rn(coef, v.y);
        v.z = __fmul_rn(coef, v.z); v.w = __fmul_rn(coef, v.w);
      }
      const int64_t g = plane + (int64_t)gr * W + gc;
      if (vec_ok && gc + 3 < W) {
        st_stream4(out + g, v);
      } else {
        const float t[4] = {v.x, v.y, v.z, v.w};
        for (int c = 0; c < 4 && gc + c < W; ++c) out[g + c] = t[c];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ columns
// CTA = one column strip (TC columns x all H rows) of one plane; a lane owns a PAIR of adjacent columns
// (natural float2 in shared memory) and slides a register window down 8 output rows at a time.
// RESIDUAL: bufA = h1 (+ zero halo) -> r = y - V h1 -> bufB (+ zero halo) -> h2 = V^T r -> out (in place ok)
// PLAIN   : bufA = in (+ zero halo) -> out = taps_f (*) in
template <bool RESIDUAL>
__global__ void __launch_bounds__(kThreads)
conv_cols(const float* in, const float* __restrict__ y, float* out,  // in may alias out (strip-local)
          float* __restrict__ err_part, int C, int H, int W, int TC, int64_t obs_repeat,
          const __grid_constant__ Taps tf, const __grid_constant__ Taps ta) {
  extern __shared__ __align__(16) float2 smem2[];
  __shared__ float red[32];
  float* smem = reinterpret_cast<float*>(smem2);
  const int strip = blockIdx.x, strips = gridDim.x;
  const int64_t pl = blockIdx.y;  // plane index over L*C
  const int c0 = strip * TC;
  const int64_t plane = pl * H * W;
  const int H8 = (H + 7) & ~7;
  const int rowsA = H8 + tf.k;          // buffer row a <-> image row a + tf.lo
  const int rowsB = H8 + ta.k;          // RESIDUAL only: row b <-> image row b + ta.lo
  float* bufA = smem;
  float* bufB = smem + (size_t)rowsA * TC;
  const bool vec4 = (W & 3) == 0, vec2 = (W & 1) == 0;
  const int tc4 = TC >> 2;

  for (int idx = threadIdx.x; idx < rowsA * tc4; idx += kThreads) {
    const int a = idx / tc4, c4 = idx - a * tc4;
    const int gr = a + tf.lo, gc = c0 + 4 * c4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr >= 0 && gr < H) {
      const int64_t g = plane + (int64_t)gr * W + gc;
      if (vec4 && gc + 3 < W) {
        v = *reinterpret_cast<const float4*>(in + g);
      } else {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < 4 && gc + c < W; ++c) t[c] = in[g + c];
        v = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
    *reinterpret_cast<float4*>(bufA + (size_t)a * TC + 4 * c4) = v;
  }
  if (RESIDUAL)
    for (int idx = threadIdx.x; idx < rowsB * tc4; idx += kThreads)
      *reinterpret_cast<float4*>(bufB + 4 * (size_t)idx) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  const int npair = TC >> 1;
  const int ntask = (H8 >> 3) * npair;
  float e2 = 0.f;
  const int64_t yplane = RESIDUAL ? ((pl / C) / obs_repeat * C + pl % C) * (int64_t)H * W : 0;

  for (int q = threadIdx.x; q < ntask; q += kThreads) {
    const int cp = q % npair, g = q / npair;
    const float* col = bufA + (size_t)(8 * g) * TC + 2 * cp;
    float2 acc[8], win[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j] = make_float2(0.f, 0.f);
      win[j] = *reinterpret_cast<const float2*>(col + (size_t)j * TC);
    }
    for (int k = 0; k < tf.k; k += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) win[8 + j] = *reinterpret_cast<const float2*>(col + (size_t)(k + 8 + j) * TC);
      fma2_block(acc, win, tf.ww + k);
#pragma unroll
      for (int j = 0; j < 8; ++j) win[j] = win[j + 8];
    }
    const int gc = c0 + 2 * cp;
    if (gc >= W) continue;
    const bool both = gc + 1 < W;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gr = 8 * g + j;
      if (gr >= H) continue;
      const int64_t g0 = (int64_t)gr * W + gc;
      if (RESIDUAL) {
        float2 yv;
        if (vec2 && both) {
          yv = __ldg(reinterpret_cast<const float2*>(y + yplane + g0));
        } else {
          yv.x = __ldg(y + yplane + g0);
          yv.y = both ? __ldg(y + yplane + g0 + 1) : 0.f;
        }
        float2 r;
        r.x = __fsub_rn(yv.x, acc[j].x);
        r.y = both ? __fsub_rn(yv.y, acc[j].y) : 0.f;
        e2 = fmaf(r.x, r.x, e2);
        e2 = fmaf(r.y, r.y, e2);
        *reinterpret_cast<float2*>(bufB + (size_t)(gr - ta.lo) * TC + 2 * cp) = r;
      } else if (vec2 && both) {
        *reinterpret_cast<float2*>(out + plane + g0) = acc[j];
      } else {
        out[plane + g0] = acc[j].x;
        if (both) out[plane + g0 + 1] = acc[j].y;
      }
    }
  }
  if (!RESIDUAL) return;
  __syncthreads();

  for (int q = threadIdx.x; q < ntask; q += kThreads) {
    const int cp = q % npair, g = q / npair;
    const float* col = bufB + (size_t)(8 * g) * TC + 2 * cp;
    float2 acc[8], win[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j] = make_float2(0.f, 0.f);
      win[j] = *reinterpret_cast<const float2*>(col + (size_t)j * TC);
    }
    for (int k = 0; k < ta.k; k += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) win[8 + j] = *reinterpret_cast<const float2*>(col + (size_t)(k + 8 + j) * TC);
      fma2_block(acc, win, ta.ww + k);
#pragma unroll
      for (int j = 0; j < 8; ++j) win[j] = win[j + 8];
    }
    const int gc = c0 + 2 * cp;
    if (gc >= W) continue;
    const bool both = gc + 1 < W;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gr = 8 * g + j;
      if (gr >= H) continue;
      const int64_t g0 = plane + (int64_t)gr * W + gc;
      if (vec2 && both) {
        *reinterpret_cast<float2*>(out + g0) = acc[j];
      } else {
        out[g0] = acc[j].x;
        if (both) out[g0 + 1] = acc[j].y;
      }
    }
  }
  const float tot = block_sum(e2, red);
  if (threadIdx.x == 0) {
    const int64_t l = pl / C;
    const int ch = (int)(pl % C);
    err_part[l * (int64_t)(C * strips) + ch * strips + strip] = tot;
  }
}

// ========================================================================================== pipelined fast path
// Persistent CTAs walk over tiles.  While tile i is computed, tile i+1 is already being copied into the
// other shared-memory stage by the bulk-copy engine (cp.async.bulk global->shared, one instruction per
// row segment, completion signalled on an mbarrier with complete_tx) -- no per-element copy instructions,
// no register staging, and the global-load latency that dominated the one-tile-per-CTA kernels (ncu:
// >30 % of stall samples on the first use of a loaded value) hides behind the FFMA2 work.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// The vote makes the loop condition warp-uniform *for the compiler*: code after the wait stays eligible
// for the uniform datapath (LDCU + FFMA2 with a UR operand), which is what keeps FFMA2 at 2 cycles
// (3 vector register pairs per FFMA2 cost 3 register-file cycles; 2 pairs + 1 uniform pair cost 2).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!__all_sync(0xffffffffu, ok));
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}

// The FFMA2 work of one task: 8 outputs x (taps.k) taps on float2 pairs.  K > 0: the tap loop is fully
// unrolled -- the tap pairs become immediate-offset constant-bank loads into uniform registers that the
// compiler hoists and widens (22 LDCU for 40 taps), the window loads are software-pipelined, and nothing but
// FFMA2 R,R,UR,R + LDS is left (36 registers).  K == 0: generic runtime loop.
template <int K>
__device__ __forceinline__ void row_block(float2 (&acc)[8], const float2* __restrict__ row, const Taps& taps) {
  float2 win[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const float4 v = *reinterpret_cast<const float4*>(row + 2 * m);
    win[2 * m] = make_float2(v.x, v.y); win[2 * m + 1] = make_float2(v.z, v.w);
  }
  if constexpr (K > 0) {
#pragma unroll
    for (int c = 0; c < K; c += 8) {
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float4 v = *reinterpret_cast<const float4*>(row + c + 8 + 2 * m);
        win[8 + 2 * m] = make_float2(v.x, v.y); win[9 + 2 * m] = make_float2(v.z, v.w);
      }
      fma2_block(acc, win, taps.ww + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) win[j] = win[j + 8];
    }
  } else {
    for (int c = 0; c < taps.k; c += 8) {
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float4 v = *reinterpret_cast<const float4*>(row + c + 8 + 2 * m);
        win[8 + 2 * m] = make_float2(v.x, v.y); win[9 + 2 * m] = make_float2(v.z, v.w);
      }
      fma2_block(acc, win, taps.ww + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) win[j] = win[j + 8];
    }
  }
}

template <int K, int TC>
__device__ __forceinline__ void col_block(float2 (&acc)[8], const float* __restrict__ col, const Taps& taps) {
  float2 win[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[j] = make_float2(0.f, 0.f);
    win[j] = *reinterpret_cast<const float2*>(col + j * TC);
  }
  if constexpr (K > 0) {
#pragma unroll
    for (int k = 0; k < K; k += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) win[8 + j] = *reinterpret_cast<const float2*>(col + (k + 8 + j) * TC);
      fma2_block(acc, win, taps.ww + k);
#pragma unroll
      for (int j = 0; j < 8; ++j) win[j] = win[j + 8];
    }
  } else {
    for (int k = 0; k < taps.k; k += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) win[8 + j] = *reinterpret_cast<const float2*>(col + (k + 8 + j) * TC);
      fma2_block(acc, win, taps.ww + k);
#pragma unroll
      for (int j = 0; j < 8; ++j) win[j] = win[j + 8];
    }
  }
}

constexpr int kPipeRows = 16;   // rows per tile of the pipelined row pass: 8 row pairs (rho, rho + 8)
constexpr int kPipeHdr = 128;   // bytes reserved at the start of dynamic smem for the two mbarriers

// NOTE on loop shapes in the two kernels below: every per-thread loop ahead of an FFMA2 block is fully
// unrolled (compile-time trip count, guards inside) or has a runtime trip count with no thread-dependent
// branch.  Otherwise ptxas treats the warp as possibly divergent from there on, stops using the uniform
// datapath, and FFMA2 with three vector-register pairs costs 3 register-file cycles instead of the 2 of
// FFMA2 R, R, UR, R.

// ---- rows.  The batch is one tall matrix of total_rows = planes * H rows x W columns (the padding is
// horizontal only, so tiles may straddle planes).  A tile is 16 full rows = ONE contiguous 16*W*4-byte
// block per input array: the producer is a single cp.async.bulk per array and tile.  Per CTA: 2 raw
// stages + one compute tile in the row-pair-interleaved float2 layout of conv_rows (pairs (rho, rho+8);
// pitch2 % 16 == 2 => conflict-free LDS.128) whose halo columns are zeroed once.  A conversion pass
// (Tweedie or plain copy) moves a landed stage into the compute tile.  Needs W % 8 == 0, W <= 512.
// HALF (Tweedie mode only): `in` / `eps` point to bf16 arrays (the bf16 sampler state); the raw stages hold bf16 rows
// and the conversion pass widens them -- everything after it is unchanged.
template <int MODE, int K, bool HALF = false>
__global__ void __launch_bounds__(kThreads)
conv_rows_pipe(const float* __restrict__ in, const float* __restrict__ eps, float* __restrict__ out,
               int64_t total_rows, int W, int pitch2, int64_t num_tiles, const __grid_constant__ Taps taps,
               float sa, float s1, float coef, const float* __restrict__ dsc) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (MODE == ROWS_TWEEDIE) { PSX_TRK(0, 0) }
  step_scalars_k1(dsc, sa, s1, coef);
  const TweedieC tc = make_tc(s1, sa);
  constexpr int kArrays = MODE == ROWS_TWEEDIE ? 2 : 1;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float2* comp = reinterpret_cast<float2*>(smem_raw + kPipeHdr);              // 8 x pitch2 float2
  float* raw = reinterpret_cast<float*>(comp + (kPipeRows / 2) * pitch2);       // 2 stages x kArrays x 16 x W
  const int tile_floats = HALF ? (kPipeRows * W) >> 1 : kPipeRows * W;  // one array of one stage, in 4-byte units
  const int w4 = W >> 2;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int64_t tile, int stage) {
    if (threadIdx.x == 0) {
      const int64_t r0 = tile * kPipeRows;
      const int64_t left = total_rows - r0;
      constexpr uint32_t esize = HALF ? 2u : 4u;
      const uint32_t nb = (uint32_t)(left < kPipeRows ? left : kPipeRows) * (uint32_t)W * esize;
      float* xs = raw + (size_t)stage * kArrays * tile_floats;
      mbar_expect_tx(&bars[stage], nb * kArrays);
      bulk_g2s(xs, reinterpret_cast<const char*>(in) + r0 * W * esize, nb, &bars[stage]);
      if (kArrays == 2) bulk_g2s(xs + tile_floats, reinterpret_cast<const char*>(eps) + r0 * W * esize, nb, &bars[stage]);
    }
  };

  int64_t tile = blockIdx.x;
  if (tile < num_tiles) issue(tile, 0);
  {  // zero the compute tile once while the first loads are in flight (its halo columns are never written
     // again; the raw stages are written by the bulk copies only)
    const int n4 = ((kPipeRows / 2) * pitch2) >> 1;
#pragma unroll
    for (int t = 0; t < 11; ++t) {  // 8 x (512 + 136 + 14) float2 / 2 <= 11 * 256
      const int i = threadIdx.x + t * kThreads;
      if (i < n4) reinterpret_cast<float4*>(comp)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (MODE == ROWS_TWEEDIE) { PSX_TRK(0, 1) }
  for (int it = 0; tile < num_tiles; ++it, tile += gridDim.x) {
    const int stage = it & 1;
    if (tile + gridDim.x < num_tiles) issue(tile + gridDim.x, stage ^ 1);
    mbar_wait(&bars[stage], (uint32_t)(it >> 1) & 1u);
    if (MODE == ROWS_TWEEDIE && it == 0) { PSX_TRK(0, 2) }
    __syncthreads();  // B1: previous tile's compute is done with `comp` (first pass: the zero fill is done)
    const float* xs = raw + (size_t)stage * kArrays * tile_floats;
    const int64_t r0 = tile * kPipeRows;
    const int64_t left = total_rows - r0;
    const int rows_valid = left < kPipeRows ? (int)left : kPipeRows;
#pragma unroll
    for (int t = 0; t < 4; ++t) {  // 8 row pairs x W/4 column quads <= 4 * 256 for W <= 512
      const int idx = threadIdx.x + t * kThreads;
      const int rp = idx / w4, c4 = idx - rp * w4;
      if (rp < kPipeRows / 2) {
        const bool va = rp < rows_valid, vb = rp + 8 < rows_valid;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        auto ld4 = [&](const float* base, int row) {  // 4 consecutive columns of one raw row
          if (HALF) {
            const uint2 h = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(base) + ((size_t)row * W + 4 * c4) * 2);
            return make_float4(bf16_lo(h.x), bf16_hi(h.x), bf16_lo(h.y), bf16_hi(h.y));
          }
          return *reinterpret_cast<const float4*>(base + row * W + 4 * c4);
        };
        if (va) a = ld4(xs, rp);
        if (vb) b = ld4(xs, rp + 8);
        if (MODE == ROWS_TWEEDIE) {
          float4 ea = make_float4(0.f, 0.f, 0.f, 0.f), eb = ea;
          if (va) ea = ld4(xs + tile_floats, rp);
          if (vb) eb = ld4(xs + tile_floats, rp + 8);
          a.x = tweedie(a.x, ea.x, tc); a.y = tweedie(a.y, ea.y, tc);
          a.z = tweedie(a.z, ea.z, tc); a.w = tweedie(a.w, ea.w, tc);
          b.x = tweedie(b.x, eb.x, tc); b.y = tweedie(b.y, eb.y, tc);
          b.z = tweedie(b.z, eb.z, tc); b.w = tweedie(b.w, eb.w, tc);
        }
        float2* dst = comp + rp * pitch2 - taps.lo + 4 * c4;  // smem column s <-> image column s + lo
        *reinterpret_cast<float4*>(dst) = make_float4(a.x, b.x, a.y, b.y);
        *reinterpret_cast<float4*>(dst + 2) = make_float4(a.z, b.z, a.w, b.w);
      }
    }
    __syncthreads();  // B2: `comp` complete; raw[stage] fully consumed (may be refilled from now on)
    if (MODE == ROWS_TWEEDIE && it == 0) { PSX_TRK(0, 3) }

    const int ncg = W >> 3;
#pragma unroll
    for (int t = 0; t < 2; ++t) {  // W/8 column groups x 8 row pairs <= 2 * 256 tasks
      const int q_raw = threadIdx.x + t * kThreads;
      const bool q_ok = q_raw < 8 * ncg;
      const int q = q_ok ? q_raw : 8 * ncg - 1;  // every thread computes (uniform control flow); stores are guarded
      const int rp = q & 7, cg = q >> 3;
      const float2* row = comp + rp * pitch2 + 8 * cg;
      float2 acc[8];
      row_block<K>(acc, row, taps);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool ok = q_ok && rp + 8 * h < rows_valid;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = h ? acc[j].y : acc[j].x;
          if (MODE == ROWS_COT) o[j] = __fmul_rn(coef, o[j]);
        }
        float* dst = out + (r0 + rp + 8 * h) * W + 8 * cg;
        if (ok) {
          st_stream4(dst, make_float4(o[0], o[1], o[2], o[3]));
          st_stream4(dst + 4, make_float4(o[4], o[5], o[6], o[7]));
        }
      }
      if (W <= 256) break;  // uniform: one round covers all tasks
    }
    // no trailing barrier: B1 of the next iteration orders this compute before the next conversion pass,
    // and the stage refilled next (this one) was fully read before B2.
    if (MODE == ROWS_TWEEDIE && it == 0) { PSX_TRK(0, 4) }
  }
  if (MODE == ROWS_TWEEDIE) { PSX_TRK(0, 5) }
}

// ---- columns.  Tile = one 32-column strip (all H rows) of one plane, fetched by ONE 2-D TMA box load
// (cp.async.bulk.tensor.2d, box = 32 x H) into a stage whose zero halo rows persist across tiles; a
// lane owns a column pair and 8 output rows.  r lives in one extra buffer; y is prefetched straight from
// global (it is shared by all samples: L2 hits).  Needs W % 32 == 0, H % 16 == 0, H <= 256 (partial last round guarded).
constexpr int kColTC = 32;

template <bool RESIDUAL, int ROUNDS, int K>  // ROUNDS = H / 8 * 16 / kThreads: tasks per thread and pass (1 or 2)
__global__ void __launch_bounds__(kThreads, 2)
conv_cols_pipe(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ y, float* out,
               float* __restrict__ err_part, int C, int H, int W, int strips, int64_t num_tiles,
               int64_t obs_repeat, const __grid_constant__ Taps tf, const __grid_constant__ Taps ta) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float red[32];
  constexpr int TC = kColTC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* smem = reinterpret_cast<float*>(smem_raw + kPipeHdr);
  const int rowsA = H + tf.k;
  const int stage_floats = rowsA * TC;
  float* bufB = smem + 2 * (size_t)stage_floats;  // RESIDUAL only, (H + ta.k) rows

  {  // zero the halo rows once (8 float4 per row): stage rows outside [-lo, -lo + H), same for bufB
#pragma unroll
    for (int t = 0; t < 5; ++t) {  // (136 + 8) halo rows * 8 float4 <= 5 * 256
      const int i = threadIdx.x + t * kThreads;
      const int r = i >> 3, c4 = i & 7;
      if (r < tf.k) {
        const int row = r < -tf.lo ? r : r + H;
        *reinterpret_cast<float4*>(smem + (size_t)row * TC + 4 * c4) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(smem + stage_floats + (size_t)row * TC + 4 * c4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (RESIDUAL && r < ta.k) {
        const int row = r < -ta.lo ? r : r + H;
        *reinterpret_cast<float4*>(bufB + (size_t)row * TC + 4 * c4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int64_t tile, int stage) {
    if (threadIdx.x == 0) {
      const int64_t pl = tile / strips;
      const int c0 = (int)(tile - pl * strips) * TC;
      mbar_expect_tx(&bars[stage], (uint32_t)H * TC * 4u);
      tma_load_2d(smem + (size_t)stage * stage_floats - (size_t)tf.lo * TC, &tmap, c0, (int)(pl * H), &bars[stage]);
    }
  };

  constexpr int npair = TC >> 1;
  int64_t tile = blockIdx.x;
  if (tile < num_tiles) issue(tile, 0);
  for (int it = 0; tile < num_tiles; ++it, tile += gridDim.x) {
    const int stage = it & 1;
    if (tile + gridDim.x < num_tiles) issue(tile + gridDim.x, stage ^ 1);
    mbar_wait(&bars[stage], (uint32_t)(it >> 1) & 1u);
    __syncthreads();
    const float* A = smem + (size_t)stage * stage_floats;
    const int64_t pl = tile / strips;
    const int strip = (int)(tile - pl * strips);
    const int c0 = strip * TC;
    const int64_t plane = pl * H * W;
    const int64_t yplane = RESIDUAL ? ((pl / C) / obs_repeat * C + pl % C) * (int64_t)H * W : 0;
    float e2 = 0.f;

    const int ntask = (H >> 3) * npair;  // <= ROUNDS * 256; the last round may be partial (H = 144 .. 240)
#pragma unroll
    for (int t = 0; t < ROUNDS; ++t) {
      const int q_raw = threadIdx.x + t * kThreads;
      const bool ok = q_raw < ntask;
      const int q = ok ? q_raw : ntask - 1;  // every thread computes (uniform control flow); stores are guarded
      const int cp = q % npair, g = q / npair;
      const int gc = c0 + 2 * cp;
      float2 yv[8];
      if (RESIDUAL) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          yv[j] = __ldg(reinterpret_cast<const float2*>(y + yplane + (int64_t)(8 * g + j) * W + gc));
      }
      const float* col = A + (size_t)(8 * g) * TC + 2 * cp;
      float2 acc[8];
      col_block<K, TC>(acc, col, tf);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int gr = 8 * g + j;
        if (RESIDUAL) {
          float2 r;
          r.x = ok ? __fsub_rn(yv[j].x, acc[j].x) : 0.f;
          r.y = ok ? __fsub_rn(yv[j].y, acc[j].y) : 0.f;
          e2 = fmaf(r.x, r.x, e2);
          e2 = fmaf(r.y, r.y, e2);
          if (ok) *reinterpret_cast<float2*>(bufB + (size_t)(gr - ta.lo) * TC + 2 * cp) = r;
        } else if (ok) {
          *reinterpret_cast<float2*>(out + plane + (int64_t)gr * W + gc) = acc[j];
        }
      }
    }
    if (RESIDUAL) {
      __syncthreads();
#pragma unroll
      for (int t = 0; t < ROUNDS; ++t) {
        const int q_raw = threadIdx.x + t * kThreads;
        const bool ok = q_raw < ntask;
        const int q = ok ? q_raw : ntask - 1;
        const int cp = q % npair, g = q / npair;
        const float* col = bufB + (size_t)(8 * g) * TC + 2 * cp;
        float2 acc[8];
        col_block<K, TC>(acc, col, ta);
        const int gc = c0 + 2 * cp;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (ok) *reinterpret_cast<float2*>(out + plane + (int64_t)(8 * g + j) * W + gc) = acc[j];
      }
      const float tot = block_sum(e2, red);  // contains a __syncthreads
      if (threadIdx.x == 0) {
        const int64_t l = pl / C;
        const int ch = (int)(pl % C);
        err_part[l * (int64_t)(C * strips) + ch * strips + strip] = tot;
      }
    }
    __syncthreads();  // stage + bufB + red fully consumed
  }
}

// ========================================================================================== 16-output blocks
// ncu on the 8-output kernels: L1TEX/shared-memory data pipe 62-72 % busy while the FMA pipe sat at 40-55 %:
// per tile the window re-reads (48 float2 per 8x2 outputs), the raw->interleaved conversion pass and the TMA
// fill together cost about as many shared-memory wavefronts as the FFMA2 work costs FMA-pipe cycles.  The
// kernels below halve the window traffic (16 outputs per task: 56 float2 per 16x2 outputs) and drop the
// conversion pass for the last row kernel by making its producer (the column kernel) write the row-pair
// INTERLEAVED layout straight to global memory:   IL[(rho >> 1) * 2W + 2c + (rho & 1)] = X[rho][c].
template <int K>
__device__ __forceinline__ void row_block16(float2 (&acc)[16], const float2* __restrict__ row, const Taps& taps) {
  float2 win[24];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const float4 v = *reinterpret_cast<const float4*>(row + 2 * m);
    win[2 * m] = make_float2(v.x, v.y); win[2 * m + 1] = make_float2(v.z, v.w);
  }
#pragma unroll
  for (int c = 0; c < K; c += 8) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const float4 v = *reinterpret_cast<const float4*>(row + c + 16 + 2 * m);
      win[16 + 2 * m] = make_float2(v.x, v.y); win[17 + 2 * m] = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 w = taps.ww[c + i];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = __ffma2_rn(w, win[i + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) win[j] = win[j + 8];
  }
}

template <int K, int TC>
__device__ __forceinline__ void col_block16(float2 (&acc)[16], const float* __restrict__ col, const Taps& taps) {
  float2 win[24];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    acc[j] = make_float2(0.f, 0.f);
    win[j] = *reinterpret_cast<const float2*>(col + j * TC);
  }
#pragma unroll
  for (int k = 0; k < K; k += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) win[16 + j] = *reinterpret_cast<const float2*>(col + (k + 16 + j) * TC);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 w = taps.ww[k + i];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = __ffma2_rn(w, win[i + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) win[j] = win[j + 8];
  }
}

// ---- columns, fused V / residual / V^T, 16 rows x 2 columns per task, h2 written row-pair interleaved.
// Strip width TC = 32 for planes up to 256 rows, 16 for up to 512 rows (at most one task per thread and pass:
// (H / 16) * (TC / 2) <= 256); needs W % TC == 0, H % 16 == 0, tf.k == ta.k == K.
template <int K, int TC>
__global__ void __launch_bounds__(kThreads, 2)
conv_cols16(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ y, float* __restrict__ out_il,
            float* __restrict__ err_part, int C, int H, int W, int strips, int64_t num_tiles, int64_t obs_repeat,
            int64_t sample0, int box_h, const __grid_constant__ Taps tf, const __grid_constant__ Taps ta) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float red[32];
  PSX_TRK(1, 0)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* smem = reinterpret_cast<float*>(smem_raw + kPipeHdr);
  const int rowsA = H + K;
  const int stage_floats = rowsA * TC;
  float* bufB = smem + 2 * (size_t)stage_floats;  // (H + K) rows

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int64_t tile, int stage) {
    if (threadIdx.x == 0) {
      const int64_t pl = tile / strips;
      const int c0 = (int)(tile - pl * strips) * TC;
      mbar_expect_tx(&bars[stage], (uint32_t)H * TC * 4u);
      float* dst = smem + (size_t)stage * stage_floats - (size_t)tf.lo * TC;
      for (int h0 = 0; h0 < H; h0 += box_h)  // a TMA box holds at most 256 rows: taller planes take two boxes
        tma_load_2d(dst + (size_t)h0 * TC, &tmap, c0, (int)(pl * H) + h0, &bars[stage]);
    }
  };

  constexpr int npair = TC >> 1;
  const int ntask = (H >> 4) * npair;  // <= 256
  const bool live = threadIdx.x < ntask;
  const int q = live ? threadIdx.x : ntask - 1;
  const int cp = q % npair, g = q / npair;
  int64_t tile = blockIdx.x;
  if (tile < num_tiles) issue(tile, 0);
#pragma unroll
  for (int t = 0; t < 5; ++t) {  // zero the halo rows once, while the first box load is in flight (the TMA
    const int i = threadIdx.x + t * kThreads;  // writes interior rows only): (136 + 8) rows * 8 float4 <= 5 * 256
    const int r = i / (TC / 4), c4 = i % (TC / 4);
    if (r < K) {
      const int ra = r < -tf.lo ? r : r + H, rb = r < -ta.lo ? r : r + H;
      *reinterpret_cast<float4*>(smem + (size_t)ra * TC + 4 * c4) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(smem + stage_floats + (size_t)ra * TC + 4 * c4) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(bufB + (size_t)rb * TC + 4 * c4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  PSX_TRK(1, 1)
  for (int it = 0; tile < num_tiles; ++it, tile += gridDim.x) {
    const int stage = it & 1;
    if (tile + gridDim.x < num_tiles) issue(tile + gridDim.x, stage ^ 1);
    mbar_wait(&bars[stage], (uint32_t)(it >> 1) & 1u);
    if (it == 0) { PSX_TRK(1, 2) }
    __syncthreads();
    const float* A = smem + (size_t)stage * stage_floats;
    const int64_t pl = tile / strips;
    const int strip = (int)(tile - pl * strips);
    const int gc = strip * TC + 2 * cp;
    const int64_t yplane = ((sample0 + pl / C) / obs_repeat * C + pl % C) * (int64_t)H * W;  // pl: plane within this launch
    float e2 = 0.f;
    {
      float2 acc[16];
      col_block16<K, TC>(acc, A + (size_t)(16 * g) * TC + 2 * cp, tf);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int gr = 16 * g + j;
        const float2 yv = __ldg(reinterpret_cast<const float2*>(y + yplane + (int64_t)gr * W + gc));
        float2 r;
        r.x = live ? __fsub_rn(yv.x, acc[j].x) : 0.f;
        r.y = live ? __fsub_rn(yv.y, acc[j].y) : 0.f;
        e2 = fmaf(r.x, r.x, e2);
        e2 = fmaf(r.y, r.y, e2);
        if (live) *reinterpret_cast<float2*>(bufB + (size_t)(gr - ta.lo) * TC + 2 * cp) = r;
      }
    }
    __syncthreads();
    if (it == 0) { PSX_TRK(1, 3) }
    {
      float2 acc[16];
      col_block16<K, TC>(acc, bufB + (size_t)(16 * g) * TC + 2 * cp, ta);
      float* dst = out_il + ((pl * H + 16 * g) >> 1) * (int64_t)(2 * W) + 2 * gc;
#pragma unroll
      for (int m = 0; m < 8; ++m)
        if (live)
          *reinterpret_cast<float4*>(dst + (int64_t)m * 2 * W) =
              make_float4(acc[2 * m].x, acc[2 * m + 1].x, acc[2 * m].y, acc[2 * m + 1].y);
    }
    const float tot = block_sum(e2, red);  // contains a __syncthreads
    if (threadIdx.x == 0) {
      const int64_t l = pl / C;
      const int ch = (int)(pl % C);
      err_part[l * (int64_t)(C * strips) + ch * strips + strip] = tot;
    }
    __syncthreads();  // stage + bufB + red fully consumed
    if (it == 0) { PSX_TRK(1, 4) }
  }
  PSX_TRK(1, 5)
}

// ---- last row pass on row-pair-interleaved input: the TMA bulk copies land directly in the compute
// layout (no conversion pass, no second barrier), 3 stages, 16 outputs x 2 rows per task, 128 threads.
// Needs W % 16 == 0, W <= 256 * ROUNDS (ROUNDS task rounds per tile), (planes * H) % 2 == 0.
constexpr int kIlThreads = 128;
constexpr int kIlStages = 3;

template <int K, int ROUNDS, bool HALF = false>  // HALF: `out` is a bf16 array (the bf16 cotangent)
__global__ void __launch_bounds__(kIlThreads)
conv_rows_il(const float* __restrict__ in_il, float* __restrict__ out, int64_t total_rows, int W, int pitch2,
             int64_t num_tiles, const __grid_constant__ Taps taps, float coef,
             const float* __restrict__ dsc) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PSX_TRK(2, 0)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float2* stages = reinterpret_cast<float2*>(smem_raw + kPipeHdr);
  const int stage_f2 = (kPipeRows / 2) * pitch2;
  const int64_t total_pairs = total_rows >> 1;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kIlStages; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int64_t tile, int stage) {
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      const int64_t p0 = tile * (kPipeRows / 2);
      const int64_t left = total_pairs - p0;
      const int valid = left < kPipeRows / 2 ? (int)left : kPipeRows / 2;
      const uint32_t nb = (uint32_t)W * 8u;  // one row pair: 2W floats
      if (lane == 0) mbar_expect_tx(&bars[stage], nb * (uint32_t)valid);
      __syncwarp();
      if (lane < valid)
        bulk_g2s(stages + (size_t)stage * stage_f2 + lane * pitch2 - taps.lo, in_il + (p0 + lane) * 2 * W, nb,
                 &bars[stage]);
    }
  };

  const int ncg = W >> 4;  // ROUNDS * kIlThreads >= 8 * ncg tasks: row pair rp = q & 7, 16-column group cg = q >> 3
  int64_t tile = blockIdx.x;
#pragma unroll
  for (int d = 0; d < kIlStages - 1; ++d)
    if (tile + (int64_t)d * gridDim.x < num_tiles) issue(tile + (int64_t)d * gridDim.x, d);
  {  // zero the halo columns of every stage once, while the first loads are in flight (the bulk copies write the
     // interior columns [-lo, -lo + W) only): per row pair the float2 columns [0, -lo) and [-lo + W, pitch2).
     // pitch2 - W <= K + 16 <= 64 float2 = 32 float4 per row pair: shifts only, no integer division.
    const int left = -taps.lo, per_row4 = (pitch2 - W) >> 1;  // left, pitch2 - W even
#pragma unroll
    for (int t = 0; t < (kIlStages * (kPipeRows / 2) * 32) / kIlThreads; ++t) {
      const int i = threadIdx.x + t * kIlThreads;
      const int row = i >> 5, c2 = 2 * (i & 31);
      if ((i & 31) < per_row4) {
        const int col = c2 < left ? c2 : c2 + W;
        *reinterpret_cast<float4*>(stages + (size_t)row * pitch2 + col) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  PSX_TRK(2, 1)
  for (int it = 0; tile < num_tiles; ++it, tile += gridDim.x) {
    const int stage = it % kIlStages;
    __syncthreads();  // everyone is done with the stage that is refilled next (first pass: halo zero fill done)
    const int64_t nxt = tile + (int64_t)(kIlStages - 1) * gridDim.x;
    if (nxt < num_tiles) issue(nxt, (it + kIlStages - 1) % kIlStages);
    mbar_wait(&bars[stage], (uint32_t)(it / kIlStages) & 1u);
    if (it == 0) { PSX_TRK(2, 2) }
#pragma unroll
    for (int round = 0; round < ROUNDS; ++round) {
      const int q_raw = threadIdx.x + round * kIlThreads;
      const bool q_ok = q_raw < 8 * ncg;
      const int q = q_ok ? q_raw : 8 * ncg - 1;  // every thread computes (uniform control flow); stores are guarded
      const int rp = q & 7, cg = q >> 3;
      float2 acc[16];
      row_block16<K>(acc, stages + (size_t)stage * stage_f2 + rp * pitch2 + 16 * cg, taps);
      if (it == 0 && round == 0) { PSX_TRK(2, 3) }
      const int64_t pair = tile * (kPipeRows / 2) + rp;
      if (q_ok && pair < total_pairs) {
        step_scalars_coef(dsc, coef);  // after the FFMA2 block: keeps the tap pairs on the uniform datapath
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float* dst = out + (2 * pair + h) * W + 16 * cg;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            float4 v;
            v.x = __fmul_rn(coef, h ? acc[4 * m].y : acc[4 * m].x);
            v.y = __fmul_rn(coef, h ? acc[4 * m + 1].y : acc[4 * m + 1].x);
            v.z = __fmul_rn(coef, h ? acc[4 * m + 2].y : acc[4 * m + 2].x);
            v.w = __fmul_rn(coef, h ? acc[4 * m + 3].y : acc[4 * m + 3].x);
            if (HALF) {  // 4 bf16 = 8 bytes at element offset (row * W + 16 cg + 4 m)
              char* d16 = reinterpret_cast<char*>(out) + (((2 * pair + h) * W + 16 * cg + 4 * m) << 1);
              *reinterpret_cast<uint2*>(d16) = make_uint2(bf16_pack2(v.x, v.y), bf16_pack2(v.z, v.w));
            } else {
              st_stream4(dst + 4 * m, v);
            }
          }
        }
      }
    }
    if (it == 0) { PSX_TRK(2, 4) }
  }
  PSX_TRK(2, 5)
}

// ========================================================================================== cluster-fused K1
// One launch for the whole blur K1 on 256 x 256 planes.  A thread-block CLUSTER of 8 CTAs owns one plane
// (CTA c: rows [32c, 32c + 32)); the plane never leaves shared memory between the four passes:
//   TMA bulk: x_t, eps rows -> smem | Tweedie -> row-pair-interleaved tile | H pass | h1 -> row-major buffer |
//   cluster.sync + halo rows pulled from the two neighbour CTAs over distributed shared memory | V pass,
//   r = y - V h1, |r|^2 | r -> buffer, halo exchange | V^T pass -> interleaved tile | H^T pass -> cot.
// HBM traffic is exactly the algorithmic 16 B/element (x_t, eps, y in; cot out); no workspace.
// Needs H == W == 256, all four tap sets with k == K and equal offsets per direction.
#ifdef PSX_TRACE
__device__ long long psx_trace[4096 * 16];
#define PSX_TICK(slot)                                                                 \
  if (threadIdx.x == 0 && blockIdx.x < 4096) {                                         \
    long long t_;                                                                      \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                             \
    psx_trace[blockIdx.x * 16 + (slot)] = t_;                                          \
  }
#else
#define PSX_TICK(slot)
#endif

constexpr int kFusedCL = 8;      // CTAs per cluster (= per plane)
constexpr int kFusedRB = 32;     // rows per CTA
constexpr int kFusedW = 256;
constexpr int kFusedHP = 260;    // pitch (floats) of the row-major h1 / r buffer: % 32 == 4

template <int K>
__global__ void __cluster_dims__(kFusedCL, 1, 1) __launch_bounds__(kThreads, 2)
blur_k1_fused(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ y,
              float* __restrict__ cot, float* __restrict__ err_part, int C, int64_t obs_repeat, int pitch2,
              const __grid_constant__ Taps fh, const __grid_constant__ Taps fv, const __grid_constant__ Taps av,
              const __grid_constant__ Taps ah, float sa, float s1, float coef, const float* __restrict__ dsc) {
  namespace cg = cooperative_groups;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float red[32];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int64_t pl = blockIdx.x / kFusedCL;  // plane index over L * C
  constexpr int W = kFusedW, RB = kFusedRB, HP = kFusedHP, H = kFusedCL * kFusedRB;
  step_scalars_k1(dsc, sa, s1, coef);
  const TweedieC tc = make_tc(s1, sa);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  float2* comp = reinterpret_cast<float2*>(smem_raw + kPipeHdr);              // 16 row pairs x pitch2
  float* hb = reinterpret_cast<float*>(comp + (RB / 2) * pitch2);               // (RB + K) rows x HP; raw x|eps alias
  const int64_t gbase = pl * (int64_t)H * W + (int64_t)rank * RB * W;           // first own element

  PSX_TICK(0)
  // ---- P0: zero the interleaved tile (its halo columns stay zero) and the outer halo rows of the edge CTAs
  // (inner halo rows are always fully overwritten by the neighbours' pushes), arm the barrier, start the load
  {
    const int n4 = ((RB / 2) * pitch2) >> 1;
#pragma unroll
    for (int t = 0; t < 11; ++t) {
      const int i = threadIdx.x + t * kThreads;
      if (i < n4) reinterpret_cast<float4*>(comp)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int t = 0; t < 11; ++t) {  // K (40) halo rows x 65 quads (pitch 260) = 2600 <= 11 * 256
      const int i = threadIdx.x + t * kThreads;
      const int hr = i / (HP / 4), c4 = i - hr * (HP / 4);
      if (hr < K) {
        const bool top = hr < -fv.lo;
        const int b = top ? hr : hr + RB;
        if ((top && rank == 0) || (!top && rank == kFusedCL - 1))
          *reinterpret_cast<float4*>(hb + (size_t)b * HP + 4 * c4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  float* own = hb + (size_t)(-fv.lo) * HP;  // own rows of the h1 / r buffer; also the landing zone of raw x_t
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    constexpr uint32_t nb = RB * W * 4;
    mbar_expect_tx(bar, nb);
    bulk_g2s(own, x + gbase, nb, bar);  // 32 KB contiguous (pitch W) inside the 33 KB own-row region
  }
  PSX_TICK(1)
  cluster.sync();  // #0: every CTA of the cluster is running => remote shared memory may be written from now on
  float* up = rank > 0 ? cluster.map_shared_rank(hb, rank - 1) : nullptr;
  float* dn = rank < kFusedCL - 1 ? cluster.map_shared_rank(hb, rank + 1) : nullptr;

  PSX_TICK(2)
  // ---- P2: Tweedie + interleave (row pairs (2rp, 2rp+1)) -> comp; eps straight from global
  {
    float4 ea[4], eb[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int idx = threadIdx.x + t * kThreads;
      const int rp = idx >> 6, c4 = idx & 63;
      ea[t] = ld_stream4(eps + gbase + (2 * rp) * W + 4 * c4);
      eb[t] = ld_stream4(eps + gbase + (2 * rp + 1) * W + 4 * c4);
    }
    mbar_wait(bar, 0);
#pragma unroll
    for (int t = 0; t < 4; ++t) {  // 16 row pairs x 64 column quads = 4 * 256
      const int idx = threadIdx.x + t * kThreads;
      const int rp = idx >> 6, c4 = idx & 63;
      float4 a = *reinterpret_cast<const float4*>(own + (2 * rp) * W + 4 * c4);
      float4 b = *reinterpret_cast<const float4*>(own + (2 * rp + 1) * W + 4 * c4);
      a.x = tweedie(a.x, ea[t].x, tc); a.y = tweedie(a.y, ea[t].y, tc);
      a.z = tweedie(a.z, ea[t].z, tc); a.w = tweedie(a.w, ea[t].w, tc);
      b.x = tweedie(b.x, eb[t].x, tc); b.y = tweedie(b.y, eb[t].y, tc);
      b.z = tweedie(b.z, eb[t].z, tc); b.w = tweedie(b.w, eb[t].w, tc);
      float2* dst = comp + rp * pitch2 - fh.lo + 4 * c4;
      *reinterpret_cast<float4*>(dst) = make_float4(a.x, b.x, a.y, b.y);
      *reinterpret_cast<float4*>(dst + 2) = make_float4(a.z, b.z, a.w, b.w);
    }
  }
  PSX_TICK(3)
  __syncthreads();  // comp ready; raw x_t dead => own rows reusable

  // own row i (0..31) is also row i + RB - lo of the upper neighbour's buffer when i < K + lo ... see below:
  //   upper neighbour (rank-1) bottom halo: buffer row i + RB - lo   for i in [0, K + lo)
  //   lower neighbour (rank+1) top halo   : buffer row i - RB - lo   for i in [RB + lo, RB)
  // ---- P3: H pass (16 row pairs x 16 column groups of 16) -> h1 into own rows, boundary rows pushed next door
  const int hrp = threadIdx.x & 15, hcg = threadIdx.x >> 4;
  {
    float2 acc[16];
    row_block16<K>(acc, comp + hrp * pitch2 + 16 * hcg, fh);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * hrp + h;
      float4 v[4];
#pragma unroll
      for (int m = 0; m < 4; ++m)
        v[m] = make_float4(h ? acc[4 * m].y : acc[4 * m].x, h ? acc[4 * m + 1].y : acc[4 * m + 1].x,
                           h ? acc[4 * m + 2].y : acc[4 * m + 2].x, h ? acc[4 * m + 3].y : acc[4 * m + 3].x);
      float* dst = hb + (size_t)(-fv.lo + i) * HP + 16 * hcg;
#pragma unroll
      for (int m = 0; m < 4; ++m) *reinterpret_cast<float4*>(dst + 4 * m) = v[m];
      if (up && i < K + fv.lo) {
        float* rd = up + (size_t)(i + RB - fv.lo) * HP + 16 * hcg;
#pragma unroll
        for (int m = 0; m < 4; ++m) *reinterpret_cast<float4*>(rd + 4 * m) = v[m];
      }
      if (dn && i >= RB + fv.lo) {
        float* rd = dn + (size_t)(i - RB - fv.lo) * HP + 16 * hcg;
#pragma unroll
        for (int m = 0; m < 4; ++m) *reinterpret_cast<float4*>(rd + 4 * m) = v[m];
      }
    }
  }
  PSX_TICK(4)
  cluster.sync();  // #1: own h1 rows and both halos (pushed by the neighbours) are in place

  PSX_TICK(5)
  // ---- P5: V pass, r = y - V h1, |r|^2 (128 column pairs x 2 groups of 16 rows)
  const int vcp = threadIdx.x & 127, vg = threadIdx.x >> 7;
  float e2 = 0.f;
  float2 r[16];
  {
    col_block16<K, HP>(r, hb + (size_t)(16 * vg) * HP + 2 * vcp, fv);
    const float* yp = y + ((pl / C) / obs_repeat * C + pl % C) * (int64_t)H * W +
                      (int64_t)(rank * RB + 16 * vg) * W + 2 * vcp;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float2 yv = __ldg(reinterpret_cast<const float2*>(yp + (int64_t)j * W));
      r[j].x = __fsub_rn(yv.x, r[j].x);
      r[j].y = __fsub_rn(yv.y, r[j].y);
      e2 = fmaf(r[j].x, r[j].x, e2);
      e2 = fmaf(r[j].y, r[j].y, e2);
    }
  }
  PSX_TICK(6)
  cluster.sync();  // #2: nobody (here or next door) still reads h1
  PSX_TICK(7)
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int i = 16 * vg + j;
    *reinterpret_cast<float2*>(hb + (size_t)(-av.lo + i) * HP + 2 * vcp) = r[j];
    if (up && i < K + av.lo) *reinterpret_cast<float2*>(up + (size_t)(i + RB - av.lo) * HP + 2 * vcp) = r[j];
    if (dn && i >= RB + av.lo) *reinterpret_cast<float2*>(dn + (size_t)(i - RB - av.lo) * HP + 2 * vcp) = r[j];
  }
  PSX_TICK(8)
  cluster.sync();  // #3: every CTA's r rows and halos are in place

  PSX_TICK(9)
  // ---- P7: V^T pass -> h2 into comp (interleaved)
  {
    col_block16<K, HP>(r, hb + (size_t)(16 * vg) * HP + 2 * vcp, av);
#pragma unroll
    for (int m = 0; m < 8; ++m)
      *reinterpret_cast<float4*>(comp + (8 * vg + m) * pitch2 - ah.lo + 2 * vcp) =
          make_float4(r[2 * m].x, r[2 * m + 1].x, r[2 * m].y, r[2 * m + 1].y);
  }
  __syncthreads();

  PSX_TICK(10)
  // ---- P8: H^T pass -> cot
  {
    float2 acc[16];
    row_block16<K>(acc, comp + hrp * pitch2 + 16 * hcg, ah);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float* dst = cot + gbase + (int64_t)(2 * hrp + h) * W + 16 * hcg;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        float4 v;
        v.x = __fmul_rn(coef, h ? acc[4 * m].y : acc[4 * m].x);
        v.y = __fmul_rn(coef, h ? acc[4 * m + 1].y : acc[4 * m + 1].x);
        v.z = __fmul_rn(coef, h ? acc[4 * m + 2].y : acc[4 * m + 2].x);
        v.w = __fmul_rn(coef, h ? acc[4 * m + 3].y : acc[4 * m + 3].x);
        st_stream4(dst + 4 * m, v);
      }
    }
  }
  PSX_TICK(11)
  const float tot = block_sum(e2, red);
  if (threadIdx.x == 0) {
    const int64_t l = pl / C;
    const int ch = (int)(pl % C);
    err_part[l * (int64_t)(C * kFusedCL) + ch * kFusedCL + rank] = tot;
  }
  PSX_TICK(12)
  cluster.sync();  // no CTA may exit while a neighbour could still read its shared memory
  PSX_TICK(13)
}

// host: 2-D tensor map over the tall (planes*H) x W fp32 matrix, box = 32 columns x H rows
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
static bool make_strip_map(CUtensorMap* map, const float* base, int64_t rows, int W, int box_w, int box_h) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)W * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------------ planning
static int row_tw(int W) {
  int tw = (W + 31) & ~31;
  return tw > 256 ? 256 : tw;
}
static int row_pitch2(int in_w) {
  int p = in_w;                  // multiple of 8, in float2 units
  while ((p & 15) != 2) p += 2;  // pitch2 % 16 == 2 -> LDS.128 with lane <-> row pair is conflict-free
  return p;
}
static size_t cols_smem(const psx_op* op, int TC, bool residual, const Taps& tf, const Taps& ta) {
  const int H8 = (op->H + 7) & ~7;
  size_t rows = (size_t)H8 + tf.k + (residual ? (size_t)H8 + ta.k : 0);
  return rows * TC * sizeof(float);
}

int sepblur_plan(psx_op* op) {
  // widest column strip whose fused (two-buffer) footprint fits in shared memory
  const size_t limit = 200 * 1024;
  op->col_tc = 0;
  for (int tc : {32, 16, 8}) {
    if (cols_smem(op, tc, true, op->fv, op->av) <= limit) { op->col_tc = tc; break; }
  }
  // planes taller / wider than 256 (up to 512) run the strip kernels with 16-column strips (launch_pre_sepblur)
  if ((op->H > 256 || op->W > 256) && op->H <= 512 && op->W <= 512 && op->H % 32 == 0 && op->W % 32 == 0 &&
      op->col_tc >= 16)
    op->col_tc = 16;
  if (!op->col_tc) return fail(PSX_ERR_UNSUPPORTED, "separable blur: image too tall for the column kernel");
  op->err_parts = op->C * ceil_div(op->W, op->col_tc);
  return PSX_OK;
}

int sm_count();

template <int MODE>
static int run_rows(const psx_op* op, const Taps& t, const float* in, const float* eps, float* out,
                    int64_t planes, float sa, float s1, float w, const float* dsc, cudaStream_t st, bool half = false) {
  if (half) {  // bf16 inputs: the pipelined Tweedie kernel with the two instantiated tap counts only
    if (MODE != ROWS_TWEEDIE || (op->W & 7) != 0 || op->W > 512 || (t.k != 40 && t.k != 16))
      return fail(PSX_ERR_UNSUPPORTED, "blur K1 on a bf16 state: unsupported geometry");
    const int pitch2 = row_pitch2(op->W + t.k);
    const size_t smem = kPipeHdr + (size_t)(kPipeRows / 2) * pitch2 * sizeof(float2) +
                        (size_t)2 * 2 * kPipeRows * op->W * 2;
    const int64_t total_rows = planes * op->H;
    const int64_t num_tiles = (total_rows + kPipeRows - 1) / kPipeRows;
    const float coef = (float)((double)w / (double)sa);
#define PSX_ROWS_H(KK)                                                                                                   \
  {                                                                                                                      \
    static int occ = 0;                                                                                                  \
    if (!occ) {                                                                                                          \
      cudaFuncSetAttribute(conv_rows_pipe<ROWS_TWEEDIE, KK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_rows_pipe<ROWS_TWEEDIE, KK, true>, kThreads, smem) !=  \
              cudaSuccess || occ < 1)                                                                                    \
        occ = 1;                                                                                                         \
    }                                                                                                                    \
    int64_t grid = (int64_t)occ * sm_count();                                                                            \
    if (grid > num_tiles) grid = num_tiles;                                                                              \
    conv_rows_pipe<ROWS_TWEEDIE, KK, true><<<(unsigned)grid, kThreads, smem, st>>>(in, eps, out, total_rows, op->W,      \
                                                                                 pitch2, num_tiles, t, sa, s1, coef, dsc); \
  }
    if (t.k == 40) PSX_ROWS_H(40) else PSX_ROWS_H(16)
#undef PSX_ROWS_H
    return check_cuda(cudaGetLastError(), "conv_rows_pipe (bf16) launch");
  }
  if ((op->W & 7) == 0 && op->W <= 512 && !env_opts().no_pipe) {
    const int pitch2 = row_pitch2(op->W + t.k);
    const int arrays = MODE == ROWS_TWEEDIE ? 2 : 1;
    const size_t smem = kPipeHdr + (size_t)(kPipeRows / 2) * pitch2 * sizeof(float2) +
                        (size_t)2 * arrays * kPipeRows * op->W * sizeof(float);  // raw stages double as read slack
    const int64_t total_rows = planes * op->H;
    const int64_t num_tiles = (total_rows + kPipeRows - 1) / kPipeRows;
    const float coef = (float)((double)w / (double)sa);
#define PSX_ROWS(KK)                                                                                           \
  {                                                                                                            \
    static int occ = 0;                                                                                        \
    if (!occ) {                                                                                                \
      cudaFuncSetAttribute(conv_rows_pipe<MODE, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_rows_pipe<MODE, KK>, kThreads, smem) !=     \
              cudaSuccess || occ < 1)                                                                          \
        occ = 1;                                                                                               \
    }                                                                                                          \
    int64_t grid = (int64_t)occ * sm_count();                                                                  \
    if (grid > num_tiles) grid = num_tiles;                                                                    \
    conv_rows_pipe<MODE, KK><<<(unsigned)grid, kThreads, smem, st>>>(in, eps, out, total_rows, op->W, pitch2,  \
                                                                     num_tiles, t, sa, s1, coef, dsc);         \
  }
    // the occupancy cache is per instantiation and assumes one (W, taps) geometry per process and mode;
    // other geometries only change smem by a few KB and keep the same CTA count per SM in practice
    if (t.k == 40) PSX_ROWS(40) else if (t.k == 16) PSX_ROWS(16) else if (t.k == 64) PSX_ROWS(64) else PSX_ROWS(0)
#undef PSX_ROWS
    return check_cuda(cudaGetLastError(), "conv_rows_pipe launch");
  }
  const int TW = row_tw(op->W);
  const int pitch = row_pitch2(TW + t.k);
  const size_t smem = (size_t)(kRowTH / 2) * pitch * sizeof(float2);
  const float coef = (float)((double)w / (double)sa);  // cot = (w / sa) * H^T h2, one rounding
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(conv_rows<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_done = true;
  }
  dim3 grid(ceil_div(op->W, TW), ceil_div(op->H, kRowTH), (unsigned)planes);
  conv_rows<MODE><<<grid, kThreads, smem, st>>>(in, eps, out, op->H, op->W, TW, pitch, t, sa, s1, coef, dsc);
  return check_cuda(cudaGetLastError(), "conv_rows launch");
}

template <bool RESIDUAL>
static int run_cols(const psx_op* op, const Taps& tf, const Taps& ta, const float* in, const float* y,
                    float* out, float* err_part, int64_t planes, int64_t obs_repeat, cudaStream_t st) {
  {
    const size_t stage = ((size_t)op->H + tf.k) * kColTC * sizeof(float);
    const size_t smem_pipe = kPipeHdr + 2 * stage + (RESIDUAL ? ((size_t)op->H + ta.k) * kColTC * sizeof(float) : 0) +
                             8 * kColTC * sizeof(float);  // 8 slack rows for the unconditional window prefetch
    CUtensorMap map;
    // the |r|^2 partials are laid out per strip of op->col_tc columns: the pipelined kernel only runs at that width
    if ((op->W % kColTC) == 0 && (op->H % 16) == 0 && op->H <= 256 && op->col_tc == kColTC && !env_opts().no_pipe &&
        make_strip_map(&map, in, planes * op->H, op->W, kColTC, op->H)) {
      const int rounds = ((op->H >> 3) * (kColTC >> 1) + kThreads - 1) / kThreads;  // the last round may be partial
      const int strips = op->W / kColTC;
      const int64_t num_tiles = planes * strips;
#define PSX_COLS(R, KK)                                                                                      \
  {                                                                                                          \
    static int occ = 0;                                                                                      \
    if (!occ) {                                                                                              \
      cudaFuncSetAttribute(conv_cols_pipe<RESIDUAL, R, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                           200 * 1024);                                                                      \
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_cols_pipe<RESIDUAL, R, KK>, kThreads,     \
                                                        smem_pipe) != cudaSuccess || occ < 1)                \
        occ = 1;                                                                                             \
    }                                                                                                        \
    int64_t grid = (int64_t)occ * sm_count();                                                                \
    if (grid > num_tiles) grid = num_tiles;                                                                  \
    conv_cols_pipe<RESIDUAL, R, KK><<<(unsigned)grid, kThreads, smem_pipe, st>>>(                            \
        map, y, out, err_part, op->C, op->H, op->W, strips, num_tiles, obs_repeat, tf, ta);                  \
  }
      const int kk = (tf.k == ta.k || !RESIDUAL) ? tf.k : 0;
      if (rounds == 2) {
        if (kk == 40) PSX_COLS(2, 40) else if (kk == 16) PSX_COLS(2, 16) else if (kk == 64) PSX_COLS(2, 64) else PSX_COLS(2, 0)
      } else if (rounds == 1) {
        if (kk == 40) PSX_COLS(1, 40) else if (kk == 16) PSX_COLS(1, 16) else if (kk == 64) PSX_COLS(1, 64) else PSX_COLS(1, 0)
      }
      if (rounds == 1 || rounds == 2) return check_cuda(cudaGetLastError(), "conv_cols_pipe launch");
#undef PSX_COLS
    }
  }
  const int TC = op->col_tc;
  const size_t smem = cols_smem(op, TC, RESIDUAL, tf, ta);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(conv_cols<RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  dim3 grid(ceil_div(op->W, TC), (unsigned)planes);
  conv_cols<RESIDUAL><<<grid, kThreads, smem, st>>>(in, y, out, err_part, op->C, op->H, op->W, TC,
                                                    obs_repeat, tf, ta);
  return check_cuda(cudaGetLastError(), "conv_cols launch");
}

// cols16 + rows_il tail of K1 (h1 in ws, row-major)  ->  cot.  Returns -1 when the geometry does not qualify.
template <int K, int TC>
static int run_fast_tail(const psx_op* op, const float* y, float* ws, float* ws2, float* cot, float* err_part, int64_t planes,
                         int64_t sample0, int64_t obs_repeat, float w, float sa, const float* dsc, cudaStream_t st,
                         bool half = false) {
  const int W = op->W, H = op->H;
  CUtensorMap map;
  const int box_h = H > 256 ? H / 2 : H;  // TMA boxes hold at most 256 rows
  if (!make_strip_map(&map, ws, planes * H, W, TC, box_h)) return -1;
  constexpr int ROUNDS = TC == 32 ? 1 : 2;  // TC = 16 serves planes up to 512 wide: 256 row tasks per 16-row tile
  const size_t smem_c = kPipeHdr + 3 * ((size_t)H + K) * TC * sizeof(float);
  const int pitch2 = row_pitch2(W + K);
  const size_t smem_r = kPipeHdr + (size_t)kIlStages * (kPipeRows / 2) * pitch2 * sizeof(float2);
  static int occ_c = 0, occ_r = 0;
  if (!occ_c) {
    cudaFuncSetAttribute(conv_cols16<K, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_rows_il<K, ROUNDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, conv_cols16<K, TC>, kThreads, smem_c) != cudaSuccess || occ_c < 1)
      occ_c = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_r, conv_rows_il<K, ROUNDS>, kIlThreads, smem_r) != cudaSuccess || occ_r < 1)
      occ_r = 1;
  }
  const int strips = W / TC;
  const int64_t tiles_c = planes * strips;
  int64_t grid_c = (int64_t)occ_c * sm_count();
  if (grid_c > tiles_c) grid_c = tiles_c;
  // NOT in place: the interleaved h2 of (row pair, column c) sits at offset 2 c of the double row, i.e. on cells of
  // two OTHER strips' h1 -- which their (persistent, unsynchronised) CTAs may not have fetched yet.  h2 -> ws2.
  conv_cols16<K, TC><<<(unsigned)grid_c, kThreads, smem_c, st>>>(map, y, ws2, err_part, op->C, H, W, strips, tiles_c,
                                                             obs_repeat, sample0, box_h, op->fv, op->av);
  int rc = check_cuda(cudaGetLastError(), "conv_cols16 launch");
  if (rc) return rc;
  const int64_t total_rows = planes * H;
  const int64_t tiles_r = (total_rows + kPipeRows - 1) / kPipeRows;
  int64_t grid_r = (int64_t)occ_r * sm_count();
  if (grid_r > tiles_r) grid_r = tiles_r;
  const float coef = (float)((double)w / (double)sa);
  if (half) {  // bf16 cotangent out; same shared-memory footprint, so the occupancy figure carries over
    static bool attr_h = false;
    if (!attr_h) {
      cudaFuncSetAttribute(conv_rows_il<K, ROUNDS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_h = true;
    }
    conv_rows_il<K, ROUNDS, true><<<(unsigned)grid_r, kIlThreads, smem_r, st>>>(ws2, cot, total_rows, W, pitch2, tiles_r,
                                                                                op->ah, coef, dsc);
  } else {
    conv_rows_il<K, ROUNDS><<<(unsigned)grid_r, kIlThreads, smem_r, st>>>(ws2, cot, total_rows, W, pitch2, tiles_r, op->ah, coef, dsc);
  }
  return check_cuda(cudaGetLastError(), "conv_rows_il launch");
}

// K1 launches that can emit the bridge mean for K2 (psx_dps_pre_mean): the tensor-core blur at batches that leave
// SMs idle
bool fuses_mean(const psx_op* op, int64_t L) {
  return op->kind == PSX_OP_SEPBLUR && tcblur_available(op) && !env_opts().no_tc && tcblur_mean_fits(op, L);
}

// half: x, eps and cot are bf16 arrays (passed through the float* parameters); the intermediates in ws stay fp32.
int launch_pre_sepblur(const psx_op* op, const float* x, const float* eps, const float* y, int64_t L,
                       int64_t obs_repeat, float sa, float s1, float w, const float* dsc, float* cot, float* err_part,
                       float* x0_out, float* ws, cudaStream_t st, bool half, float* mean_out, float c_ell,
                       float c_s, const float* zn, float sd) {
  if (x0_out) return fail(PSX_ERR_UNSUPPORTED, "psx_dps_pre: d_x0_out is not produced for blur operators");
  const int64_t planes = L * op->C;
  // tensor-core single launch (psx_tcblur.cu): 256 x 256 planes, symmetric taps shared by rows and columns
  if (mean_out && (half || !fuses_mean(op, L)))
    return fail(PSX_ERR_UNSUPPORTED, "psx_dps_pre_mean: this operator's K1 does not emit the bridge mean at this "
                                     "batch size (psx_op_fuses_mean)");
  if (!half && tcblur_available(op) && !env_opts().no_tc)
    return launch_pre_sepblur_tc(op, x, eps, y, L, obs_repeat, sa, s1, w, dsc, cot, err_part, mean_out, c_ell, c_s,
                                 zn, sd, st);

  // cluster-fused single launch for 256 x 256 planes
  if (op->H == kFusedCL * kFusedRB && op->W == kFusedW && op->fh.k == 40 && op->fv.k == 40 && op->ah.k == 40 &&
      op->av.k == 40 && op->fh.lo == op->ah.lo && op->fv.lo == op->av.lo && -op->fv.lo <= kFusedRB &&
      op->err_parts == op->C * kFusedCL && env_opts().fused && !half) {
    // opt-in: measured 48.0 us vs 46.1 us for the three-launch path at L = 16 (profiles/README.md) -- each CTA runs
    // its nine phases back to back (21 us per CTA, 6 us of it in cluster barriers) and 48 planes need two waves
    // of the 33 clusters that fit, so the single launch does not pay off yet despite its minimal HBM traffic.
    constexpr int K = 40;
    const int pitch2 = row_pitch2(kFusedW + K);
    const size_t smem = kPipeHdr + (size_t)(kFusedRB / 2) * pitch2 * sizeof(float2) +
                        (size_t)(kFusedRB + K) * kFusedHP * sizeof(float);
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(blur_k1_fused<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr = true;
    }
    const float coef = (float)((double)w / (double)sa);
    blur_k1_fused<K><<<(unsigned)(planes * kFusedCL), kThreads, smem, st>>>(
        x, eps, y, cot, err_part, op->C, obs_repeat, pitch2, op->fh, op->fv, op->av, op->ah, sa, s1, coef, dsc);
    return check_cuda(cudaGetLastError(), "blur_k1_fused launch");
  }
  const int kk = op->fv.k;
  // planes up to 256 x 256: 32-column strips; up to 512 x 512 (configs 4-5): 16-column strips, two TMA boxes per strip
  const bool small = op->W <= 256 && op->H <= 256;
  const bool fast = op->W % 32 == 0 && op->W <= 512 && op->H % 16 == 0 && op->H <= 512 && op->av.k == kk &&
                    op->ah.k == kk && op->col_tc == (small ? 32 : 16) && (kk == 40 || kk == 16) &&
                    (small || op->H % 32 == 0) && !env_opts().no_pipe && !env_opts().no_fast16 &&
                    encode_fn() != nullptr;
  if (fast) {
    // The three launches of one group leave SMs idle in every kernel's ramp-up and tail (about a third of K1 at
    // L = 16, tools/micro/k1_trace_main.cu).  Samples are independent, so K1 runs as `parts` groups on the caller's
    // stream plus side streams: one group's tails and launch gaps are filled by the other groups' kernels.
    // Measured (profiles/README.md): inside a replayed CUDA graph the fork/join is free and two groups cut K1 by 8 %
    // at L = 16; launched eagerly the extra event calls cost more than the overlap returns until L >= 32.
    const int forced = env_opts().split;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    const int want = forced ? forced : ((cap == cudaStreamCaptureStatusActive || L >= 32) ? 2 : 1);
    int parts = want < 1 ? 1 : (want > op->aux_n + 1 ? op->aux_n + 1 : want);
    while (parts > 1 && (L % parts != 0 || L / parts < 2)) --parts;
    const int64_t Lp = L / parts, planes_p = Lp * op->C;
    std::unique_lock<std::mutex> lock;
    if (parts > 1) {
      lock = std::unique_lock<std::mutex>(*op->aux_mu);
      if (int rc = check_cuda(cudaEventRecord(op->ev_fork, st), "K1 fork event")) return rc;
    }
    for (int p = 0; p < parts; ++p) {
      cudaStream_t s = p == 0 ? st : op->aux_stream[p - 1];
      int rc = p > 0 ? check_cuda(cudaStreamWaitEvent(s, op->ev_fork, 0), "K1 fork wait") : PSX_OK;
      const int64_t l0 = p * Lp, off = l0 * op->n;
      const int64_t off_io = half ? off / 2 : off;  // x / eps / cot offsets in float units (n is even for these shapes)
      if (!rc) rc = run_rows<ROWS_TWEEDIE>(op, op->fh, x + off_io, eps + off_io, ws + off, planes_p, sa, s1, w, dsc, s, half);
      float* part = err_part + l0 * op->err_parts;
#define PSX_TAIL(KK, TCC) \
  run_fast_tail<KK, TCC>(op, y, ws + off, ws + L * op->n + off, cot + off_io, part, planes_p, l0, obs_repeat, w, sa, dsc, s, half)
      if (!rc) {
        rc = small ? (kk == 40 ? PSX_TAIL(40, 32) : PSX_TAIL(16, 32)) : (kk == 40 ? PSX_TAIL(40, 16) : PSX_TAIL(16, 16));
        if (rc < 0) rc = fail(PSX_ERR_CUDA, "blur K1: could not encode the strip tensor map");
      }
#undef PSX_TAIL
      if (p > 0) {
        // joined even after a failure: whatever reached the side stream must not be left unordered against the
        // caller's stream (under capture an unjoined fork invalidates the whole graph with an opaque error)
        const int e1 = check_cuda(cudaEventRecord(op->ev_join[p - 1], s), "K1 join event");
        const int e2 = e1 ? e1 : check_cuda(cudaStreamWaitEvent(st, op->ev_join[p - 1], 0), "K1 join wait");
        if (!rc) rc = e2;
      }
      if (rc) return rc;  // the groups after this one have not been forked
    }
    return PSX_OK;
  }
  if (half) return fail(PSX_ERR_UNSUPPORTED, "blur K1 on a bf16 state needs the strip-kernel geometry (W % 32 == 0, H % 16 == 0, <= 512)");
  int rc = run_rows<ROWS_TWEEDIE>(op, op->fh, x, eps, ws, planes, sa, s1, w, dsc, st);
  if (rc) return rc;
  rc = run_cols<true>(op, op->fv, op->av, ws, y, ws, err_part, planes, obs_repeat, st);
  if (rc) return rc;
  return run_rows<ROWS_COT>(op, op->ah, ws, nullptr, cot, planes, sa, s1, w, dsc, st);
}

// ------------------------------------------------------------------------------------------ sparse 2-D
enum C2Mode { C2_PLAIN = 0, C2_RESIDUAL = 1, C2_COT = 2 };
constexpr int kC2TH = 64;        // output tile: 64 rows x 64 columns, 256 threads (32 / 48 / 128 rows measured: 141 / 158 / 136 us against 136 us at L = 16, 473 / 503 / 440 against 426 us at L = 64)
constexpr int kC2TW = 64;
constexpr int kC2Rows = 2;       // rows per thread: rp, rp + 16 (4 rows per thread measured slower: 194 vs 153 us)
constexpr int kC2Threads = (kC2TH / kC2Rows) * (kC2TW / 8);

// One chunk of a row segment: 4 taps x (8 outputs x kC2Rows rows).  The 12-register windows rotate by 4 per chunk; R is
// the chunk index modulo 3, so every register index is a compile-time constant and no value is ever moved.
template <int R>
__device__ __forceinline__ void c2_chunk(float (&acc)[kC2Rows][8], float (&win)[kC2Rows][12], const float* __restrict__ row,
                                         int row_step, const float4* __restrict__ w4) {
  constexpr int S = (8 + 4 * R) % 12;  // slot of the incoming columns
#pragma unroll
  for (int h = 0; h < kC2Rows; ++h) {
    const float4 v = *reinterpret_cast<const float4*>(row + h * row_step);
    win[h][S] = v.x; win[h][S + 1] = v.y; win[h][S + 2] = v.z; win[h][S + 3] = v.w;
  }
  const float4 t = *w4;
  const float tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int h = 0; h < kC2Rows; ++h) acc[h][q] = fmaf(tw[i], win[h][(q + i + 4 * R) % 12], acc[h][q]);
}

// The same chunk on 32-bit shared-memory addresses (tile row and tap group both in shared memory): explicit ld.shared,
// so that no generic-to-shared window arithmetic is re-derived per segment.
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
template <int R>
__device__ __forceinline__ void c2_chunk_s(float (&acc)[kC2Rows][8], float (&win)[kC2Rows][12], uint32_t row,
                                           uint32_t row_step_b, uint32_t w4) {
  constexpr int S = (8 + 4 * R) % 12;  // slot of the incoming columns
#pragma unroll
  for (int h = 0; h < kC2Rows; ++h) {
    const float4 v = lds128(row + h * row_step_b);
    win[h][S] = v.x; win[h][S + 1] = v.y; win[h][S + 2] = v.z; win[h][S + 3] = v.w;
  }
  const float4 t = lds128(w4);
  const float tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int h = 0; h < kC2Rows; ++h) acc[h][q] = fmaf(tw[i], win[h][(q + i + 4 * R) % 12], acc[h][q]);
}

// 2-D (motion) PSFs.  The PSF is a list of row segments (psx_common.cuh: RowSeg): every PSF row is a short 1-D
// correlation, evaluated as a sliding register window -- per chunk of 4 taps a thread issues 2 LDS.128 + one
// broadcast tap load for 64 FMAs (8 outputs x 2 rows), against one load per FMA of a tap-by-tap gather.
// grid = (tilesX, tilesY, planes).  Thread q: column group cg (8 columns) and rows rp, rp + 16 of the 32 x 64 tile;
// a quarter-warp covers 4 column groups x 2 consecutive rows, conflict-free for LDS.128 because pitch / 4 is odd.
// STAB: the taps and segments were copied behind the tile (nw4 > 0): they are then read with shared-memory loads from
// compile-time-known address space (a pointer that may be either costs generic LD.E plus a 16-bit load per field).
template <int MODE, bool STAB>
__global__ void __launch_bounds__(kC2Threads)
conv2d_rowseg(const float* __restrict__ in, const float* __restrict__ eps, const float* __restrict__ y,
              float* __restrict__ out, float* __restrict__ err_part, const RowSeg* __restrict__ segs,
              const float4* __restrict__ w4, int nseg, int nw4, int dy_lo, int dy_hi, int dx_lo, int dx_hi, int pitch,
              int C, int H, int W, int64_t obs_repeat, float sa, float s1, float wgt, const float* __restrict__ dsc) {
  step_scalars_k1(dsc, sa, s1, wgt);
  const TweedieC tc = make_tc(s1, sa);
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[32];
  const int th = kC2TH + dy_hi - dy_lo, tw = kC2TW + dx_hi - dx_lo;  // tw % 4 == 0, tw <= pitch
  const int r0 = blockIdx.y * kC2TH, c0 = blockIdx.x * kC2TW;
  const int64_t pl = blockIdx.z;
  const int64_t plane = pl * H * W;
  // nw4 > 0: the launcher reserved room behind the tile for a copy of the taps and segments -- shared-memory
  // broadcasts (~30 cycles) instead of L1-hit global loads (~200) at the head of every chunk's FMA chain
  float4* const sw4 = reinterpret_cast<float4*>(smem + th * pitch);
  RowSeg* const ssg = reinterpret_cast<RowSeg*>(sw4 + (STAB ? nw4 : 0));
  if (STAB) {
    for (int i = threadIdx.x; i < nw4; i += kC2Threads) sw4[i] = w4[i];
    for (int i = threadIdx.x; i < nseg; i += kC2Threads) ssg[i] = segs[i];
  }

  if ((W & 3) == 0) {
    // vector fill: a warp owns tile rows warp, warp + 4, ...; c0 + dx_lo and W are multiples of 4, so a float4 is
    // either inside the image or outside; rows are unrolled so that several rows of loads are in flight per lane
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tw4 = tw >> 2;
#pragma unroll 4
    for (int r = warp; r < th; r += kC2Threads / 32) {
      const int gr = r0 + dy_lo + r;
      const bool row_ok = gr >= 0 && gr < H;
      for (int c4 = lane; c4 < tw4; c4 += 32) {
        const int gc = c0 + dx_lo + 4 * c4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_ok && gc >= 0 && gc < W) {
          const int64_t g = plane + (int64_t)gr * W + gc;
          v = *reinterpret_cast<const float4*>(in + g);
          if (MODE == C2_RESIDUAL) {
            const float4 e = *reinterpret_cast<const float4*>(eps + g);
            v.x = tweedie(v.x, e.x, tc); v.y = tweedie(v.y, e.y, tc);
            v.z = tweedie(v.z, e.z, tc); v.w = tweedie(v.w, e.w, tc);
          }
        }
        *reinterpret_cast<float4*>(smem + r * pitch + 4 * c4) = v;
      }
    }
  } else {
    for (int idx = threadIdx.x; idx < th * tw; idx += kC2Threads) {
      const int r = idx / tw, c = idx - r * tw;
      const int gr = r0 + dy_lo + r, gc = c0 + dx_lo + c;
      float v = 0.f;
      if (gr >= 0 && gr < H && gc >= 0 && gc < W) {
        v = in[plane + (int64_t)gr * W + gc];
        if (MODE == C2_RESIDUAL) v = tweedie(v, eps[plane + (int64_t)gr * W + gc], tc);
      }
      smem[r * pitch + c] = v;
    }
  }
  __syncthreads();

  const int q = threadIdx.x;
  // q bits 0-1 and 3.. (kCgBits - 2 of them): column group; bit 2 and the rest: row pair
  constexpr int kCgBits = kC2TW == 64 ? 3 : kC2TW == 128 ? 4 : kC2TW == 256 ? 5 : -1;
  static_assert(kCgBits > 0, "tile width 64, 128 or 256");
  const int cg = (q & 3) | (((q >> 3) & ((1 << (kCgBits - 2)) - 1)) << 2);
  const int rp = ((q >> 2) & 1) | ((q >> (kCgBits + 1)) << 1);  // 0 .. kC2TH / kC2Rows - 1
  constexpr int kStepRows = kC2TH / kC2Rows;
  const int row_step = kStepRows * pitch;
  float acc[kC2Rows][8];
#pragma unroll
  for (int h = 0; h < kC2Rows; ++h)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[h][j] = 0.f;
  // The thread's tile offset takes a round trip through shared memory: left alone, ptxas re-derives it from %tid for
  // every segment (S2R + 10 integer instructions of a ~50-instruction segment prologue; ncu source view), and an
  // empty asm does not stop it.
  int boff = (rp - dy_lo) * pitch + (8 * cg - dx_lo);
  // (the same for the three shared-memory base addresses of the loop, which would otherwise be rebuilt per segment from
  // %cluster_ctaid and the shared window base on the uniform datapath: ~25 instructions)
  uint32_t base_s = (uint32_t)__cvta_generic_to_shared(smem) + 4u * (uint32_t)boff;
  uint32_t wtab_s = (uint32_t)__cvta_generic_to_shared(sw4), seg_s = (uint32_t)__cvta_generic_to_shared(ssg);
  {
    __shared__ int spin[4 * kC2Threads];
    volatile int* vb = spin;
    vb[q] = boff;
    vb[q + kC2Threads] = (int)base_s;
    vb[q + 2 * kC2Threads] = (int)wtab_s;
    vb[q + 3 * kC2Threads] = (int)seg_s;
    boff = vb[q];
    base_s = (uint32_t)vb[q + kC2Threads];
    wtab_s = (uint32_t)vb[q + 2 * kC2Threads];
    seg_s = (uint32_t)vb[q + 3 * kC2Threads];
  }
  static_assert(sizeof(RowSeg) == 8, "a segment is read as one 64-bit word");  // {dy | dx0 << 16, nch | w4_off << 16}
  if (STAB) {
    // everything the loop touches is in shared memory: 32-bit shared addresses, explicit ld.shared
    const uint32_t row_step_b = 4u * (uint32_t)row_step;
    int2 nxt;
    asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(nxt.x), "=r"(nxt.y) : "r"(seg_s));
    for (int sgi = 0; sgi < nseg; ++sgi) {
      const int2 sg = nxt;
      if (sgi + 1 < nseg)  // in flight while this segment computes
        asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(nxt.x), "=r"(nxt.y) : "r"(seg_s + 8u * (uint32_t)(sgi + 1)));
      const int sg_dy = (int)(short)(sg.x & 0xffff), sg_dx0 = sg.x >> 16;
      uint32_t row = base_s + 4u * (uint32_t)(sg_dy * pitch + sg_dx0);
      uint32_t wp = wtab_s + 16u * ((uint32_t)sg.y >> 16);
      float win[kC2Rows][12];
#pragma unroll
      for (int h = 0; h < kC2Rows; ++h) {
        const float4 a0 = lds128(row + h * row_step_b), a1 = lds128(row + h * row_step_b + 16);
        win[h][0] = a0.x; win[h][1] = a0.y; win[h][2] = a0.z; win[h][3] = a0.w;
        win[h][4] = a1.x; win[h][5] = a1.y; win[h][6] = a1.z; win[h][7] = a1.w;
      }
      row += 32;
      int nch = (int)(short)(sg.y & 0xffff);
      for (; nch >= 3; nch -= 3, row += 48, wp += 48) {  // the window rotation has period 3
        c2_chunk_s<0>(acc, win, row, row_step_b, wp);
        c2_chunk_s<1>(acc, win, row + 16, row_step_b, wp + 16);
        c2_chunk_s<2>(acc, win, row + 32, row_step_b, wp + 32);
      }
      if (nch >= 1) c2_chunk_s<0>(acc, win, row, row_step_b, wp);
      if (nch >= 2) c2_chunk_s<1>(acc, win, row + 16, row_step_b, wp + 16);
    }
  } else {
    const float* base = smem + boff;
    const int2* const seg2 = reinterpret_cast<const int2*>(segs);
    int2 nxt = seg2[0];
    for (int sgi = 0; sgi < nseg; ++sgi) {
      const int2 sg = nxt;
      if (sgi + 1 < nseg) nxt = seg2[sgi + 1];  // in flight while this segment computes
      const int sg_dy = (int)(short)(sg.x & 0xffff), sg_dx0 = sg.x >> 16;
      const float* row = base + sg_dy * pitch + sg_dx0;
      const float4* wp = w4 + ((unsigned)sg.y >> 16);
      float win[kC2Rows][12];
#pragma unroll
      for (int h = 0; h < kC2Rows; ++h) {
        const float4 a0 = *reinterpret_cast<const float4*>(row + h * row_step);
        const float4 a1 = *reinterpret_cast<const float4*>(row + h * row_step + 4);
        win[h][0] = a0.x; win[h][1] = a0.y; win[h][2] = a0.z; win[h][3] = a0.w;
        win[h][4] = a1.x; win[h][5] = a1.y; win[h][6] = a1.z; win[h][7] = a1.w;
      }
      row += 8;
      int nch = (int)(short)(sg.y & 0xffff);
      for (; nch >= 3; nch -= 3, row += 12, wp += 3) {  // the window rotation has period 3
        c2_chunk<0>(acc, win, row, row_step, wp);
        c2_chunk<1>(acc, win, row + 4, row_step, wp + 1);
        c2_chunk<2>(acc, win, row + 8, row_step, wp + 2);
      }
      if (nch >= 1) c2_chunk<0>(acc, win, row, row_step, wp);
      if (nch >= 2) c2_chunk<1>(acc, win, row + 4, row_step, wp + 1);
    }
  }

  float e2 = 0.f;
#pragma unroll
  for (int h = 0; h < kC2Rows; ++h) {
    const int gr = r0 + rp + kStepRows * h;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gc = c0 + 8 * cg + j;
      if (gr < H && gc < W) {
        const int64_t g = plane + (int64_t)gr * W + gc;
        if (MODE == C2_RESIDUAL) {
          const int64_t yo = ((pl / C) / obs_repeat * C + pl % C) * (int64_t)H * W + (int64_t)gr * W + gc;
          const float r = __fsub_rn(__ldg(y + yo), acc[h][j]);
          e2 = fmaf(r, r, e2);
          out[g] = r;
        } else if (MODE == C2_COT) {
          out[g] = __fmul_rn(wgt, acc[h][j]);
        } else {
          out[g] = acc[h][j];
        }
      }
    }
  }
  if (MODE == C2_RESIDUAL) {
    const float tot = block_sum(e2, red);
    if (threadIdx.x == 0) {
      const int tiles = gridDim.x * gridDim.y;
      const int64_t l = pl / C;
      const int ch = (int)(pl % C);
      err_part[l * (int64_t)(C * tiles) + (int64_t)ch * tiles + blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
  }
}

// ---- row segments on packed FFMA2 (conv2d_rowseg2).  Same tiling and the same sliding 12-column window as
// conv2d_rowseg, but (i) the tile is stored as ROW PAIRS P[r][c] = (T[r][c], T[r + 32][c]) -- a thread's two output rows
// rp, rp + 32 read the two halves of one float2, so one FFMA2 advances both rows and one LDS.128 brings 2 columns x 2
// rows; (ii) the segments and the duplicated taps (w, w) are a kernel PARAMETER: they are read through the constant
// bank with warp-uniform indices into uniform registers, which is what keeps FFMA2 at full rate (FFMA2 R, R, UR, R;
// three vector-register pairs cost a third register-file cycle).  Per chunk of 4 taps: 2 LDS.128 + 32 FFMA2 for 64 FMAs
// (conv2d_rowseg: 3 LDS.128 + 64 FFMA).
template <int R>
__device__ __forceinline__ void c2v2_chunk(float2 (&acc)[8], float2 (&win)[12], uint32_t row, const float2* __restrict__ ww) {
  constexpr int S = (8 + 4 * R) % 12;  // slot of the incoming columns
  const float4 a = lds128(row), b = lds128(row + 16);
  win[S] = make_float2(a.x, a.y); win[S + 1] = make_float2(a.z, a.w);
  win[S + 2] = make_float2(b.x, b.y); win[S + 3] = make_float2(b.z, b.w);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 w = ww[i];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = __ffma2_rn(w, win[(q + i + 4 * R) % 12], acc[q]);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kC2Threads)
conv2d_rowseg2(const float* __restrict__ in, const float* __restrict__ eps, const float* __restrict__ y,
               float* __restrict__ out, float* __restrict__ err_part, const __grid_constant__ C2Params prm, int dy_lo,
               int dy_hi, int dx_lo, int dx_hi, int pitch2, int C, int H, int W, int64_t obs_repeat, float sa, float s1,
               float wgt, const float* __restrict__ dsc) {
  static_assert(kC2TH == 64 && kC2TW == 64 && kC2Threads == 256, "row pairs (r, r + 32), 8 column groups x 32 pairs");
  step_scalars_k1(dsc, sa, s1, wgt);
  const TweedieC tc = make_tc(s1, sa);
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[32];
  const int pr = 32 + dy_hi - dy_lo, tw4 = (kC2TW + dx_hi - dx_lo) >> 2;  // pair rows, float4 groups per tile row
  const int r0 = blockIdx.y * kC2TH, c0 = blockIdx.x * kC2TW;
  const int64_t pl = blockIdx.z;
  const int64_t plane = pl * H * W;
  {
    // fill: a warp owns pair rows warp, warp + 8, ...; W, c0 + dx_lo are multiples of 4, so a float4 is inside the image
    // or outside; both rows of a pair are loaded by the same lane and stored as two 16-byte pieces (x0 y0 x1 y1)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 2   // (unroll 4 with three resident CTAs pinned: 132 / 420 us against 124 / 389 us at L = 16 / 64)
    for (int r = warp; r < pr; r += kC2Threads / 32) {
      const int gr0 = r0 + dy_lo + r, gr1 = gr0 + 32;
      const bool ok0 = gr0 >= 0 && gr0 < H, ok1 = gr1 >= 0 && gr1 < H;
      for (int c4 = lane; c4 < tw4; c4 += 32) {
        const int gc = c0 + dx_lo + 4 * c4;
        const bool cok = gc >= 0 && gc < W;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (ok0 && cok) {
          const int64_t g = plane + (int64_t)gr0 * W + gc;
          a = *reinterpret_cast<const float4*>(in + g);
          if (MODE == C2_RESIDUAL) {
            const float4 e = *reinterpret_cast<const float4*>(eps + g);
            a.x = tweedie(a.x, e.x, tc); a.y = tweedie(a.y, e.y, tc);
            a.z = tweedie(a.z, e.z, tc); a.w = tweedie(a.w, e.w, tc);
          }
        }
        if (ok1 && cok) {
          const int64_t g = plane + (int64_t)gr1 * W + gc;
          b = *reinterpret_cast<const float4*>(in + g);
          if (MODE == C2_RESIDUAL) {
            const float4 e = *reinterpret_cast<const float4*>(eps + g);
            b.x = tweedie(b.x, e.x, tc); b.y = tweedie(b.y, e.y, tc);
            b.z = tweedie(b.z, e.z, tc); b.w = tweedie(b.w, e.w, tc);
          }
        }
        float4* d = reinterpret_cast<float4*>(smem + 2 * (r * pitch2 + 4 * c4));
        d[0] = make_float4(a.x, b.x, a.y, b.y);
        d[1] = make_float4(a.z, b.z, a.w, b.w);
      }
    }
  }
  __syncthreads();

  // thread -> (column group cg: 8 columns, row pair rp: rows rp, rp + 32).  A quarter-warp = 8 consecutive pair rows
  // of one column group: conflict-free LDS.128 because pitch2 / 2 is odd.
  const int q = threadIdx.x;
  const int rp = (q & 7) | ((q >> 6) << 3), cg = (q >> 3) & 7;
  float2 acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = make_float2(0.f, 0.f);
  const uint32_t base_s = (uint32_t)__cvta_generic_to_shared(smem) +
                          8u * (uint32_t)((rp - dy_lo) * pitch2 + (8 * cg - dx_lo));
  const int nseg = prm.nseg;
  for (int sgi = 0; sgi < nseg; ++sgi) {
    const int2 sg = prm.seg[sgi];
    const int sg_dy = (int)(short)(sg.x & 0xffff), sg_dx0 = sg.x >> 16;
    uint32_t row = base_s + 8u * (uint32_t)(sg_dy * pitch2 + sg_dx0);
    const float2* wp = prm.ww + ((unsigned)sg.y >> 16);
    float2 win[12];
    {
      const float4 a0 = lds128(row), a1 = lds128(row + 16), a2 = lds128(row + 32), a3 = lds128(row + 48);
      win[0] = make_float2(a0.x, a0.y); win[1] = make_float2(a0.z, a0.w);
      win[2] = make_float2(a1.x, a1.y); win[3] = make_float2(a1.z, a1.w);
      win[4] = make_float2(a2.x, a2.y); win[5] = make_float2(a2.z, a2.w);
      win[6] = make_float2(a3.x, a3.y); win[7] = make_float2(a3.z, a3.w);
    }
    row += 64;
    int nch = sg.y & 0xffff;
    for (; nch >= 3; nch -= 3, row += 96, wp += 12) {  // the window rotation has period 3
      c2v2_chunk<0>(acc, win, row, wp);
      c2v2_chunk<1>(acc, win, row + 32, wp + 4);
      c2v2_chunk<2>(acc, win, row + 64, wp + 8);
    }
    if (nch >= 1) c2v2_chunk<0>(acc, win, row, wp);
    if (nch >= 2) c2v2_chunk<1>(acc, win, row + 32, wp + 4);
  }

  float e2 = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int gr = r0 + rp + 32 * h;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gc = c0 + 8 * cg + j;
      const float av = h ? acc[j].y : acc[j].x;
      if (gr < H && gc < W) {
        const int64_t g = plane + (int64_t)gr * W + gc;
        if (MODE == C2_RESIDUAL) {
          const int64_t yo = ((pl / C) / obs_repeat * C + pl % C) * (int64_t)H * W + (int64_t)gr * W + gc;
          const float r = __fsub_rn(__ldg(y + yo), av);
          e2 = fmaf(r, r, e2);
          out[g] = r;
        } else if (MODE == C2_COT) {
          out[g] = __fmul_rn(wgt, av);
        } else {
          out[g] = av;
        }
      }
    }
  }
  if (MODE == C2_RESIDUAL) {
    const float tot = block_sum(e2, red);
    if (threadIdx.x == 0) {
      const int tiles = gridDim.x * gridDim.y;
      const int64_t l = pl / C;
      const int ch = (int)(pl % C);
      err_part[l * (int64_t)(C * tiles) + (int64_t)ch * tiles + blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
  }
}

// One chunk of a COLUMN segment: 4 taps x (8 output rows x 2 columns).  The window holds 12 consecutive rows of the
// thread's column pair and rotates by 4 rows per chunk (R = chunk index modulo 3, as in c2_chunk).  ODD: the pair
// starts at an odd tile column, i.e. it is not 8-byte aligned and is read as two LDS.32.
template <bool ODD>
__device__ __forceinline__ float2 c2_pair(const float* __restrict__ p) {
  return ODD ? make_float2(p[0], p[1]) : *reinterpret_cast<const float2*>(p);
}
template <int R, bool ODD>
__device__ __forceinline__ void c2_chunk_col(float2 (&acc)[8], float2 (&win)[12], const float* __restrict__ p, int pitch,
                                             const float4* __restrict__ w4) {
  constexpr int S = (8 + 4 * R) % 12;
#pragma unroll
  for (int i = 0; i < 4; ++i) win[S + i] = c2_pair<ODD>(p + i * pitch);
  const float4 t = *w4;
  const float tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      acc[q].x = fmaf(tw[i], win[(q + i + 4 * R) % 12].x, acc[q].x);
      acc[q].y = fmaf(tw[i], win[(q + i + 4 * R) % 12].y, acc[q].y);
    }
}
template <bool ODD>
__device__ __forceinline__ void c2_colseg(float2 (&acc)[8], const float* __restrict__ p, int pitch,
                                          const float4* __restrict__ wp, int nch) {
  float2 win[12];
#pragma unroll
  for (int i = 0; i < 8; ++i) win[i] = c2_pair<ODD>(p + i * pitch);
  p += 8 * pitch;
  for (; nch >= 3; nch -= 3, p += 12 * pitch, wp += 3) {
    c2_chunk_col<0, ODD>(acc, win, p, pitch, wp);
    c2_chunk_col<1, ODD>(acc, win, p + 4 * pitch, pitch, wp + 1);
    c2_chunk_col<2, ODD>(acc, win, p + 8 * pitch, pitch, wp + 2);
  }
  if (nch >= 1) c2_chunk_col<0, ODD>(acc, win, p, pitch, wp);
  if (nch >= 2) c2_chunk_col<1, ODD>(acc, win, p + 4 * pitch, pitch, wp + 1);
}

// Column-segment form (steep motion lines, camera-shake trajectories): every PSF column is a vertical 1-D correlation
// that starts at its exact first tap -- the window slides down one row per load, so no alignment padding is needed.
// Thread: column pair cp = lane (a warp reads 256 contiguous bytes per row), 8 output rows of row group warp.
// dx_lo is a multiple of 4 here (launcher), so the tile is filled with aligned 128-bit loads.
template <int MODE>
__global__ void __launch_bounds__(kC2Threads)
conv2d_colseg(const float* __restrict__ in, const float* __restrict__ eps, const float* __restrict__ y,
              float* __restrict__ out, float* __restrict__ err_part, const RowSeg* __restrict__ segs,
              const float4* __restrict__ w4, int nseg, int nw4, int dy_lo, int dy_hi, int dx_lo, int tw, int pitch,
              int C, int H, int W, int64_t obs_repeat, float sa, float s1, float wgt, const float* __restrict__ dsc) {
  step_scalars_k1(dsc, sa, s1, wgt);
  const TweedieC tc = make_tc(s1, sa);
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[32];
  const int th = kC2TH + dy_hi - dy_lo;  // tile[r][c] = image(r0 + dy_lo + r, c0 + dx_lo + c), c < tw (tw % 4 == 0)
  const int r0 = blockIdx.y * kC2TH, c0 = blockIdx.x * kC2TW;
  const int64_t pl = blockIdx.z;
  const int64_t plane = pl * H * W;
  if (nw4 > 0) {
    float4* sw4 = reinterpret_cast<float4*>(smem + th * pitch);
    RowSeg* ssg = reinterpret_cast<RowSeg*>(sw4 + nw4);
    for (int i = threadIdx.x; i < nw4; i += kC2Threads) sw4[i] = w4[i];
    for (int i = threadIdx.x; i < nseg; i += kC2Threads) ssg[i] = segs[i];
    w4 = sw4;
    segs = ssg;
  }
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool vec = (W & 3) == 0;
#pragma unroll 4
    for (int r = warp; r < th; r += kC2Threads / 32) {
      const int gr = r0 + dy_lo + r;
      const bool row_ok = gr >= 0 && gr < H;
      for (int c4 = lane; c4 < (tw >> 2); c4 += 32) {
        const int gc = c0 + dx_lo + 4 * c4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (row_ok && vec && gc >= 0 && gc < W) {
          const int64_t g = plane + (int64_t)gr * W + gc;
          const float4 a = *reinterpret_cast<const float4*>(in + g);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
          if (MODE == C2_RESIDUAL) {
            const float4 e = *reinterpret_cast<const float4*>(eps + g);
            v[0] = tweedie(v[0], e.x, tc); v[1] = tweedie(v[1], e.y, tc);
            v[2] = tweedie(v[2], e.z, tc); v[3] = tweedie(v[3], e.w, tc);
          }
        } else if (row_ok && !vec) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (gc + j >= 0 && gc + j < W) {
              v[j] = in[plane + (int64_t)gr * W + gc + j];
              if (MODE == C2_RESIDUAL) v[j] = tweedie(v[j], eps[plane + (int64_t)gr * W + gc + j], tc);
            }
        }
        *reinterpret_cast<float4*>(smem + r * pitch + 4 * c4) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
  }
  __syncthreads();

  const int cp = threadIdx.x % (kC2TW / 2), rg = threadIdx.x / (kC2TW / 2);  // columns 2cp, 2cp + 1; rows 8rg .. 8rg + 7
  float2 acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = make_float2(0.f, 0.f);
  RowSeg nxt = segs[0];
  for (int sgi = 0; sgi < nseg; ++sgi) {
    const RowSeg sg = nxt;
    if (sgi + 1 < nseg) nxt = segs[sgi + 1];
    const int dxr = sg.dx0 - dx_lo;
    const float* p = smem + (8 * rg + sg.dy - dy_lo) * pitch + 2 * cp + dxr;
    if (dxr & 1) c2_colseg<true>(acc, p, pitch, w4 + sg.w4_off, sg.nch);
    else c2_colseg<false>(acc, p, pitch, w4 + sg.w4_off, sg.nch);
  }

  float e2 = 0.f;
  const bool pair_ok = (W & 1) == 0;  // both columns of a pair are inside together and 8-byte aligned
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int gr = r0 + 8 * rg + j, gc = c0 + 2 * cp;
    if (gr >= H || gc >= W) continue;
    const int64_t g = plane + (int64_t)gr * W + gc;
    const int64_t yo = ((pl / C) / obs_repeat * C + pl % C) * (int64_t)H * W + (int64_t)gr * W + gc;
    float2 o = acc[j];
    if (MODE == C2_RESIDUAL) {
      float2 yv;
      if (pair_ok) yv = __ldg(reinterpret_cast<const float2*>(y + yo));
      else { yv.x = __ldg(y + yo); yv.y = gc + 1 < W ? __ldg(y + yo + 1) : 0.f; }
      o.x = __fsub_rn(yv.x, o.x);
      o.y = gc + 1 < W ? __fsub_rn(yv.y, o.y) : 0.f;
      e2 = fmaf(o.x, o.x, e2);
      e2 = fmaf(o.y, o.y, e2);
    } else if (MODE == C2_COT) {
      o.x = __fmul_rn(wgt, o.x);
      o.y = __fmul_rn(wgt, o.y);
    }
    if (pair_ok) *reinterpret_cast<float2*>(out + g) = o;
    else { out[g] = o.x; if (gc + 1 < W) out[g + 1] = o.y; }
  }
  if (MODE == C2_RESIDUAL) {
    const float tot = block_sum(e2, red);
    if (threadIdx.x == 0) {
      const int tiles = gridDim.x * gridDim.y;
      const int64_t l = pl / C;
      const int ch = (int)(pl % C);
      err_part[l * (int64_t)(C * tiles) + (int64_t)ch * tiles + blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
  }
}

int conv2d_err_parts(const psx_op* op) {
  return op->C * ceil_div(op->W, kC2TW) * ceil_div(op->H, kC2TH);
}

template <int MODE, bool ADJ>
static int run_conv2d(const psx_op* op, const float* in, const float* eps, const float* y, float* out,
                      float* err_part, int64_t planes, int64_t obs_repeat, float sa, float s1, float w, const float* dsc,
                      cudaStream_t st) {
  const Psf2D& psf = ADJ ? op->psf_a : op->psf_f;
  dim3 grid(ceil_div(op->W, kC2TW), ceil_div(op->H, kC2TH), (unsigned)planes);
  const float wgt = MODE == C2_COT ? (float)((double)w / (double)sa) : w;
  const size_t tab = (size_t)psf.nw4 * sizeof(float4) + (size_t)psf.nseg * sizeof(RowSeg);
  if (psf.cols) {
    const int dx_lo4 = psf.dx_lo >= 0 ? (psf.dx_lo / 4) * 4 : -(((-psf.dx_lo) + 3) / 4) * 4;  // aligned tile origin
    const int tw = (kC2TW + psf.dx_hi - dx_lo4 + 1 + 3) & ~3;  // columns c0 + dx_lo4 .. c0 + 63 + dx_hi, rounded to 4
    const int pitch = tw + 4;   // rows 8 floats apart modulo 32 banks is irrelevant here: a warp reads one row at a time
    size_t smem = (size_t)(kC2TH + psf.dy_hi - psf.dy_lo) * pitch * sizeof(float);
    const int nw4 = smem + tab <= 160 * 1024 ? psf.nw4 : 0;
    if (nw4) smem += tab;
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(conv2d_colseg<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_done = true;
    }
    conv2d_colseg<MODE><<<grid, kC2Threads, smem, st>>>(in, eps, y, out, err_part, psf.d_segs, psf.d_w4, psf.nseg, nw4,
                                                         psf.dy_lo, psf.dy_hi, dx_lo4, tw, pitch, op->C, op->H, op->W,
                                                         obs_repeat, sa, s1, wgt, dsc);
    return check_cuda(cudaGetLastError(), "conv2d_colseg launch");
  }
  if (psf.h_v2 && (op->W & 3) == 0 && !env_opts().no_c2v2) {
    // packed-FFMA2 kernel: row-pair tile, segments / taps in the parameter bank
    int pitch2 = kC2TW + psf.dx_hi - psf.dx_lo;  // multiple of 4 -> + 2: even with pitch2 / 2 odd
    pitch2 += 2;
    const size_t smem2 = (size_t)(32 + psf.dy_hi - psf.dy_lo) * pitch2 * sizeof(float2);
    if (smem2 <= 200 * 1024) {
      static bool attr2 = false;
      if (!attr2) {
        cudaFuncSetAttribute(conv2d_rowseg2<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr2 = true;
      }
      conv2d_rowseg2<MODE><<<grid, kC2Threads, smem2, st>>>(in, eps, y, out, err_part, *psf.h_v2, psf.dy_lo, psf.dy_hi,
                                                           psf.dx_lo, psf.dx_hi, pitch2, op->C, op->H, op->W,
                                                           obs_repeat, sa, s1, wgt, dsc);
      return check_cuda(cudaGetLastError(), "conv2d_rowseg2 launch");
    }
  }
  const int tw = kC2TW + psf.dx_hi - psf.dx_lo;
  int pitch = tw;                      // multiple of 4 with pitch / 4 odd: conflict-free LDS.128 (see the kernel)
  if (((pitch >> 2) & 1) == 0) pitch += 4;
  size_t smem = (size_t)(kC2TH + psf.dy_hi - psf.dy_lo) * pitch * sizeof(float);
  const int nw4 = smem + tab <= 96 * 1024 ? psf.nw4 : 0;  // huge PSFs read their taps through L1 instead
  if (nw4) smem += tab;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(conv2d_rowseg<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv2d_rowseg<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  if (nw4)
    conv2d_rowseg<MODE, true><<<grid, kC2Threads, smem, st>>>(in, eps, y, out, err_part, psf.d_segs, psf.d_w4, psf.nseg,
                                                               nw4, psf.dy_lo, psf.dy_hi, psf.dx_lo, psf.dx_hi, pitch,
                                                               op->C, op->H, op->W, obs_repeat, sa, s1, wgt, dsc);
  else
    conv2d_rowseg<MODE, false><<<grid, kC2Threads, smem, st>>>(in, eps, y, out, err_part, psf.d_segs, psf.d_w4, psf.nseg,
                                                                nw4, psf.dy_lo, psf.dy_hi, psf.dx_lo, psf.dx_hi, pitch,
                                                                op->C, op->H, op->W, obs_repeat, sa, s1, wgt, dsc);
  return check_cuda(cudaGetLastError(), "conv2d_rowseg launch");
}

int launch_pre_conv2d(const psx_op* op, const float* x, const float* eps, const float* y, int64_t L,
                      int64_t obs_repeat, float sa, float s1, float w, const float* dsc, float* cot, float* err_part,
                      float* x0_out, float* ws, cudaStream_t st) {
  if (x0_out) return fail(PSX_ERR_UNSUPPORTED, "psx_dps_pre: d_x0_out is not produced for blur operators");
  const int64_t planes = L * op->C;
  int rc = run_conv2d<C2_RESIDUAL, false>(op, x, eps, y, ws, err_part, planes, obs_repeat, sa, s1, w, dsc, st);
  if (rc) return rc;
  return run_conv2d<C2_COT, true>(op, ws, nullptr, nullptr, cot, nullptr, planes, 1, sa, s1, w, dsc, st);
}

// ------------------------------------------------------------------------------------------ stand-alone A / A^T
int launch_op_pointwise(const psx_op* op, bool adjoint, const float* in, float* out, int64_t L, cudaStream_t st);

int launch_op(const psx_op* op, bool adjoint, const float* in, float* out, int64_t L, float* ws,
              cudaStream_t st) {
  const int64_t planes = L * op->C;
  switch (op->kind) {
    case PSX_OP_SEPBLUR: {
      if (!adjoint) {  // y = V(H(x))
        int rc = run_rows<ROWS_PLAIN>(op, op->fh, in, nullptr, ws, planes, 1.f, 0.f, 1.f, nullptr, st);
        if (rc) return rc;
        return run_cols<false>(op, op->fv, op->fv, ws, nullptr, out, nullptr, planes, 1, st);
      }
      int rc = run_cols<false>(op, op->av, op->av, in, nullptr, ws, nullptr, planes, 1, st);
      if (rc) return rc;
      return run_rows<ROWS_PLAIN>(op, op->ah, ws, nullptr, out, planes, 1.f, 0.f, 1.f, nullptr, st);
    }
    case PSX_OP_CONV2D:
      if (!adjoint)
        return run_conv2d<C2_PLAIN, false>(op, in, nullptr, nullptr, out, nullptr, planes, 1, 1.f, 0.f, 1.f, nullptr, st);
      return run_conv2d<C2_PLAIN, true>(op, in, nullptr, nullptr, out, nullptr, planes, 1, 1.f, 0.f, 1.f, nullptr, st);
    default:
      return launch_op_pointwise(op, adjoint, in, out, L, st);
  }
}

}  // namespace psx
