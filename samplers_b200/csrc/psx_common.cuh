// psx_common.cuh -- shared device helpers for libpsx (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <mutex>
#include <stdint.h>

#include <string>

#include "psx.h"

namespace psx {

constexpr int kThreads = 256;
constexpr int kMaxParts = 64;  // upper bound on per-sample partial sums (pointwise kernels)

// ---------------------------------------------------------------- error plumbing
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int check_cuda(cudaError_t e, const char* what);
int sm_count();

#define PSX_REQUIRE(cond, msg) \
  do {                         \
    if (!(cond)) return ::psx::fail(PSX_ERR_INVALID, msg); \
  } while (0)

// ---------------------------------------------------------------- operator descriptor
struct Taps {
  int k;    // number of taps kept (after pruning), padded to a multiple of 8 with zeros
  int lo;   // offset of tap 0: out[p] = sum_i w[i] * in[p + lo + i];  lo % 4 == 0
  // every tap duplicated as (w, w): the packed FFMA2 (fma.rn.f32x2) takes the pair straight from the
  // kernel-parameter constant bank through a uniform register (LDCU.64 + FFMA2 R, R, UR, R)
  float2 ww[PSX_MAX_TAPS + 9];
};

// 2-D PSF as row segments: PSF row dy is a 1-D correlation whose taps start at column offset dx0 (a multiple of 4)
// and span nch chunks of 4 taps (zero padded); w4_off indexes the float4 tap array.
struct RowSeg {
  int16_t dy, dx0;
  int16_t nch;
  uint16_t w4_off;
};
// The same row-segment PSF as a kernel parameter (conv2d_rowseg2): segments and DUPLICATED taps (w, w) live in the
// parameter constant bank, so that the tap pairs reach packed FFMA2 through uniform registers.
constexpr int kC2MaxSeg = 80, kC2MaxTap = 368;
struct C2Params {
  int nseg, pad;
  int2 seg[kC2MaxSeg];      // {dy | dx0 << 16, nch | first tap << 16}
  float2 ww[kC2MaxTap];     // (w, w), 4 per chunk
};
static_assert(sizeof(C2Params) <= 3712, "kernel parameter space (4 KB with the other arguments)");
struct Psf2D {              // one direction (forward or adjoint = flipped PSF)
  C2Params* h_v2;           // host, owned: parameter image for conv2d_rowseg2, or null (column form / too large)
  RowSeg* d_segs;           // device, owned
  float4* d_w4;             // device, owned: zero-padded taps, 4 per chunk
  int nseg, nw4;            // segments, float4 tap groups
  // cols == 0 (row segments): dy_lo/dy_hi = min / max dy, dx_lo/dx_hi = min dx0 / max (dx0 + 4 nch), multiples of 4.
  // cols == 1 (column segments: RowSeg.dx0 = the column offset, RowSeg.dy = first tap row, taps run down the
  //            column): dy_lo/dy_hi = min dy / max (dy + 4 nch), dx_lo/dx_hi = min / max column offset.
  int cols;
  int dy_lo, dy_hi;
  int dx_lo, dx_hi;
};

}  // namespace psx

struct psx_op {
  int kind;
  int64_t n, n_y;
  int C, H, W;
  int factor;                 // box
  const uint8_t* d_keep;      // mask (caller-owned)
  psx::Taps fh, fv, ah, av;   // separable blur: forward / adjoint taps (rows, cols)
  psx::Psf2D psf_f, psf_a;    // conv2d: forward / adjoint PSF as row segments (device arrays owned)
  int kh, kw;
  int err_parts;
  int col_tc;                 // sepblur: column-strip width
  // sepblur: side streams + events for running K1 as independent sample groups that overlap each other's
  // kernel tails (launch_pre_sepblur); created with the descriptor when a device is present, owned
  int aux_n;                  // number of usable side streams (0: no split)
  cudaStream_t aux_stream[3];
  cudaEvent_t ev_fork, ev_join[3];
  std::mutex* aux_mu;         // serialises the fork/join sequence of concurrent callers
  // sepblur on the tensor cores (psx_tcblur.cu): 0 = not applicable, else the zero padding (K window = 64 + 2 pad)
  int tc_pad;
  uint8_t* d_tc_img;          // device, owned: fp16 hi/lo shared-memory image of the Toeplitz block
  float tc_inv_scale;         // 1 / (power-of-two scale applied to the taps in the image)
};

namespace psx {

// ---------------------------------------------------------------- memory access
// Streaming 128-bit loads/stores: data is touched once per kernel, keep it out of L1.
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// the same load as a plain (non-volatile) asm: the compiler may batch several of them ahead of their uses
__device__ __forceinline__ float4 ld_nc4(const float* p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// 256-bit streaming load (sm_100): 32 bytes = one sector per lane
__device__ __forceinline__ void ld_nc8(const float* p, float4& a, float4& b) {
  asm("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(p));
}
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// ---------------------------------------------------------------- bf16 storage (production state)
// bf16 -> fp32 is a 16-bit shift; fp32 -> bf16 rounds to nearest even.  Two values per 32-bit word, low half first.
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t bf16_pack2(float lo, float hi) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&p);
}

// ---------------------------------------------------------------- arithmetic with torch's roundings
// networks/base.py:42-43 evaluates mul, sub, div as three separately rounded tensor ops;
// the intrinsics stop nvcc from contracting them into an FMA / a reciprocal multiply.
// The division by the per-launch constant sa uses Markstein's correction instead of the generic IEEE
// division sequence (~25 instructions incl. range checks): with rcp = RN(1/sa),
//   q0 = RN(t*rcp);  rem = t - q0*sa (exact in one FMA);  q = RN(q0 + rem*rcp)
// is the correctly rounded t/sa for normal operands (sa is in [1e-3, 1]) -- 3 instructions.
struct TweedieC {
  float s1, sa, rcp;
};
__device__ __forceinline__ TweedieC make_tc(float s1, float sa) {
  TweedieC c;
  c.s1 = s1;
  c.sa = sa;
  c.rcp = __frcp_rn(sa);
  return c;
}
__device__ __forceinline__ float tweedie(float x, float e, const TweedieC& c) {
  const float t = __fsub_rn(x, __fmul_rn(c.s1, e));
  const float q0 = __fmul_rn(t, c.rcp);
  const float rem = __fmaf_rn(-q0, c.sa, t);
  return __fmaf_rn(rem, c.rcp, q0);
}

// Step scalars: kernels take them by value; the *_dev entry points (graph-replayable launches) pass a device
// row psx_step_row = [sqrt_acp, sqrt_1m_acp, lik_weight / sqrt_acp, c_ell, c_s, std, gamma, -] that overrides them.
__device__ __forceinline__ void step_scalars_k1(const float* dsc, float& sa, float& s1, float& coef) {
  if (dsc != nullptr) {
    sa = __ldg(dsc);
    s1 = __ldg(dsc + 1);
    coef = __ldg(dsc + 2);
  }
}
__device__ __forceinline__ void step_scalars_coef(const float* dsc, float& coef) {
  if (dsc != nullptr) coef = __ldg(dsc + 2);
}
__device__ __forceinline__ void step_scalars_k2(const float* dsc, float& sa, float& s1, float& c_ell, float& c_s,
                                                float& sd, float& gamma) {
  if (dsc != nullptr) {
    sa = __ldg(dsc);
    s1 = __ldg(dsc + 1);
    c_ell = __ldg(dsc + 3);
    c_s = __ldg(dsc + 4);
    sd = __ldg(dsc + 5);
    gamma = __ldg(dsc + 6);
  }
}

// ---------------------------------------------------------------- in-kernel noise (Philox4x32-10 + Box-Muller)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    if (r) {
      k.x += 0x9E3779B9u;
      k.y += 0xBB67AE85u;
    }
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
  }
  return c;
}
__device__ __forceinline__ float philox_u01(uint32_t x) {  // (x + 0.5) * 2^-32 in (0, 1]
  return __fadd_rn(__fmul_rn(__uint2float_rn(x), 2.3283064365386963e-10f), 1.1641532182693481e-10f);
}
__device__ __forceinline__ float2 box_muller(float ua, float ub) {
  const float r = sqrtf(-2.f * logf(ua));
  float sn, cs;
  sincospif(2.f * ub, &sn, &cs);
  return make_float2(r * cs, r * sn);
}
__device__ __forceinline__ float4 philox_normal4(uint64_t group, uint64_t seed, uint64_t step) {
  const uint4 b = philox4x32_10(make_uint4((uint32_t)group, (uint32_t)(group >> 32), (uint32_t)step, (uint32_t)(step >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float2 p = box_muller(philox_u01(b.x), philox_u01(b.y)), q = box_muller(philox_u01(b.z), philox_u01(b.w));
  return make_float4(p.x, p.y, q.x, q.y);
}


// ---------------------------------------------------------------- reductions (fixed order => deterministic)
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the CTA; result valid in thread 0.  `red` must hold >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = 0.f;
  if (wid == 0) {
    t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
  }
  return t;
}

// Per-sample |r|^2 from the partial sums K1 left behind; every thread gets the value.
__device__ __forceinline__ float sum_parts(const float* part, int parts, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (wid == 0) {
    float t = 0.f;
    for (int i = lane; i < parts; i += 32) t += part[i];
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// launchers implemented in the .cu files ------------------------------------------------
int launch_pre_pointwise(const psx_op* op, const float* x, const float* eps, const float* y, int64_t L,
                         int64_t obs_repeat, float sa, float s1, float w, const float* dsc, float* cot, float* err_part,
                         float* x0_out, cudaStream_t st);
int launch_pre_box(const psx_op* op, const float* x, const float* eps, const float* y, int64_t L,
                   int64_t obs_repeat, float sa, float s1, float w, const float* dsc, float* cot, float* err_part,
                   float* x0_out, cudaStream_t st);
// mean_out != nullptr: K1 also writes the bridge mean c_ell x_t + c_s x0 (tensor-core blur at small batches, see
// fuses_mean)
int launch_pre_sepblur(const psx_op* op, const float* x, const float* eps, const float* y, int64_t L,
                       int64_t obs_repeat, float sa, float s1, float w, const float* dsc, float* cot, float* err_part,
                       float* x0_out, float* ws, cudaStream_t st, bool half = false, float* mean_out = nullptr,
                       float c_ell = 0.f, float c_s = 0.f, const float* zn = nullptr, float sd = 0.f);
bool fuses_mean(const psx_op* op, int64_t L);
bool tcblur_mean_fits(const psx_op* op, int64_t L);
int launch_post_mean(const float* mean, const float* cot, const float* vjp, const float* z, const float* err_part,
                     int err_parts, int64_t L, int64_t n, float s1, float sd, float gamma, const float* dsc,
                     float* x_next, float* err_out, cudaStream_t st);
int launch_pre_conv2d(const psx_op* op, const float* x, const float* eps, const float* y, int64_t L,
                      int64_t obs_repeat, float sa, float s1, float w, const float* dsc, float* cot, float* err_part,
                      float* x0_out, float* ws, cudaStream_t st);
int launch_op(const psx_op* op, bool adjoint, const float* in, float* out, int64_t L, float* ws,
              cudaStream_t st);
int sepblur_plan(psx_op* op);
void tcblur_plan(psx_op* op);
void tcblur_release(psx_op* op);
bool tcblur_available(const psx_op* op);
int launch_pre_sepblur_tc(const psx_op* op, const float* x, const float* eps, const float* y, int64_t L,
                          int64_t obs_repeat, float sa, float s1, float w, const float* dsc, float* cot,
                          float* err_part, float* mean_out, float c_ell, float c_s, const float* zn, float sd,
                          cudaStream_t st);
int conv2d_err_parts(const psx_op* op);

// Kernel-selection switches of the environment, read once (psx_reload_env re-reads them).
struct EnvOpts {
  bool no_pipe, no_fast16, no_tc, fused;
  bool no_c2v2;       // PSX_NO_C2V2: 2-D row-segment PSFs on the scalar-FFMA kernel (conv2d_rowseg) instead of conv2d_rowseg2
  bool tc_persist;    // PSX_TC_PERSIST: blur_k1_tc as one wave of persistent cluster pairs when the planes exceed it
  int mean_lag_ns;    // PSX_MEAN_LAG_NS: the bridge-mean CTAs of blur_k1_tc start this long after the kernel (default 5000:
                      // behind the plane CTAs' load phase, see psx_tcblur.cu)
  int split;  // PSX_SPLIT: forced number of K1 sample groups, 0 = automatic
};
const EnvOpts& env_opts();

}  // namespace psx
