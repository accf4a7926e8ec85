// psx_half.cu -- bf16-state variants of K1 (identity / mask / 4x box) and K2: the sampler state, eps, the
// cotangent and the network VJP are stored as bf16 (2 B/element), every value is widened to fp32 on load, the
// arithmetic is the fp32 kernels' arithmetic instruction for instruction, and results are rounded to bf16 (RN) on
// store.  Hence   K_bf16(inputs) == bf16_rn( K_fp32( float(inputs) ) )   bit for bit -- the property the tests check.
// Traffic: K1 2+2+4 (y stays fp32, shared) in, 2 out; K2 2*4 in, 2 out with in-kernel Philox noise: 18 B/element for
// the step against 40 B/element in fp32.  Vector width: 8 elements (16 B); n % 8 != 0 takes the scalar kernels.
#include <cuda_bf16.h>

#include <cmath>

#include "psx_common.cuh"

namespace psx {

int one_wave_parts(int slots, int64_t L, int64_t units_per_sample, int cap);
template <typename Kern>
static int resident_slots_h(Kern kernel) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
  return per_sm * sm_count();
}

struct Bf8 {  // 8 bf16 values = one 128-bit access
  uint4 raw;
};
__device__ __forceinline__ Bf8 ld_bf8(const __nv_bfloat16* p) {
  Bf8 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.raw.x), "=r"(v.raw.y), "=r"(v.raw.z), "=r"(v.raw.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_bf8(__nv_bfloat16* p, const Bf8& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.raw.x), "r"(v.raw.y),
               "r"(v.raw.z), "r"(v.raw.w)
               : "memory");
}
__device__ __forceinline__ void unpack8(const Bf8& v, float (&f)[8]) {
  const uint32_t w[4] = {v.raw.x, v.raw.y, v.raw.z, v.raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);  // .x = lo (low half), round to nearest even
  return *reinterpret_cast<const uint32_t*>(&p);
}
__device__ __forceinline__ Bf8 pack8(const float (&f)[8]) {
  Bf8 v;
  v.raw.x = pack2(f[0], f[1]); v.raw.y = pack2(f[2], f[3]);
  v.raw.z = pack2(f[4], f[5]); v.raw.w = pack2(f[6], f[7]);
  return v;
}

// ------------------------------------------------------------------------------------------------ K1
// grid = (parts, L); CTA (p, l) owns the 8-element groups [p*chunk, (p+1)*chunk) of sample l.
template <bool MASK>
__global__ void __launch_bounds__(kThreads)
k1_pointwise_h8(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ eps, const float* __restrict__ y,
                const uint8_t* __restrict__ keep, __nv_bfloat16* __restrict__ cot, float* __restrict__ err_part, int slots,
                int64_t n, int64_t chunk8, int64_t obs_repeat, float sa, float s1, float coef,
                const float* __restrict__ dsc) {
  step_scalars_k1(dsc, sa, s1, coef);
  const TweedieC tc = make_tc(s1, sa);
  __shared__ float red[32];
  const int64_t l = blockIdx.y, n8 = n >> 3;
  const int64_t beg = (int64_t)blockIdx.x * chunk8, end = min(beg + chunk8, n8);
  const __nv_bfloat16* xs = x + l * n;
  const __nv_bfloat16* es = eps + l * n;
  const float* ys = y + (l / obs_repeat) * n;
  __nv_bfloat16* cs = cot + l * n;
  float acc = 0.f;
  constexpr int U = 2;
  for (int64_t base = beg + threadIdx.x; base < end; base += (int64_t)kThreads * U) {
    Bf8 xv[U], ev[U];
    float4 ya[U], yb[U];
    uint2 kv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + (int64_t)u * kThreads;
      if (i < end) {
        xv[u] = ld_bf8(xs + 8 * i);
        ev[u] = ld_bf8(es + 8 * i);
        ya[u] = __ldg(reinterpret_cast<const float4*>(ys) + 2 * i);
        yb[u] = __ldg(reinterpret_cast<const float4*>(ys) + 2 * i + 1);
        if (MASK) kv[u] = __ldg(reinterpret_cast<const uint2*>(keep) + i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + (int64_t)u * kThreads;
      if (i < end) {
        float xf[8], ef[8], d[8];
        unpack8(xv[u], xf);
        unpack8(ev[u], ef);
        const float yf[8] = {ya[u].x, ya[u].y, ya[u].z, ya[u].w, yb[u].x, yb[u].y, yb[u].z, yb[u].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float r = __fsub_rn(yf[j], tweedie(xf[j], ef[j], tc));
          if (MASK) r = ((j < 4 ? kv[u].x >> (8 * j) : kv[u].y >> (8 * (j - 4))) & 0xffu) ? r : 0.f;
          acc = fmaf(r, r, acc);
          d[j] = __fmul_rn(coef, r);
        }
        st_bf8(cs + 8 * i, pack8(d));
      }
    }
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) err_part[l * slots + blockIdx.x] = tot;
  if (blockIdx.x == 0)  // unused slots of this sample's row must read as zero in K2
    for (int i = gridDim.x + threadIdx.x; i < slots; i += kThreads) err_part[l * slots + i] = 0.f;
}

template <bool MASK>
__global__ void __launch_bounds__(kThreads)
k1_pointwise_hs(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ eps, const float* __restrict__ y,
                const uint8_t* __restrict__ keep, __nv_bfloat16* __restrict__ cot, float* __restrict__ err_part, int slots,
                int64_t n, int64_t chunk, int64_t obs_repeat, float sa, float s1, float coef,
                const float* __restrict__ dsc) {
  step_scalars_k1(dsc, sa, s1, coef);
  const TweedieC tc = make_tc(s1, sa);
  __shared__ float red[32];
  const int64_t l = blockIdx.y;
  const int64_t beg = (int64_t)blockIdx.x * chunk, end = min(beg + chunk, n);
  float acc = 0.f;
  for (int64_t i = beg + threadIdx.x; i < end; i += kThreads) {
    const int64_t j = l * n + i;
    float r = __fsub_rn(y[(l / obs_repeat) * n + i], tweedie(__bfloat162float(x[j]), __bfloat162float(eps[j]), tc));
    if (MASK) r = keep[i] ? r : 0.f;
    acc = fmaf(r, r, acc);
    cot[j] = __float2bfloat16_rn(__fmul_rn(coef, r));
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) err_part[l * slots + blockIdx.x] = tot;
  if (blockIdx.x == 0)
    for (int i = gridDim.x + threadIdx.x; i < slots; i += kThreads) err_part[l * slots + i] = 0.f;
}

int launch_pre_pointwise_bf16(const psx_op* op, const void* x, const void* eps, const float* y, int64_t L,
                              int64_t obs_repeat, float sa, float s1, float w, const float* dsc, void* cot,
                              float* err_part, cudaStream_t st) {
  const int64_t n = op->n;
  const int slots = op->err_parts;
  const bool mask = op->kind == PSX_OP_MASK;
  const float coef = (float)((double)w / (double)sa);
  const __nv_bfloat16 *xb = (const __nv_bfloat16*)x, *eb = (const __nv_bfloat16*)eps;
  __nv_bfloat16* cb = (__nv_bfloat16*)cot;
#define PSX_K1H(KERNEL, UNITS)                                                                              \
  {                                                                                                         \
    static int rs = 0;                                                                                      \
    if (!rs) rs = resident_slots_h(KERNEL);                                                                 \
    const int parts = one_wave_parts(rs, L, (UNITS), slots);                                                \
    const int64_t chunk = ((UNITS) + parts - 1) / parts;                                                    \
    KERNEL<<<dim3(parts, (unsigned)L), kThreads, 0, st>>>(xb, eb, y, op->d_keep, cb, err_part, slots, n,    \
                                                          chunk, obs_repeat, sa, s1, coef, dsc);            \
  }
  if (n % 8 == 0) {
    if (mask) PSX_K1H(k1_pointwise_h8<true>, n / 8) else PSX_K1H(k1_pointwise_h8<false>, n / 8)
  } else {
    if (mask) PSX_K1H(k1_pointwise_hs<true>, n) else PSX_K1H(k1_pointwise_hs<false>, n)
  }
#undef PSX_K1H
  return check_cuda(cudaGetLastError(), "k1_pointwise bf16 launch");
}

// ---- K1 for the 4 x 4 box (super-resolution): one thread per coarse pixel, 4 rows x 4 bf16 (8 B) of x_t / eps in,
// one fp32 y in, 4 rows x 4 bf16 cotangents out; the arithmetic (incl. the summation order) is k1_box<4>'s.
__global__ void __launch_bounds__(kThreads)
k1_box4_h(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ eps, const float* __restrict__ y,
          __nv_bfloat16* __restrict__ cot, float* __restrict__ err_part, int slots, int planes, int H, int W,
          int64_t chunk, int64_t obs_repeat, float sa, float s1, float coef, const float* __restrict__ dsc) {
  step_scalars_k1(dsc, sa, s1, coef);
  const TweedieC tc = make_tc(s1, sa);
  __shared__ float red[32];
  constexpr int F = 4;
  const int Hc = H / F, Wc = W / F;
  const int64_t n = (int64_t)planes * H * W, ny = (int64_t)planes * Hc * Wc;
  const int64_t l = blockIdx.y;
  const int64_t beg = (int64_t)blockIdx.x * chunk, end = min(beg + chunk, ny);
  const float inv = 1.0f / (float)(F * F);
  float acc = 0.f;
  for (int64_t q = beg + threadIdx.x; q < end; q += kThreads) {
    const int cx = (int)(q % Wc);
    const int64_t t = q / Wc;
    const int cy = (int)(t % Hc);
    const int pl = (int)(t / Hc);
    const int64_t off = l * n + ((int64_t)pl * H + (int64_t)cy * F) * W + (int64_t)cx * F;
    uint2 xr[F], er[F];
#pragma unroll
    for (int dy = 0; dy < F; ++dy) {
      xr[dy] = __ldg(reinterpret_cast<const uint2*>(x + off + (int64_t)dy * W));
      er[dy] = __ldg(reinterpret_cast<const uint2*>(eps + off + (int64_t)dy * W));
    }
    float s = 0.f;
#pragma unroll
    for (int dy = 0; dy < F; ++dy) {
      const uint32_t xw[2] = {xr[dy].x, xr[dy].y}, ew[2] = {er[dy].x, er[dy].y};
#pragma unroll
      for (int dx = 0; dx < F; ++dx) {
        const float xv = __uint_as_float(dx & 1 ? xw[dx >> 1] & 0xffff0000u : xw[dx >> 1] << 16);
        const float ev = __uint_as_float(dx & 1 ? ew[dx >> 1] & 0xffff0000u : ew[dx >> 1] << 16);
        s = __fadd_rn(s, tweedie(xv, ev, tc));
      }
    }
    const float avg = __fmul_rn(s, inv);
    const float r = __fsub_rn(__ldg(y + (l / obs_repeat) * ny + q), avg);
    acc = fmaf(r, r, acc);
    const float d = __fmul_rn(coef, __fmul_rn(r, inv));
    const uint32_t dd = pack2(d, d);
#pragma unroll
    for (int dy = 0; dy < F; ++dy) *reinterpret_cast<uint2*>(cot + off + (int64_t)dy * W) = make_uint2(dd, dd);
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) err_part[l * slots + blockIdx.x] = tot;
  if (blockIdx.x == 0)
    for (int i = gridDim.x + threadIdx.x; i < slots; i += kThreads) err_part[l * slots + i] = 0.f;
}

int launch_pre_box_bf16(const psx_op* op, const void* x, const void* eps, const float* y, int64_t L, int64_t obs_repeat,
                        float sa, float s1, float w, const float* dsc, void* cot, float* err_part, cudaStream_t st) {
  if (op->factor != 4 || op->W % 4 != 0) return fail(PSX_ERR_UNSUPPORTED, "psx_dps_pre_bf16: box factor 4 (W % 4 == 0) only");
  const int slots = op->err_parts;
  const float coef = (float)((double)w / (double)sa);
  static int rs = 0;
  if (!rs) rs = resident_slots_h(k1_box4_h);
  const int parts = one_wave_parts(rs, L, op->n_y, slots);
  const int64_t chunk = (op->n_y + parts - 1) / parts;
  k1_box4_h<<<dim3(parts, (unsigned)L), kThreads, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)eps, y,
                                                          (__nv_bfloat16*)cot, err_part, slots, op->C, op->H, op->W, chunk,
                                                          obs_repeat, sa, s1, coef, dsc);
  return check_cuda(cudaGetLastError(), "k1_box bf16 launch");
}

// ------------------------------------------------------------------------------------------------ K2
// ZMODE: 0 no noise, 1 bf16 noise tensor, 2 in-kernel Philox (the fp32 field of psx_philox_normal, used unrounded)
template <int ZMODE, bool VEC>
__global__ void __launch_bounds__(kThreads)
k2_post_h(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ eps, const __nv_bfloat16* __restrict__ cot,
          const __nv_bfloat16* __restrict__ vjp, const __nv_bfloat16* __restrict__ z, const float* __restrict__ err_part,
          int err_parts, int64_t n, int64_t chunk, float sa, float s1, float c_ell, float c_s, float sd, float gamma,
          __nv_bfloat16* __restrict__ x_next, float* __restrict__ err_out, const float* __restrict__ dsc, uint64_t seed,
          uint64_t step, const uint64_t* __restrict__ rng) {
  step_scalars_k2(dsc, sa, s1, c_ell, c_s, sd, gamma);
  if (ZMODE == 2 && rng != nullptr) {
    seed = rng[0];
    step = rng[1];
  }
  const TweedieC tc = make_tc(s1, sa);
  __shared__ float red[32];
  const int64_t l = blockIdx.y;
  float scale = gamma;  // err_parts == 0: fixed guidance scale (PGDM); otherwise DPS: gamma / (|r| + 1e-9)
  if (err_parts > 0) {
    const float e2 = sum_parts(err_part + l * err_parts, err_parts, red);
    const float err = sqrtf(e2);
    scale = __fdiv_rn(gamma, __fadd_rn(err, 1e-9f));
    if (err_out && blockIdx.x == 0 && threadIdx.x == 0) err_out[l] = err;
  }
  auto update = [&](float xv, float ev, float dv, float vv, float zv) {
    const float x0 = tweedie(xv, ev, tc);
    float m = __fadd_rn(__fmul_rn(c_ell, xv), __fmul_rn(c_s, x0));
    if (ZMODE != 0) m = __fadd_rn(m, __fmul_rn(sd, zv));
    const float g = __fadd_rn(dv, __fmul_rn(-s1, vv));
    return __fadd_rn(m, __fmul_rn(scale, g));
  };
  const int64_t so = l * n;
  if (VEC) {
    const int64_t n8 = n >> 3;
    const int64_t beg = (int64_t)blockIdx.x * chunk, end = min(beg + chunk, n8);
    for (int64_t i = beg + threadIdx.x; i < end; i += kThreads) {
      const Bf8 xv = ld_bf8(x + so + 8 * i), ev = ld_bf8(eps + so + 8 * i), dv = ld_bf8(cot + so + 8 * i),
                vv = ld_bf8(vjp + so + 8 * i);
      float zf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (ZMODE == 1) unpack8(ld_bf8(z + so + 8 * i), zf);
      if (ZMODE == 2) {  // element group g = flat index / 4: two Philox blocks per 8 elements
        const uint64_t g0 = (uint64_t)(so >> 2) + 2 * (uint64_t)i;  // so % 8 == 0 in the vector kernel
        const float4 a = philox_normal4(g0, seed, step), b = philox_normal4(g0 + 1, seed, step);
        zf[0] = a.x; zf[1] = a.y; zf[2] = a.z; zf[3] = a.w; zf[4] = b.x; zf[5] = b.y; zf[6] = b.z; zf[7] = b.w;
      }
      float xf[8], ef[8], df[8], vf[8], o[8];
      unpack8(xv, xf); unpack8(ev, ef); unpack8(dv, df); unpack8(vv, vf);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = update(xf[j], ef[j], df[j], vf[j], zf[j]);
      st_bf8(x_next + so + 8 * i, pack8(o));
    }
  } else {
    const int64_t beg = (int64_t)blockIdx.x * chunk, end = min(beg + chunk, n);
    for (int64_t i = beg + threadIdx.x; i < end; i += kThreads) {
      const int64_t j = so + i;
      float zv = 0.f;
      if (ZMODE == 1) zv = __bfloat162float(z[j]);
      if (ZMODE == 2) {
        const float4 zz = philox_normal4((uint64_t)(j >> 2), seed, step);
        const int c = (int)(j & 3);
        zv = c == 0 ? zz.x : c == 1 ? zz.y : c == 2 ? zz.z : zz.w;
      }
      x_next[j] = __float2bfloat16_rn(update(__bfloat162float(x[j]), __bfloat162float(eps[j]), __bfloat162float(cot[j]),
                                             __bfloat162float(vjp[j]), zv));
    }
  }
}

int launch_post_bf16(const void* x, const void* eps, const void* cot, const void* vjp, const void* z, const float* err_part,
                     int err_parts, int64_t L, int64_t n, float sa, float s1, float c_ell, float c_s, float sd,
                     float gamma, const float* dsc, void* x_next, float* err_out, int zmode, uint64_t seed,
                     uint64_t step, const uint64_t* rng, cudaStream_t st) {
  typedef const __nv_bfloat16* P;
#define PSX_K2H(KERNEL, UNITS)                                                                                  \
  {                                                                                                             \
    static int rs = 0;                                                                                          \
    if (!rs) rs = resident_slots_h(KERNEL);                                                                     \
    const int parts = one_wave_parts(rs, L, (UNITS), 1 << 20);                                                  \
    const int64_t chunk = ((UNITS) + parts - 1) / parts;                                                        \
    KERNEL<<<dim3(parts, (unsigned)L), kThreads, 0, st>>>((P)x, (P)eps, (P)cot, (P)vjp, (P)z, err_part,         \
                                                          err_parts, n, chunk, sa, s1, c_ell, c_s, sd, gamma,   \
                                                          (__nv_bfloat16*)x_next, err_out, dsc, seed, step, rng); \
  }
  if (n % 8 == 0) {
    if (zmode == 2) PSX_K2H((k2_post_h<2, true>), n / 8) else if (zmode == 1) PSX_K2H((k2_post_h<1, true>), n / 8)
    else PSX_K2H((k2_post_h<0, true>), n / 8)
  } else {
    if (zmode == 2) PSX_K2H((k2_post_h<2, false>), n) else if (zmode == 1) PSX_K2H((k2_post_h<1, false>), n)
    else PSX_K2H((k2_post_h<0, false>), n)
  }
#undef PSX_K2H
  return check_cuda(cudaGetLastError(), "k2_post bf16 launch");
}

}  // namespace psx
