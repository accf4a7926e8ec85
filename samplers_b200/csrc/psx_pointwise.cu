// psx_pointwise.cu -- HBM-bound kernels of the DPS step (sm_100a):
//   K1 for identity / inpainting-mask / box super-resolution operators,
//   K2 (bridge update + guidance + injected noise), the final Tweedie estimate,
//   and the stand-alone forward / adjoint of those operators.
// Every kernel streams each tensor exactly once with 128-bit accesses.
#include <cmath>

#include "psx_common.cuh"

#ifndef PSX_K2M_U
#define PSX_K2M_U 4  // float4 groups per thread and batch in k2_post_mean (L <= 16: 1 / 2 / 4 -> 10.6 / 10.45 / 10.0 us)
#endif
namespace psx {

// =========================================================================== K1: identity / mask
// grid = (parts, L); CTA (p, l) owns the float4 range [p*chunk, (p+1)*chunk) of sample l.
template <bool MASK, bool WRITE_X0>
__global__ void __launch_bounds__(kThreads)
k1_pointwise_v4(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ y,
                const uint8_t* __restrict__ keep, float* __restrict__ cot, float* __restrict__ err_part,
                int slots, float* __restrict__ x0_out, int64_t n, int64_t chunk4, int64_t obs_repeat, float sa,
                float s1, float coef, const float* __restrict__ dsc) {
  step_scalars_k1(dsc, sa, s1, coef);
  const TweedieC tc = make_tc(s1, sa);
  __shared__ float red[32];
  const int64_t l = blockIdx.y;
  const int64_t n4 = n >> 2;
  const int64_t beg = (int64_t)blockIdx.x * chunk4;
  const int64_t end = min(beg + chunk4, n4);
  const float* xs = x + l * n;
  const float* es = eps + l * n;
  const float* ys = y + (l / obs_repeat) * n;
  float* cs = cot + l * n;
  float acc = 0.f;

  constexpr int U = 4;
  for (int64_t base = beg + threadIdx.x; base < end; base += (int64_t)kThreads * U) {
    float4 xv[U], ev[U], yv[U];
    uchar4 kv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + (int64_t)u * kThreads;
      if (i < end) {
        xv[u] = ld_stream4(xs + 4 * i);
        ev[u] = ld_stream4(es + 4 * i);
        yv[u] = __ldg(reinterpret_cast<const float4*>(ys) + i);
        if (MASK) kv[u] = __ldg(reinterpret_cast<const uchar4*>(keep) + i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + (int64_t)u * kThreads;
      if (i < end) {
        float4 x0, r, d;
        x0.x = tweedie(xv[u].x, ev[u].x, tc);
        x0.y = tweedie(xv[u].y, ev[u].y, tc);
        x0.z = tweedie(xv[u].z, ev[u].z, tc);
        x0.w = tweedie(xv[u].w, ev[u].w, tc);
        r.x = __fsub_rn(yv[u].x, x0.x);
        r.y = __fsub_rn(yv[u].y, x0.y);
        r.z = __fsub_rn(yv[u].z, x0.z);
        r.w = __fsub_rn(yv[u].w, x0.w);
        if (MASK) {
          r.x = kv[u].x ? r.x : 0.f;
          r.y = kv[u].y ? r.y : 0.f;
          r.z = kv[u].z ? r.z : 0.f;
          r.w = kv[u].w ? r.w : 0.f;
        }
        acc = fmaf(r.x, r.x, acc);
        acc = fmaf(r.y, r.y, acc);
        acc = fmaf(r.z, r.z, acc);
        acc = fmaf(r.w, r.w, acc);
        d.x = __fmul_rn(coef, r.x);
        d.y = __fmul_rn(coef, r.y);
        d.z = __fmul_rn(coef, r.z);
        d.w = __fmul_rn(coef, r.w);
        st_stream4(cs + 4 * i, d);
        if (WRITE_X0) st_stream4(x0_out + l * n + 4 * i, x0);
      }
    }
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) err_part[l * slots + blockIdx.x] = tot;
  if (blockIdx.x == 0)  // unused slots of this sample's row must read as zero in K2
    for (int i = gridDim.x + threadIdx.x; i < slots; i += kThreads) err_part[l * slots + i] = 0.f;
}

// scalar fallback for n % 4 != 0 (tiny / odd shapes)
template <bool MASK>
__global__ void __launch_bounds__(kThreads)
k1_pointwise_s(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ y,
               const uint8_t* __restrict__ keep, float* __restrict__ cot, float* __restrict__ err_part,
               int slots, float* __restrict__ x0_out, int64_t n, int64_t chunk, int64_t obs_repeat, float sa,
               float s1, float coef, const float* __restrict__ dsc) {
  step_scalars_k1(dsc, sa, s1, coef);
  const TweedieC tc = make_tc(s1, sa);
  __shared__ float red[32];
  const int64_t l = blockIdx.y;
  const int64_t beg = (int64_t)blockIdx.x * chunk, end = min(beg + chunk, n);
  float acc = 0.f;
  for (int64_t i = beg + threadIdx.x; i < end; i += kThreads) {
    const float x0 = tweedie(x[l * n + i], eps[l * n + i], tc);
    float r = __fsub_rn(y[(l / obs_repeat) * n + i], x0);
    if (MASK) r = keep[i] ? r : 0.f;
    acc = fmaf(r, r, acc);
    cot[l * n + i] = __fmul_rn(coef, r);
    if (x0_out) x0_out[l * n + i] = x0;
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) err_part[l * slots + blockIdx.x] = tot;
  if (blockIdx.x == 0)  // unused slots of this sample's row must read as zero in K2
    for (int i = gridDim.x + threadIdx.x; i < slots; i += kThreads) err_part[l * slots + i] = 0.f;
}

// ---- grid sizing: ONE wave.  A kernel this short (tens of microseconds) loses 10-25 % to a partial
// second wave, so the per-sample part count is chosen such that parts * L <= resident CTA slots.
// SM count of the CURRENT device (cached per device ordinal; the per-kernel occupancy figures cached at the launch
// sites are blocks per SM times this and assume what the north star names: one process per GPU of a homogeneous box).
int sm_count() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!cached[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}
template <typename K>
int resident_slots(K kernel) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1)
    per_sm = 2;
  return per_sm * sm_count();
}
// parts per sample.  Small launches (fewer than ~3 waves of 4-vectors-per-thread CTAs): exactly one wave,
// every CTA resident and equally loaded.  Large launches: many short CTAs (measured faster at L = 64:
// staggered CTA lifetimes keep more loads in flight than one long-lived wave).  Never more than `cap`
// (slots in the partial-sum row) and never less than one vector per thread.
int one_wave_parts(int slots, int64_t L, int64_t units_per_sample, int cap) {
  const int64_t by_size = (units_per_sample + kThreads - 1) / kThreads;
  int64_t fine = (units_per_sample + 4 * kThreads - 1) / (4 * kThreads);
  int64_t parts = fine * L >= 3 * (int64_t)slots ? fine : slots / L;
  if (parts > by_size) parts = by_size;
  if (parts > cap) parts = cap;
  return parts < 1 ? 1 : (int)parts;
}

int pointwise_parts(int64_t) { return kMaxParts; }  // size of the per-sample partial-sum row

int launch_pre_pointwise(const psx_op* op, const float* x, const float* eps, const float* y, int64_t L,
                         int64_t obs_repeat, float sa, float s1, float w, const float* dsc, float* cot, float* err_part,
                         float* x0_out, cudaStream_t st) {
  const int64_t n = op->n;
  const int slots = op->err_parts;
  const bool mask = op->kind == PSX_OP_MASK;
  const float coef = (float)((double)w / (double)sa);  // cot = (w / sa) * r, one rounding
  if (n % 4 == 0) {
#define PSX_LAUNCH(M, X)                                                                               \
  {                                                                                                    \
    static int rs = 0;                                                                                 \
    if (!rs) rs = resident_slots(k1_pointwise_v4<M, X>);                                               \
    const int parts = one_wave_parts(rs, L, n / 4, slots);                                             \
    const int64_t chunk4 = (n / 4 + parts - 1) / parts;                                                \
    k1_pointwise_v4<M, X><<<dim3(parts, (unsigned)L), kThreads, 0, st>>>(                              \
        x, eps, y, op->d_keep, cot, err_part, slots, x0_out, n, chunk4, obs_repeat, sa, s1, coef, dsc);\
  }
    if (mask) {
      if (x0_out) PSX_LAUNCH(true, true) else PSX_LAUNCH(true, false)
    } else {
      if (x0_out) PSX_LAUNCH(false, true) else PSX_LAUNCH(false, false)
    }
#undef PSX_LAUNCH
  } else {
    const int parts = one_wave_parts(2 * sm_count(), L, n, slots);
    const int64_t chunk = (n + parts - 1) / parts;
    dim3 grid(parts, (unsigned)L);
    if (mask)
      k1_pointwise_s<true><<<grid, kThreads, 0, st>>>(x, eps, y, op->d_keep, cot, err_part, slots, x0_out, n,
                                                      chunk, obs_repeat, sa, s1, coef, dsc);
    else
      k1_pointwise_s<false><<<grid, kThreads, 0, st>>>(x, eps, y, op->d_keep, cot, err_part, slots, x0_out, n,
                                                       chunk, obs_repeat, sa, s1, coef, dsc);
  }
  return check_cuda(cudaGetLastError(), "k1_pointwise launch");
}

// =========================================================================== K1: box super-resolution
// One thread per coarse pixel: F x F fine pixels of x_t / eps in, one y in, F x F cotangents out.
// A warp covers 32 consecutive coarse pixels of a row => 32*F contiguous floats per fine row.
template <int F>
struct RowVec;
template <> struct RowVec<2> { using T = float2; };
template <> struct RowVec<4> { using T = float4; };

template <int F>
__global__ void __launch_bounds__(kThreads)
k1_box(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ y,
       const uint8_t* __restrict__ keep, float* __restrict__ cot, float* __restrict__ err_part, int slots,
       float* __restrict__ x0_out, int planes, int H, int W, int64_t chunk, int64_t obs_repeat, float sa, float s1, float coef, const float* __restrict__ dsc) {
  step_scalars_k1(dsc, sa, s1, coef);
  const TweedieC tc = make_tc(s1, sa);
  __shared__ float red[32];
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
    float x0[F][F];
    float s = 0.f;
#pragma unroll
    for (int dy = 0; dy < F; ++dy) {
      float xv[F], ev[F];
      if constexpr (F == 4) {
        float4 a = ld_stream4(x + off + (int64_t)dy * W), b = ld_stream4(eps + off + (int64_t)dy * W);
        xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
        ev[0] = b.x; ev[1] = b.y; ev[2] = b.z; ev[3] = b.w;
      } else if constexpr (F == 2) {
        float2 a = __ldg(reinterpret_cast<const float2*>(x + off + (int64_t)dy * W));
        float2 b = __ldg(reinterpret_cast<const float2*>(eps + off + (int64_t)dy * W));
        xv[0] = a.x; xv[1] = a.y; ev[0] = b.x; ev[1] = b.y;
      } else {
#pragma unroll
        for (int dx = 0; dx < F; ++dx) {
          xv[dx] = __ldg(x + off + (int64_t)dy * W + dx);
          ev[dx] = __ldg(eps + off + (int64_t)dy * W + dx);
        }
      }
#pragma unroll
      for (int dx = 0; dx < F; ++dx) {
        x0[dy][dx] = tweedie(xv[dx], ev[dx], tc);
        s = __fadd_rn(s, x0[dy][dx]);
      }
    }
    const float avg = __fmul_rn(s, inv);  // F*F is a power of two for F = 2, 4, 8: exact scaling
    // composed operator mask o box (keep != nullptr): (A x0)[q] = 0 at a dropped coarse pixel -- the residual there is
    // y itself (it counts in |r|^2 exactly as in the oracle; a pre-masked observation makes it 0) and A^T drops it
    const bool kept = keep == nullptr || __ldg(keep + q) != 0;
    const float r = __fsub_rn(__ldg(y + (l / obs_repeat) * ny + q), kept ? avg : 0.f);
    acc = fmaf(r, r, acc);
    const float d = kept ? __fmul_rn(coef, __fmul_rn(r, inv)) : 0.f;
#pragma unroll
    for (int dy = 0; dy < F; ++dy) {
      if constexpr (F == 4) {
        st_stream4(cot + off + (int64_t)dy * W, make_float4(d, d, d, d));
        if (x0_out) st_stream4(x0_out + off + (int64_t)dy * W,
                               make_float4(x0[dy][0], x0[dy][1], x0[dy][2], x0[dy][3]));
      } else {
#pragma unroll
        for (int dx = 0; dx < F; ++dx) {
          cot[off + (int64_t)dy * W + dx] = d;
          if (x0_out) x0_out[off + (int64_t)dy * W + dx] = x0[dy][dx];
        }
      }
    }
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) err_part[l * slots + blockIdx.x] = tot;
  if (blockIdx.x == 0)  // unused slots of this sample's row must read as zero in K2
    for (int i = gridDim.x + threadIdx.x; i < slots; i += kThreads) err_part[l * slots + i] = 0.f;
}

// generic factor (not a power of two): division by F*F as the oracle does
__global__ void __launch_bounds__(kThreads)
k1_box_any(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ y,
           const uint8_t* __restrict__ keep, float* __restrict__ cot, float* __restrict__ err_part, int slots,
           float* __restrict__ x0_out, int planes, int H, int W, int F, int64_t chunk, int64_t obs_repeat, float sa, float s1, float coef, const float* __restrict__ dsc) {
  step_scalars_k1(dsc, sa, s1, coef);
  const TweedieC tc = make_tc(s1, sa);
  __shared__ float red[32];
  const int Hc = H / F, Wc = W / F;
  const int64_t n = (int64_t)planes * H * W, ny = (int64_t)planes * Hc * Wc;
  const int64_t l = blockIdx.y;
  const int64_t beg = (int64_t)blockIdx.x * chunk, end = min(beg + chunk, ny);
  const float ff = (float)(F * F);
  float acc = 0.f;
  for (int64_t q = beg + threadIdx.x; q < end; q += kThreads) {
    const int cx = (int)(q % Wc);
    const int64_t t = q / Wc;
    const int cy = (int)(t % Hc);
    const int pl = (int)(t / Hc);
    const int64_t off = l * n + ((int64_t)pl * H + (int64_t)cy * F) * W + (int64_t)cx * F;
    float s = 0.f;
    for (int dy = 0; dy < F; ++dy)
      for (int dx = 0; dx < F; ++dx) {
        const float v = tweedie(x[off + (int64_t)dy * W + dx], eps[off + (int64_t)dy * W + dx], tc);
        if (x0_out) x0_out[off + (int64_t)dy * W + dx] = v;
        s = __fadd_rn(s, v);
      }
    const bool kept = keep == nullptr || keep[q] != 0;
    const float r = __fsub_rn(y[(l / obs_repeat) * ny + q], kept ? __fdiv_rn(s, ff) : 0.f);
    acc = fmaf(r, r, acc);
    const float d = kept ? __fmul_rn(coef, __fdiv_rn(r, ff)) : 0.f;
    for (int dy = 0; dy < F; ++dy)
      for (int dx = 0; dx < F; ++dx) cot[off + (int64_t)dy * W + dx] = d;
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) err_part[l * slots + blockIdx.x] = tot;
  if (blockIdx.x == 0)  // unused slots of this sample's row must read as zero in K2
    for (int i = gridDim.x + threadIdx.x; i < slots; i += kThreads) err_part[l * slots + i] = 0.f;
}

int box_parts(int64_t) { return kMaxParts; }

int launch_pre_box(const psx_op* op, const float* x, const float* eps, const float* y, int64_t L,
                   int64_t obs_repeat, float sa, float s1, float w, const float* dsc, float* cot, float* err_part,
                   float* x0_out, cudaStream_t st) {
  const int slots = op->err_parts;
  const int F = op->factor;
  const float coef = (float)((double)w / (double)sa);
#define PSX_BOX(KERNEL, ...)                                                                          \
  {                                                                                                   \
    static int rs = 0;                                                                                \
    if (!rs) rs = resident_slots(KERNEL);                                                             \
    const int parts = one_wave_parts(rs, L, op->n_y, slots);                                          \
    const int64_t chunk = (op->n_y + parts - 1) / parts;                                              \
    KERNEL<<<dim3(parts, (unsigned)L), kThreads, 0, st>>>(x, eps, y, op->d_keep, cot, err_part, slots, x0_out, \
                                                          op->C, op->H, op->W, __VA_ARGS__ chunk,     \
                                                          obs_repeat, sa, s1, coef, dsc);             \
  }
  if (F == 4 && op->W % 4 == 0) PSX_BOX(k1_box<4>, )
  else if (F == 2) PSX_BOX(k1_box<2>, )
  else if (F == 8) PSX_BOX(k1_box<8>, )
  else PSX_BOX(k1_box_any, F, )
#undef PSX_BOX
  return check_cuda(cudaGetLastError(), "k1_box launch");
}

// =========================================================================== K2
// x_next = c_ell*x_t + c_s*x0 + std*z + gamma/(|r|+1e-9) * (cot - s1*vjp)
// Rounding order follows bridge_kernels.py:41 (mean), :59 (+ std*z), dps.py:121-122 (+ scale*grad).
// ---- in-kernel noise (production mode): Philox4x32-10 keyed by (seed), counter = (element group,
// step); four outputs -> four N(0,1) values by Box-Muller.  Layout and arithmetic are restated in oracle/philox.py
// (pinned to the Random123 known-answer vectors); the field depends on (seed, step, element index) only, so eager
// launches and graph replays draw identical noise.
__global__ void __launch_bounds__(kThreads)
k_philox_normal(float* __restrict__ out, int64_t total, uint64_t seed, uint64_t step) {
  const int64_t groups = (total + 3) >> 2;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    const float4 z = philox_normal4((uint64_t)g, seed, step);
    const float v[4] = {z.x, z.y, z.z, z.w};
    for (int c = 0; c < 4; ++c)
      if (4 * g + c < total) out[4 * g + c] = v[c];
  }
}

int launch_philox_normal(float* out, int64_t total, uint64_t seed, uint64_t step, cudaStream_t st) {
  int64_t b = ((total + 3) / 4 + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)sm_count() * 16;
  k_philox_normal<<<(unsigned)(b > cap ? cap : (b < 1 ? 1 : b)), kThreads, 0, st>>>(out, total, seed, step);
  return check_cuda(cudaGetLastError(), "k_philox_normal launch");
}

// ZMODE: 0 = no noise term, 1 = noise read from `z`, 2 = noise drawn in the kernel (Philox; `rng` = device
// {seed, step} when not null, else the by-value pair).
template <int ZMODE>
__global__ void __launch_bounds__(kThreads)
k2_post_v4(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ cot,
           const float* __restrict__ vjp, const float* __restrict__ z, const float* __restrict__ err_part,
           int err_parts, int64_t n, int64_t chunk4, float sa, float s1, float c_ell, float c_s,
           float sd,
           float gamma, float* __restrict__ x_next, float* __restrict__ err_out, const float* __restrict__ dsc,
           uint64_t seed, uint64_t step, const uint64_t* __restrict__ rng) {
  step_scalars_k2(dsc, sa, s1, c_ell, c_s, sd, gamma);
  if (ZMODE == 2 && rng != nullptr) {
    seed = rng[0];
    step = rng[1];
  }
  const TweedieC tc = make_tc(s1, sa);
  __shared__ float red[32];
  const int64_t l = blockIdx.y;
  const int64_t n4 = n >> 2;
  const int64_t beg = (int64_t)blockIdx.x * chunk4, end = min(beg + chunk4, n4);
  const int64_t so = l * n;
  constexpr int U = 2;
  constexpr int64_t kStride = (int64_t)kThreads * U;
  float4 xv[U], ev[U], dv[U], vv[U], zv[U];
#define PSX_K2_LOAD(BASE)                                         \
  _Pragma("unroll") for (int u = 0; u < U; ++u) {                 \
    const int64_t i = (BASE) + (int64_t)u * kThreads;             \
    if (i < end) {                                                \
      xv[u] = ld_stream4(x + so + 4 * i);                         \
      ev[u] = ld_stream4(eps + so + 4 * i);                       \
      dv[u] = ld_stream4(cot + so + 4 * i);                       \
      vv[u] = ld_stream4(vjp + so + 4 * i);                       \
      if (ZMODE == 1) zv[u] = ld_stream4(z + so + 4 * i);         \
    }                                                             \
  }
  // The first batch is requested BEFORE the reduction of the partial sums (a dependent global round trip plus a block
  // barrier): the guidance scale is only needed by the last operation of an element.
  int64_t base = beg + threadIdx.x;
  PSX_K2_LOAD(base)
  float scale = gamma;  // err_parts == 0: fixed guidance scale (PGDM); otherwise DPS: gamma / (|r| + 1e-9)
  if (err_parts > 0) {
    const float e2 = sum_parts(err_part + l * err_parts, err_parts, red);
    const float err = sqrtf(e2);
    scale = __fdiv_rn(gamma, __fadd_rn(err, 1e-9f));
    if (err_out && blockIdx.x == 0 && threadIdx.x == 0) err_out[l] = err;
  }
  for (; base < end; base += kStride) {
    if (ZMODE == 2) {  // generated while the loads are in flight
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = base + (int64_t)u * kThreads;
        if (i < end) zv[u] = philox_normal4((uint64_t)(l * n4 + i), seed, step);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + (int64_t)u * kThreads;
      if (i < end) {
        float4 o;
#define PSX_K2_LANE(c)                                                                         \
  {                                                                                            \
    const float x0 = tweedie(xv[u].c, ev[u].c, tc);                                        \
    float m = __fadd_rn(__fmul_rn(c_ell, xv[u].c), __fmul_rn(c_s, x0));                        \
    if (ZMODE != 0) m = __fadd_rn(m, __fmul_rn(sd, zv[u].c));                                  \
    const float g = __fadd_rn(dv[u].c, __fmul_rn(-s1, vv[u].c));                               \
    o.c = __fadd_rn(m, __fmul_rn(scale, g));                                                   \
  }
        PSX_K2_LANE(x) PSX_K2_LANE(y) PSX_K2_LANE(z) PSX_K2_LANE(w)
#undef PSX_K2_LANE
        st_stream4(x_next + so + 4 * i, o);
      }
    }
    PSX_K2_LOAD(base + kStride)
  }
#undef PSX_K2_LOAD
}

template <int ZMODE>
__global__ void __launch_bounds__(kThreads)
k2_post_s(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ cot,
          const float* __restrict__ vjp, const float* __restrict__ z, const float* __restrict__ err_part,
          int err_parts, int64_t n, int64_t chunk, float sa, float s1, float c_ell, float c_s, float sd,
          float gamma, float* __restrict__ x_next, float* __restrict__ err_out, const float* __restrict__ dsc,
          uint64_t seed, uint64_t step, const uint64_t* __restrict__ rng) {
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
  const int64_t beg = (int64_t)blockIdx.x * chunk, end = min(beg + chunk, n);
  for (int64_t i = beg + threadIdx.x; i < end; i += kThreads) {
    const int64_t j = l * n + i;
    const float x0 = tweedie(x[j], eps[j], tc);
    float m = __fadd_rn(__fmul_rn(c_ell, x[j]), __fmul_rn(c_s, x0));
    if (ZMODE == 1) m = __fadd_rn(m, __fmul_rn(sd, z[j]));
    if (ZMODE == 2) {  // ragged n: every element evaluates its group's Philox block and keeps one lane
      const float4 zz = philox_normal4((uint64_t)(j >> 2), seed, step);
      const int c = (int)(j & 3);
      m = __fadd_rn(m, __fmul_rn(sd, c == 0 ? zz.x : c == 1 ? zz.y : c == 2 ? zz.z : zz.w));
    }
    const float g = __fadd_rn(cot[j], __fmul_rn(-s1, vjp[j]));
    x_next[j] = __fadd_rn(m, __fmul_rn(scale, g));
  }
}

// zmode: 0 none, 1 tensor z, 2 Philox (seed, step by value, or from the device pair rng)
int launch_post(const float* x, const float* eps, const float* cot, const float* vjp, const float* z,
                const float* err_part, int err_parts, int64_t L, int64_t n, float sa, float s1,
                float c_ell, float c_s, float sd, float gamma, const float* dsc, float* x_next,
                float* err_out, int zmode, uint64_t seed, uint64_t step, const uint64_t* rng, cudaStream_t st) {
#define PSX_POST(KERNEL, UNITS)                                                                        \
  {                                                                                                    \
    static int rs = 0;                                                                                 \
    if (!rs) rs = resident_slots(KERNEL);                                                              \
    const int parts = one_wave_parts(rs, L, (UNITS), 1 << 20);                                         \
    const int64_t chunk = ((UNITS) + parts - 1) / parts;                                               \
    KERNEL<<<dim3(parts, (unsigned)L), kThreads, 0, st>>>(x, eps, cot, vjp, z, err_part, err_parts, n, \
                                                          chunk, sa, s1, c_ell, c_s, sd, gamma, x_next, \
                                                          err_out, dsc, seed, step, rng);              \
  }
  if (n % 4 == 0) {
    if (zmode == 2) PSX_POST(k2_post_v4<2>, n / 4) else if (zmode == 1) PSX_POST(k2_post_v4<1>, n / 4)
    else PSX_POST(k2_post_v4<0>, n / 4)
  } else {
    if (zmode == 2) PSX_POST(k2_post_s<2>, n) else if (zmode == 1) PSX_POST(k2_post_s<1>, n)
    else PSX_POST(k2_post_s<0>, n)
  }
#undef PSX_POST
  return check_cuda(cudaGetLastError(), "k2_post launch");
}

// K2 behind a K1 that has already written the bridge mean (psx_dps_pre_mean: the tensor-core blur):
//     x_next = (mean + std*z) + scale * (cot - s1*vjp)
// -- the arithmetic and roundings of k2_post_v4 from `m` on, so the pair (K1 with mean, this) is bit-identical to
// (K1, k2_post_v4).  20 instead of 24 B per element through the HBM-bound kernel.  ZMODE: 0 = no noise term, 1 = z read.
template <int ZMODE>
__global__ void __launch_bounds__(kThreads)
k2_post_mean(const float* __restrict__ mean, const float* __restrict__ cot, const float* __restrict__ vjp,
             const float* __restrict__ z, const float* __restrict__ err_part, int err_parts, int64_t n,
             int64_t chunk4, float s1, float sd, float gamma, float* __restrict__ x_next,
             float* __restrict__ err_out, const float* __restrict__ dsc) {
  if (dsc != nullptr) {
    s1 = __ldg(dsc + 1);
    sd = __ldg(dsc + 5);
    gamma = __ldg(dsc + 6);
  }
  __shared__ float red[32];
  const int64_t l = blockIdx.y;
  const int64_t n4 = n >> 2;
  const int64_t beg = (int64_t)blockIdx.x * chunk4, end = min(beg + chunk4, n4);
  const int64_t so = l * n;
  constexpr int U = PSX_K2M_U;
  constexpr int64_t kStride = (int64_t)kThreads * U;
  float4 mv[U], dv[U], vv[U], zv[U];
#define PSX_K2M_LOAD(BASE)                                        \
  _Pragma("unroll") for (int u = 0; u < U; ++u) {                 \
    const int64_t i = (BASE) + (int64_t)u * kThreads;             \
    if (i < end) {                                                \
      mv[u] = ld_stream4(mean + so + 4 * i);                      \
      dv[u] = ld_stream4(cot + so + 4 * i);                       \
      vv[u] = ld_stream4(vjp + so + 4 * i);                       \
      if (ZMODE == 1) zv[u] = ld_stream4(z + so + 4 * i);         \
    }                                                             \
  }
  int64_t base = beg + threadIdx.x;
  PSX_K2M_LOAD(base)  // ahead of the partial-sum reduction, as in k2_post_v4
  float scale = gamma;
  if (err_parts > 0) {
    const float e2 = sum_parts(err_part + l * err_parts, err_parts, red);
    const float err = sqrtf(e2);
    scale = __fdiv_rn(gamma, __fadd_rn(err, 1e-9f));
    if (err_out && blockIdx.x == 0 && threadIdx.x == 0) err_out[l] = err;
  }
  for (; base < end; base += kStride) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + (int64_t)u * kThreads;
      if (i < end) {
        float4 o;
#define PSX_K2M_LANE(c)                                                        \
  {                                                                            \
    float m = mv[u].c;                                                         \
    if (ZMODE != 0) m = __fadd_rn(m, __fmul_rn(sd, zv[u].c));                  \
    const float g = __fadd_rn(dv[u].c, __fmul_rn(-s1, vv[u].c));               \
    o.c = __fadd_rn(m, __fmul_rn(scale, g));                                   \
  }
        PSX_K2M_LANE(x) PSX_K2M_LANE(y) PSX_K2M_LANE(z) PSX_K2M_LANE(w)
#undef PSX_K2M_LANE
        st_stream4(x_next + so + 4 * i, o);
      }
    }
    PSX_K2M_LOAD(base + kStride)
  }
#undef PSX_K2M_LOAD
}

int launch_post_mean(const float* mean, const float* cot, const float* vjp, const float* z, const float* err_part,
                     int err_parts, int64_t L, int64_t n, float s1, float sd, float gamma, const float* dsc,
                     float* x_next, float* err_out, cudaStream_t st) {
  if (n % 4 != 0) return fail(PSX_ERR_UNSUPPORTED, "psx_dps_post_mean: n must be a multiple of 4");
#define PSX_POSTM(KERNEL)                                                                                   \
  {                                                                                                         \
    static int rs = 0;                                                                                      \
    if (!rs) rs = resident_slots(KERNEL);                                                                   \
    const int parts = one_wave_parts(rs, L, n / 4, 1 << 20);                                                \
    const int64_t chunk = (n / 4 + parts - 1) / parts;                                                      \
    KERNEL<<<dim3(parts, (unsigned)L), kThreads, 0, st>>>(mean, cot, vjp, z, err_part, err_parts, n, chunk, \
                                                          s1, sd, gamma, x_next, err_out, dsc);             \
  }
  if (z != nullptr) PSX_POSTM(k2_post_mean<1>) else PSX_POSTM(k2_post_mean<0>)
#undef PSX_POSTM
  return check_cuda(cudaGetLastError(), "k2_post_mean launch");
}

// =========================================================================== final Tweedie (+ posterior moments)
// One thread per pixel (quad), looping over the L local samples: x0 goes to the gather slot, the
// per-pixel sum / sum of squares are the all-reduce send buffers for the posterior mean / variance.
template <int V>
__global__ void __launch_bounds__(kThreads)
k_tweedie_final(const float* __restrict__ x, const float* __restrict__ eps, int64_t L, int64_t n, float sa,
                float s1, float* __restrict__ x0, float* __restrict__ sum, float* __restrict__ sumsq) {
  const TweedieC tc = make_tc(s1, sa);
  const int64_t nv = n / V;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (int64_t)gridDim.x * blockDim.x) {
    float s[V], q[V];
#pragma unroll
    for (int c = 0; c < V; ++c) s[c] = q[c] = 0.f;
    for (int64_t l = 0; l < L; ++l) {
      float xv[V], ev[V], o[V];
      if constexpr (V == 4) {
        float4 a = ld_stream4(x + l * n + 4 * i), b = ld_stream4(eps + l * n + 4 * i);
        xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
        ev[0] = b.x; ev[1] = b.y; ev[2] = b.z; ev[3] = b.w;
      } else {
        xv[0] = x[l * n + i];
        ev[0] = eps[l * n + i];
      }
#pragma unroll
      for (int c = 0; c < V; ++c) {
        o[c] = tweedie(xv[c], ev[c], tc);
        s[c] += o[c];
        q[c] = fmaf(o[c], o[c], q[c]);
      }
      if constexpr (V == 4) st_stream4(x0 + l * n + 4 * i, make_float4(o[0], o[1], o[2], o[3]));
      else x0[l * n + i] = o[0];
    }
    if (sum) {
      if constexpr (V == 4) st_stream4(sum + 4 * i, make_float4(s[0], s[1], s[2], s[3]));
      else sum[i] = s[0];
    }
    if (sumsq) {
      if constexpr (V == 4) st_stream4(sumsq + 4 * i, make_float4(q[0], q[1], q[2], q[3]));
      else sumsq[i] = q[0];
    }
  }
}

// When no moments are requested parallelise over samples as well (pure elementwise).
__global__ void __launch_bounds__(kThreads)
k_tweedie_flat(const float* __restrict__ x, const float* __restrict__ eps, int64_t total, float sa, float s1,
               float* __restrict__ x0) {
  const TweedieC tc = make_tc(s1, sa);
  const int64_t t4 = total >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = ld_stream4(x + 4 * i), b = ld_stream4(eps + 4 * i), o;
    o.x = tweedie(a.x, b.x, tc); o.y = tweedie(a.y, b.y, tc);
    o.z = tweedie(a.z, b.z, tc); o.w = tweedie(a.w, b.w, tc);
    st_stream4(x0 + 4 * i, o);
  }
  if (blockIdx.x == 0)
    for (int64_t i = (t4 << 2) + threadIdx.x; i < total; i += blockDim.x) x0[i] = tweedie(x[i], eps[i], tc);
}

int launch_tweedie(const float* x, const float* eps, int64_t L, int64_t n, float sa, float s1, float* x0,
                   float* sum, float* sumsq, cudaStream_t st) {
  if (!sum && !sumsq) {
    const int64_t total = L * n;
    int blocks = ceil_div(total / 4 + 1, kThreads);
    blocks = blocks > 148 * 16 ? 148 * 16 : blocks;
    k_tweedie_flat<<<blocks, kThreads, 0, st>>>(x, eps, total, sa, s1, x0);
  } else if (n % 4 == 0) {
    int blocks = ceil_div(n / 4, kThreads);
    k_tweedie_final<4><<<blocks, kThreads, 0, st>>>(x, eps, L, n, sa, s1, x0, sum, sumsq);
  } else {
    int blocks = ceil_div(n, kThreads);
    k_tweedie_final<1><<<blocks, kThreads, 0, st>>>(x, eps, L, n, sa, s1, x0, sum, sumsq);
  }
  return check_cuda(cudaGetLastError(), "tweedie launch");
}

// =========================================================================== PSLD tail + glue
template <bool HAS_Z, bool HAS_G>
__global__ void __launch_bounds__(kThreads)
k_bridge_update(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ z,
                const float* __restrict__ grad, int64_t total, float sa, float s1, float c_ell, float c_s,
                float sd, float gs, float* __restrict__ x_next) {
  const TweedieC tc = make_tc(s1, sa);
  const int64_t t4 = total >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 xv = ld_stream4(x + 4 * i), ev = ld_stream4(eps + 4 * i);
    float4 zv = make_float4(0.f, 0.f, 0.f, 0.f), gv = zv, o;
    if (HAS_Z) zv = ld_stream4(z + 4 * i);
    if (HAS_G) gv = ld_stream4(grad + 4 * i);
#define PSX_BU(c)                                                                       \
  {                                                                                     \
    const float x0 = tweedie(xv.c, ev.c, tc);                                           \
    float m = __fadd_rn(__fmul_rn(c_ell, xv.c), __fmul_rn(c_s, x0));                    \
    if (HAS_Z) m = __fadd_rn(m, __fmul_rn(sd, zv.c));                                   \
    if (HAS_G) m = __fadd_rn(m, __fmul_rn(gs, gv.c));                                   \
    o.c = m;                                                                            \
  }
    PSX_BU(x) PSX_BU(y) PSX_BU(z) PSX_BU(w)
    st_stream4(x_next + 4 * i, o);
  }
  if (blockIdx.x == 0)
    for (int64_t i = (t4 << 2) + threadIdx.x; i < total; i += blockDim.x) {
      const float x0 = tweedie(x[i], eps[i], tc);
      float m = __fadd_rn(__fmul_rn(c_ell, x[i]), __fmul_rn(c_s, x0));
      if (HAS_Z) m = __fadd_rn(m, __fmul_rn(sd, z[i]));
      if (HAS_G) m = __fadd_rn(m, __fmul_rn(gs, grad[i]));
      x_next[i] = m;
    }
#undef PSX_BU
}

int launch_bridge_update(const float* x, const float* eps, const float* z, const float* grad, int64_t total,
                         float sa, float s1, float c_ell, float c_s, float sd, float gs, float* x_next,
                         cudaStream_t st) {
  int blocks = ceil_div(total / 4 + 1, kThreads);
  blocks = blocks > sm_count() * 8 ? sm_count() * 8 : blocks;
  const bool hz = z != nullptr, hg = grad != nullptr;
#define PSX_GO(Z, G) \
  k_bridge_update<Z, G><<<blocks, kThreads, 0, st>>>(x, eps, z, grad, total, sa, s1, c_ell, c_s, sd, gs, x_next)
  if (hz && hg) PSX_GO(true, true); else if (hz) PSX_GO(true, false); else if (hg) PSX_GO(false, true);
  else PSX_GO(false, false);
#undef PSX_GO
  return check_cuda(cudaGetLastError(), "bridge_update launch");
}

template <bool HAS_C>
__global__ void __launch_bounds__(kThreads)
k_lincomb3(const float* __restrict__ a, float ca, const float* __restrict__ b, float cb,
           const float* __restrict__ c, float cc, float* __restrict__ out, int64_t total,
           const float* __restrict__ d_num, const float* __restrict__ d_den) {
  // device-side factor of cc (psx_lincomb3_dev): cc * num / den with torch's roundings (one multiply, one division);
  // a zero denominator (a residual norm of exactly 0) switches the term off instead of producing inf / nan
  if (d_num != nullptr) {
    const float den = d_den != nullptr ? __ldg(d_den) : 1.f;
    cc = den != 0.f ? __fdiv_rn(__fmul_rn(cc, __ldg(d_num)), den) : 0.f;
  }
  const int64_t t4 = total >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 av = ld_stream4(a + 4 * i), bv = ld_stream4(b + 4 * i);
    float4 o;
    o.x = __fadd_rn(__fmul_rn(ca, av.x), __fmul_rn(cb, bv.x)); o.y = __fadd_rn(__fmul_rn(ca, av.y), __fmul_rn(cb, bv.y));
    o.z = __fadd_rn(__fmul_rn(ca, av.z), __fmul_rn(cb, bv.z)); o.w = __fadd_rn(__fmul_rn(ca, av.w), __fmul_rn(cb, bv.w));
    if (HAS_C) {
      const float4 cv = ld_stream4(c + 4 * i);
      o.x = __fadd_rn(o.x, __fmul_rn(cc, cv.x)); o.y = __fadd_rn(o.y, __fmul_rn(cc, cv.y));
      o.z = __fadd_rn(o.z, __fmul_rn(cc, cv.z)); o.w = __fadd_rn(o.w, __fmul_rn(cc, cv.w));
    }
    st_stream4(out + 4 * i, o);
  }
  if (blockIdx.x == 0)
    for (int64_t i = (t4 << 2) + threadIdx.x; i < total; i += blockDim.x) {
      float o = __fadd_rn(__fmul_rn(ca, a[i]), __fmul_rn(cb, b[i]));
      if (HAS_C) o = __fadd_rn(o, __fmul_rn(cc, c[i]));
      out[i] = o;
    }
}

int launch_lincomb3(const float* a, float ca, const float* b, float cb, const float* c, float cc, float* out,
                    int64_t total, cudaStream_t st, const float* d_num, const float* d_den) {
  int blocks = ceil_div(total / 4 + 1, kThreads);
  blocks = blocks > sm_count() * 8 ? sm_count() * 8 : blocks;
  if (c) k_lincomb3<true><<<blocks, kThreads, 0, st>>>(a, ca, b, cb, c, cc, out, total, d_num, d_den);
  else k_lincomb3<false><<<blocks, kThreads, 0, st>>>(a, ca, b, cb, c, cc, out, total, d_num, d_den);
  return check_cuda(cudaGetLastError(), "lincomb3 launch");
}

// =========================================================================== ReSample kernels
__device__ __forceinline__ float div_const(float t, float d, float rcp) {  // correctly rounded t / d (see TweedieC)
  const float q0 = __fmul_rn(t, rcp);
  return __fmaf_rn(__fmaf_rn(-q0, d, t), rcp, q0);
}

template <bool HAS_Z>
__global__ void __launch_bounds__(kThreads)
k_ddim_eps(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ z, int64_t total,
           float sat, float soma, float oma, float sap, float dir, float sig, float* __restrict__ x_prev,
           float* __restrict__ pred, float* __restrict__ pseudo) {
  const float rcp = __frcp_rn(sat);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = x[i], ev = eps[i];
    const float p0 = div_const(__fsub_rn(xv, __fmul_rn(soma, ev)), sat, rcp);
    const float ps = div_const(__fsub_rn(xv, __fmul_rn(oma, ev)), sat, rcp);
    float o = __fadd_rn(__fmul_rn(sap, p0), __fmul_rn(dir, ev));
    if (HAS_Z) o = __fadd_rn(o, __fmul_rn(sig, z[i]));
    x_prev[i] = o;
    if (pred) pred[i] = p0;
    if (pseudo) pseudo[i] = ps;
  }
}

__global__ void __launch_bounds__(kThreads)
k_stoch_resample(const float* __restrict__ p, const float* __restrict__ xt, const float* __restrict__ nz,
                 int64_t total, float cp, float cx, float den, float kn, float* __restrict__ out) {
  const float rcp = __frcp_rn(den);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float num = __fadd_rn(__fmul_rn(cp, p[i]), __fmul_rn(cx, xt[i]));
    out[i] = __fadd_rn(div_const(num, den, rcp), __fmul_rn(nz[i], kn));
  }
}

// AdamW with torch's operation order (torch/optim/adamw.py, single-tensor path):
//   p *= 1 - lr*wd;  m += (g - m)*(1 - b1);  v = v*b2 + (1 - b2)*g*g
//   p += -(lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(kThreads)
k_adamw(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
        int64_t total, float decay, float w1, float b2, float w2, float step_size, float bc2_sqrt, float aeps,
        int* __restrict__ flags, int flag_in, const float* __restrict__ loss_parts, int64_t n_parts,
        float loss_scale, float loss_thr) {
  __shared__ float red[32];
  if (flags) {
    if (flags[flag_in]) {
      if (blockIdx.x == 0 && threadIdx.x == 0) flags[1 - flag_in] = 1;
      return;
    }
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float pi = __fmul_rn(p[i], decay);
    const float mi = __fadd_rn(m[i], __fmul_rn(__fsub_rn(gi, m[i]), w1));
    const float vi = __fadd_rn(__fmul_rn(v[i], b2), __fmul_rn(__fmul_rn(w2, gi), gi));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bc2_sqrt), aeps);
    pi = __fadd_rn(pi, __fmul_rn(-step_size, __fdiv_rn(mi, denom)));
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
  if (flags && blockIdx.x == 0) {  // fixed-order sum of the loss partials of THIS iteration
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n_parts; i += kThreads) acc += loss_parts[i];
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) flags[1 - flag_in] = (__fmul_rn(tot, loss_scale) < loss_thr) ? 1 : 0;
  }
}

static int ew_blocks(int64_t total) {
  int b = ceil_div(total, kThreads);
  return b > sm_count() * 8 ? sm_count() * 8 : (b < 1 ? 1 : b);
}

int launch_ddim_eps(const float* x, const float* eps, const float* z, int64_t total, float sat, float soma,
                    float oma, float sap, float dir, float sig, float* x_prev, float* pred, float* pseudo,
                    cudaStream_t st) {
  if (z) k_ddim_eps<true><<<ew_blocks(total), kThreads, 0, st>>>(x, eps, z, total, sat, soma, oma, sap, dir, sig, x_prev, pred, pseudo);
  else k_ddim_eps<false><<<ew_blocks(total), kThreads, 0, st>>>(x, eps, z, total, sat, soma, oma, sap, dir, sig, x_prev, pred, pseudo);
  return check_cuda(cudaGetLastError(), "ddim_eps launch");
}

int launch_stoch_resample(const float* p, const float* xt, const float* nz, int64_t total, float cp, float cx,
                          float den, float kn, float* out, cudaStream_t st) {
  k_stoch_resample<<<ew_blocks(total), kThreads, 0, st>>>(p, xt, nz, total, cp, cx, den, kn, out);
  return check_cuda(cudaGetLastError(), "stochastic_resample launch");
}

int launch_adamw(float* p, const float* g, float* m, float* v, int64_t total, float lr, float b1, float b2,
                 float aeps, float wd, int step, int* flags, int flag_in, const float* loss_parts, int64_t n_parts,
                 float loss_scale, float loss_thr, cudaStream_t st) {
  // scalar prep in double like Python floats, then rounded once to fp32
  const double bc1 = 1.0 - std::pow((double)b1, step), bc2 = 1.0 - std::pow((double)b2, step);
  const float decay = (float)(1.0 - (double)lr * (double)wd);
  const float w1 = (float)(1.0 - (double)b1), w2 = (float)(1.0 - (double)b2);
  const float step_size = (float)((double)lr / bc1), bc2_sqrt = (float)std::sqrt(bc2);
  k_adamw<<<ew_blocks(total), kThreads, 0, st>>>(p, g, m, v, total, decay, w1, b2, w2, step_size, bc2_sqrt, aeps,
                                                 flags, flag_in, loss_parts, n_parts, loss_scale, loss_thr);
  return check_cuda(cudaGetLastError(), "adamw launch");
}

// =========================================================================== stand-alone operators
__global__ void __launch_bounds__(kThreads)
k_mask_apply(const float* __restrict__ in, const uint8_t* __restrict__ keep, float* __restrict__ out,
             int64_t n, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = keep[i % n] ? in[i] : 0.f;
}

__global__ void __launch_bounds__(kThreads)
k_box_apply(const float* __restrict__ x, float* __restrict__ y, const uint8_t* __restrict__ keep, int64_t n_y,
            int planes, int H, int W, int F, int64_t total_y) {
  const int Hc = H / F, Wc = W / F;
  const bool pow2 = (F & (F - 1)) == 0;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total_y;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int cx = (int)(q % Wc);
    const int64_t t = q / Wc;
    const int cy = (int)(t % Hc);
    const int64_t pl = t / Hc;  // plane index including the sample
    const int64_t off = (pl * H + (int64_t)cy * F) * W + (int64_t)cx * F;
    float s = 0.f;
    for (int dy = 0; dy < F; ++dy)
      for (int dx = 0; dx < F; ++dx) s = __fadd_rn(s, x[off + (int64_t)dy * W + dx]);
    const float avg = pow2 ? __fmul_rn(s, 1.0f / (float)(F * F)) : __fdiv_rn(s, (float)(F * F));
    y[q] = (keep == nullptr || keep[q % n_y]) ? avg : 0.f;
  }
}

__global__ void __launch_bounds__(kThreads)
k_box_adjoint(const float* __restrict__ y, float* __restrict__ x, const uint8_t* __restrict__ keep, int64_t n_y,
              int planes, int H, int W, int F, int64_t total_x) {
  const int Hc = H / F, Wc = W / F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_x;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int px = (int)(i % W);
    const int64_t t = i / W;
    const int py = (int)(t % H);
    const int64_t pl = t / H;
    const int64_t q = (pl * Hc + py / F) * Wc + px / F;
    x[i] = (keep == nullptr || keep[q % n_y]) ? __fdiv_rn(y[q], (float)(F * F)) : 0.f;
  }
}

__global__ void __launch_bounds__(kThreads)
k_gather(const float* __restrict__ in, const int64_t* __restrict__ idx, float* __restrict__ out, int64_t n,
         int64_t m, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = i / m, j = i - l * m;
    out[i] = in[l * n + __ldg(idx + j)];
  }
}

__global__ void __launch_bounds__(kThreads)
k_scatter(const float* __restrict__ in, const int64_t* __restrict__ idx, float* __restrict__ out, int64_t n,
          int64_t m, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = i / m, j = i - l * m;
    out[l * n + __ldg(idx + j)] = in[i];
  }
}

int launch_gather(bool scatter, const float* in, const int64_t* idx, float* out, int64_t L, int64_t n,
                  int64_t m, cudaStream_t st) {
  const int64_t total = L * m;
  int blocks = ceil_div(total, kThreads);
  blocks = blocks > 148 * 16 ? 148 * 16 : (blocks < 1 ? 1 : blocks);
  if (scatter) {
    int rc = check_cuda(cudaMemsetAsync(out, 0, (size_t)L * n * sizeof(float), st), "scatter memset");
    if (rc) return rc;
    if (total > 0) k_scatter<<<blocks, kThreads, 0, st>>>(in, idx, out, n, m, total);
  } else if (total > 0) {
    k_gather<<<blocks, kThreads, 0, st>>>(in, idx, out, n, m, total);
  }
  return check_cuda(cudaGetLastError(), "gather/scatter launch");
}

int launch_op_pointwise(const psx_op* op, bool adjoint, const float* in, float* out, int64_t L,
                        cudaStream_t st) {
  const int64_t total_x = L * op->n, total_y = L * op->n_y;
  auto blocks_for = [](int64_t t) {
    int b = ceil_div(t, kThreads);
    return b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b);
  };
  switch (op->kind) {
    case PSX_OP_IDENTITY:
      if (in != out)
        return check_cuda(cudaMemcpyAsync(out, in, total_x * sizeof(float), cudaMemcpyDeviceToDevice, st),
                          "identity copy");
      return PSX_OK;
    case PSX_OP_MASK:
      k_mask_apply<<<blocks_for(total_x), kThreads, 0, st>>>(in, op->d_keep, out, op->n, total_x);
      break;
    case PSX_OP_BOX:
      if (!adjoint)
        k_box_apply<<<blocks_for(total_y), kThreads, 0, st>>>(in, out, op->d_keep, op->n_y, op->C, op->H, op->W,
                                                              op->factor, total_y);
      else
        k_box_adjoint<<<blocks_for(total_x), kThreads, 0, st>>>(in, out, op->d_keep, op->n_y, op->C, op->H, op->W,
                                                                op->factor, total_x);
      break;
    default:
      return fail(PSX_ERR_INVALID, "launch_op_pointwise: wrong operator kind");
  }
  return check_cuda(cudaGetLastError(), "operator launch");
}

}  // namespace psx
