// psx_conv.cu -- shared-memory-tiled convolution kernels (sm_100a):
//   separable blur  A = V . H  (rows then columns), adjoint = H^T . V^T,
//   sparse-tap depthwise 2-D correlation (motion PSFs),
//   and K1 (psx_dps_pre) built from them:
//     rows<TWEEDIE>  : x0 = (x_t - s1 eps)/sa on the fly, h1 = H x0           -> workspace
//     cols<RESIDUAL> : r = y - V h1, |r|^2 partials, h2 = V^T r  (strip-local) -> workspace (in place)
//     rows<COT>      : cot = w * H^T h2 / sa                                   -> d_cot
// All passes are zero-padded "same" cross-correlations; halos are zero-filled in shared memory.
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
                                            bool tw, float s1, float sa) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (gr >= H) return v;
  const int64_t g = plane + (int64_t)gr * W + gc;
  if (vec_ok && gc >= 0 && gc + 3 < W) {
    v = ld_stream4(in + g);
    if (tw) {
      const float4 e = ld_stream4(eps + g);
      v.x = tweedie(v.x, e.x, s1, sa); v.y = tweedie(v.y, e.y, s1, sa);
      v.z = tweedie(v.z, e.z, s1, sa); v.w = tweedie(v.w, e.w, s1, sa);
    }
  } else {
    float t[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      t[c] = 0.f;
      if (gc + c >= 0 && gc + c < W) {
        t[c] = in[g + c];
        if (tw) t[c] = tweedie(t[c], eps[g + c], s1, sa);
      }
    }
    v = make_float4(t[0], t[1], t[2], t[3]);
  }
  return v;
}

// ------------------------------------------------------------------------------------------ rows
// grid = (ceil(W/TW), ceil(H/32), planes_total); dynamic smem = 16 * pitch2 float2.
// The tile is stored ROW-PAIR INTERLEAVED: smem2[rp][s] = (row 2rp, row 2rp+1) at image column
// c0 + taps.lo + s, so that one FFMA2 advances two output rows with an aligned register pair for
// every tap.  pitch2 % 16 == 2 makes the LDS.128 of 8 consecutive row pairs hit 8 distinct 16 B banks.
template <int MODE>
__global__ void __launch_bounds__(kThreads)
conv_rows(const float* __restrict__ in, const float* __restrict__ eps, float* __restrict__ out, int H, int W,
          int TW, int pitch2, const __grid_constant__ Taps taps, float sa, float s1, float coef) {
  extern __shared__ __align__(16) float2 smem2[];
  const int c0 = blockIdx.x * TW;
  const int r0 = blockIdx.y * kRowTH;
  const int64_t plane = (int64_t)blockIdx.z * H * W;
  const int in_w4 = (TW + taps.k) >> 2;
  const bool vec_ok = (W & 3) == 0;

  for (int idx = threadIdx.x; idx < (kRowTH / 2) * in_w4; idx += kThreads) {
    const int rp = idx / in_w4, c4 = idx - rp * in_w4;
    const int gc = c0 + taps.lo + 4 * c4, gr = r0 + 2 * rp;
    const float4 a = load_row4(in, eps, plane, gr, gc, H, W, vec_ok, MODE == ROWS_TWEEDIE, s1, sa);
    const float4 b = load_row4(in, eps, plane, gr + 1, gc, H, W, vec_ok, MODE == ROWS_TWEEDIE, s1, sa);
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
  if (!op->col_tc) return fail(PSX_ERR_UNSUPPORTED, "separable blur: image too tall for the column kernel");
  op->err_parts = op->C * ceil_div(op->W, op->col_tc);
  return PSX_OK;
}

template <int MODE>
static int run_rows(const psx_op* op, const Taps& t, const float* in, const float* eps, float* out,
                    int64_t planes, float sa, float s1, float w, cudaStream_t st) {
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
  conv_rows<MODE><<<grid, kThreads, smem, st>>>(in, eps, out, op->H, op->W, TW, pitch, t, sa, s1, coef);
  return check_cuda(cudaGetLastError(), "conv_rows launch");
}

template <bool RESIDUAL>
static int run_cols(const psx_op* op, const Taps& tf, const Taps& ta, const float* in, const float* y,
                    float* out, float* err_part, int64_t planes, int64_t obs_repeat, cudaStream_t st) {
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

int launch_pre_sepblur(const psx_op* op, const float* x, const float* eps, const float* y, int64_t L,
                       int64_t obs_repeat, float sa, float s1, float w, float* cot, float* err_part,
                       float* x0_out, float* ws, cudaStream_t st) {
  if (x0_out) return fail(PSX_ERR_UNSUPPORTED, "psx_dps_pre: d_x0_out is not produced for blur operators");
  const int64_t planes = L * op->C;
  int rc = run_rows<ROWS_TWEEDIE>(op, op->fh, x, eps, ws, planes, sa, s1, w, st);
  if (rc) return rc;
  rc = run_cols<true>(op, op->fv, op->av, ws, y, ws, err_part, planes, obs_repeat, st);
  if (rc) return rc;
  return run_rows<ROWS_COT>(op, op->ah, ws, nullptr, cot, planes, sa, s1, w, st);
}

// ------------------------------------------------------------------------------------------ sparse 2-D
enum C2Mode { C2_PLAIN = 0, C2_RESIDUAL = 1, C2_COT = 2 };
constexpr int kC2Tile = 32;
constexpr int kC2Chunk = 512;

// grid = (tilesX, tilesY, planes); each thread: 4 rows of one column of the 32 x 32 tile.
// ADJ = false: out[p] = sum_j w_j in[p + off_j];  ADJ = true: out[p] = sum_j w_j in[p - off_j].
template <int MODE, bool ADJ>
__global__ void __launch_bounds__(kThreads)
conv2d_sparse(const float* __restrict__ in, const float* __restrict__ eps, const float* __restrict__ y,
              float* __restrict__ out, float* __restrict__ err_part, const Tap2D* __restrict__ taps,
              int ntaps, int C, int H, int W, int kh, int kw, int64_t obs_repeat, float sa, float s1,
              float wgt) {
  extern __shared__ __align__(16) float smem[];
  __shared__ Tap2D stap[kC2Chunk];
  __shared__ float red[32];
  const int hy = kh / 2, hx = kw / 2;
  const int tw = kC2Tile + kw - 1, th = kC2Tile + kh - 1;
  const int r0 = blockIdx.y * kC2Tile, c0 = blockIdx.x * kC2Tile;
  const int64_t pl = blockIdx.z;
  const int64_t plane = pl * H * W;

  for (int idx = threadIdx.x; idx < th * tw; idx += kThreads) {
    const int r = idx / tw, c = idx - r * tw;
    const int gr = r0 - hy + r, gc = c0 - hx + c;
    float v = 0.f;
    if (gr >= 0 && gr < H && gc >= 0 && gc < W) {
      v = in[plane + (int64_t)gr * W + gc];
      if (MODE == C2_RESIDUAL) v = tweedie(v, eps[plane + (int64_t)gr * W + gc], s1, sa);
    }
    smem[idx] = v;
  }

  const int lx = threadIdx.x & 31, ly = (threadIdx.x >> 5) * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int base = 0; base < ntaps; base += kC2Chunk) {
    __syncthreads();
    const int cnt = min(kC2Chunk, ntaps - base);
    for (int i = threadIdx.x; i < cnt; i += kThreads) stap[i] = taps[base + i];
    __syncthreads();
    for (int i = 0; i < cnt; ++i) {
      const Tap2D t = stap[i];
      const int dy = ADJ ? -t.dy : t.dy, dx = ADJ ? -t.dx : t.dx;
      const float* p = smem + (ly + hy + dy) * tw + (lx + hx + dx);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(t.w, p[j * tw], acc[j]);
    }
  }

  float e2 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int gr = r0 + ly + j, gc = c0 + lx;
    if (gr < H && gc < W) {
      const int64_t g = plane + (int64_t)gr * W + gc;
      if (MODE == C2_RESIDUAL) {
        const int64_t yo = ((pl / C) / obs_repeat * C + pl % C) * (int64_t)H * W + (int64_t)gr * W + gc;
        const float r = __fsub_rn(__ldg(y + yo), acc[j]);
        e2 = fmaf(r, r, e2);
        out[g] = r;
      } else if (MODE == C2_COT) {
        out[g] = __fmul_rn(wgt, acc[j]);
      } else {
        out[g] = acc[j];
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

int conv2d_err_parts(const psx_op* op) {
  return op->C * ceil_div(op->W, kC2Tile) * ceil_div(op->H, kC2Tile);
}

template <int MODE, bool ADJ>
static int run_conv2d(const psx_op* op, const float* in, const float* eps, const float* y, float* out,
                      float* err_part, int64_t planes, int64_t obs_repeat, float sa, float s1, float w,
                      cudaStream_t st) {
  const size_t smem = (size_t)(kC2Tile + op->kh - 1) * (kC2Tile + op->kw - 1) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(conv2d_sparse<MODE, ADJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr_done = true;
  }
  dim3 grid(ceil_div(op->W, kC2Tile), ceil_div(op->H, kC2Tile), (unsigned)planes);
  const float wgt = MODE == C2_COT ? (float)((double)w / (double)sa) : w;
  conv2d_sparse<MODE, ADJ><<<grid, kThreads, smem, st>>>(in, eps, y, out, err_part, op->d_taps_f,
                                                         op->n_taps2d, op->C, op->H, op->W, op->kh, op->kw,
                                                         obs_repeat, sa, s1, wgt);
  return check_cuda(cudaGetLastError(), "conv2d_sparse launch");
}

int launch_pre_conv2d(const psx_op* op, const float* x, const float* eps, const float* y, int64_t L,
                      int64_t obs_repeat, float sa, float s1, float w, float* cot, float* err_part,
                      float* x0_out, float* ws, cudaStream_t st) {
  if (x0_out) return fail(PSX_ERR_UNSUPPORTED, "psx_dps_pre: d_x0_out is not produced for blur operators");
  const int64_t planes = L * op->C;
  int rc = run_conv2d<C2_RESIDUAL, false>(op, x, eps, y, ws, err_part, planes, obs_repeat, sa, s1, w, st);
  if (rc) return rc;
  return run_conv2d<C2_COT, true>(op, ws, nullptr, nullptr, cot, nullptr, planes, 1, sa, s1, w, st);
}

// ------------------------------------------------------------------------------------------ stand-alone A / A^T
int launch_op_pointwise(const psx_op* op, bool adjoint, const float* in, float* out, int64_t L, cudaStream_t st);

int launch_op(const psx_op* op, bool adjoint, const float* in, float* out, int64_t L, float* ws,
              cudaStream_t st) {
  const int64_t planes = L * op->C;
  switch (op->kind) {
    case PSX_OP_SEPBLUR: {
      if (!adjoint) {  // y = V(H(x))
        int rc = run_rows<ROWS_PLAIN>(op, op->fh, in, nullptr, ws, planes, 1.f, 0.f, 1.f, st);
        if (rc) return rc;
        return run_cols<false>(op, op->fv, op->fv, ws, nullptr, out, nullptr, planes, 1, st);
      }
      int rc = run_cols<false>(op, op->av, op->av, in, nullptr, ws, nullptr, planes, 1, st);
      if (rc) return rc;
      return run_rows<ROWS_PLAIN>(op, op->ah, ws, nullptr, out, planes, 1.f, 0.f, 1.f, st);
    }
    case PSX_OP_CONV2D:
      if (!adjoint)
        return run_conv2d<C2_PLAIN, false>(op, in, nullptr, nullptr, out, nullptr, planes, 1, 1.f, 0.f, 1.f, st);
      return run_conv2d<C2_PLAIN, true>(op, in, nullptr, nullptr, out, nullptr, planes, 1, 1.f, 0.f, 1.f, st);
    default:
      return launch_op_pointwise(op, adjoint, in, out, L, st);
  }
}

}  // namespace psx
