// psx_api.cu -- extern "C" entry points of libpsx (see include/psx.h).
// Host-side validation, operator descriptors, tap preparation; all device work is asynchronous
// on the caller's stream.  No CPU implementation exists behind these calls.
#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstring>
#include <new>
#include <vector>

#include "psx_common.cuh"

namespace psx {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
// Every kernel launch of the library is followed by exactly one check_cuda(cudaGetLastError(), "<kernel> launch"):
// that is where the library counts its own launches (psx_kernel_launches; a launch recorded into a CUDA graph is
// counted once, at capture).
std::atomic<long long> g_kernel_launches{0};
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) {
    const size_t n = std::strlen(what);
    if (n >= 6 && std::strcmp(what + n - 6, "launch") == 0) g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    return PSX_OK;
  }
  g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return PSX_ERR_CUDA;
}

int pointwise_parts(int64_t n);
int box_parts(int64_t ny);
int launch_post(const float*, const float*, const float*, const float*, const float*, const float*, int,
                int64_t, int64_t, float, float, float, float, float, float, const float*, float*, float*, int,
                uint64_t, uint64_t, const uint64_t*, cudaStream_t);
int launch_philox_normal(float*, int64_t, uint64_t, uint64_t, cudaStream_t);
int launch_pre_pointwise_bf16(const psx_op*, const void*, const void*, const float*, int64_t, int64_t, float, float, float,
                              const float*, void*, float*, cudaStream_t);
int launch_pre_box_bf16(const psx_op*, const void*, const void*, const float*, int64_t, int64_t, float, float, float,
                        const float*, void*, float*, cudaStream_t);
int launch_post_bf16(const void*, const void*, const void*, const void*, const void*, const float*, int, int64_t, int64_t,
                     float, float, float, float, float, float, const float*, void*, float*, int, uint64_t, uint64_t,
                     const uint64_t*, cudaStream_t);
int launch_tweedie(const float*, const float*, int64_t, int64_t, float, float, float*, float*, float*,
                   cudaStream_t);
int launch_add_noise(float*, const float*, int64_t, float, float, cudaStream_t);
int launch_image_to_u8(const float*, uint8_t*, int64_t, int, int64_t, cudaStream_t);
int launch_image_from_u8(const uint8_t*, float*, int64_t, int, int64_t, cudaStream_t);
int launch_gather(bool, const float*, const int64_t*, float*, int64_t, int64_t, int64_t, cudaStream_t);
int launch_bridge_update(const float*, const float*, const float*, const float*, int64_t, float, float, float,
                         float, float, float, float*, cudaStream_t);
int launch_lincomb3(const float*, float, const float*, float, const float*, float, float*, int64_t, cudaStream_t,
                    const float* d_num = nullptr, const float* d_den = nullptr);
int launch_ddim_eps(const float*, const float*, const float*, int64_t, float, float, float, float, float, float,
                    float*, float*, float*, cudaStream_t);
int launch_stoch_resample(const float*, const float*, const float*, int64_t, float, float, float, float, float*,
                          cudaStream_t);
int launch_adamw(float*, const float*, float*, float*, int64_t, float, float, float, float, float, int, int*, int,
                 const float*, int64_t, float, float, cudaStream_t);

// Taps whose magnitude is below 2^-30 of the largest tap are dropped at the two ends: their total
// contribution (<= k * 2^-30 * max|w| * max|x|) is far below half an fp32 ulp of the result.
static int make_taps(const float* w, int k, bool flip, Taps* out) {
  std::vector<float> v(w, w + k);
  if (flip) for (int i = 0; i < k / 2; ++i) std::swap(v[i], v[k - 1 - i]);
  float mx = 0.f;
  for (float t : v) mx = std::fmax(mx, std::fabs(t));
  int first = 0, last = k - 1;
  const float thr = mx * 9.313225746154785e-10f;  // 2^-30
  while (first < last && std::fabs(v[first]) <= thr) ++first;
  while (last > first && std::fabs(v[last]) <= thr) --last;
  int lo = first - k / 2;                       // offset of the first kept tap
  int lo4 = (int)std::floor(lo / 4.0) * 4;      // rows kernel wants lo % 4 == 0
  int front = lo - lo4;
  int kept = last - first + 1;
  int kk = (front + kept + 7) & ~7;
  if (kk > PSX_MAX_TAPS + 9) return fail(PSX_ERR_INVALID, "too many taps");
  std::memset(out, 0, sizeof(Taps));
  out->k = kk;
  out->lo = lo4;
  for (int i = 0; i < kept; ++i) out->ww[front + i] = make_float2(v[first + i], v[first + i]);
  return PSX_OK;
}

namespace {
EnvOpts g_env;
std::atomic<bool> g_env_read{false};
std::mutex g_env_mu;
void read_env() {
  EnvOpts e;
  e.no_pipe = getenv("PSX_NO_PIPE") != nullptr;
  e.no_fast16 = getenv("PSX_NO_FAST16") != nullptr;
  e.no_tc = getenv("PSX_NO_TC") != nullptr;
  e.tc_persist = getenv("PSX_TC_PERSIST") != nullptr;
  e.no_c2v2 = getenv("PSX_NO_C2V2") != nullptr;
  const char* lag = getenv("PSX_MEAN_LAG_NS");
  e.mean_lag_ns = lag ? atoi(lag) : 5000;
  e.fused = getenv("PSX_FUSED") != nullptr;
  const char* sp = getenv("PSX_SPLIT");
  e.split = sp ? atoi(sp) : 0;
  g_env = e;
}
}  // namespace
const EnvOpts& env_opts() {
  if (!g_env_read.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lock(g_env_mu);
    if (!g_env_read.load(std::memory_order_relaxed)) {
      read_env();
      g_env_read.store(true, std::memory_order_release);
    }
  }
  return g_env;
}

}  // namespace psx

using namespace psx;

extern "C" {

PSX_API int psx_abi_version(void) { return PSX_ABI_VERSION; }
PSX_API long long psx_kernel_launches(void) { return g_kernel_launches.load(std::memory_order_relaxed); }
PSX_API void psx_reload_env(void) {
  std::lock_guard<std::mutex> lock(g_env_mu);
  read_env();
  g_env_read.store(true, std::memory_order_release);
}
PSX_API const char* psx_last_error(void) { return g_last_error.c_str(); }

static psx_op* new_op(int kind) {
  psx_op* op = new (std::nothrow) psx_op();
  if (op) {
    std::memset(op, 0, sizeof(psx_op));
    op->kind = kind;
  }
  return op;
}

PSX_API int psx_op_create_identity(int64_t n, psx_op** out) {
  PSX_REQUIRE(out && n > 0, "psx_op_create_identity: n must be positive");
  psx_op* op = new_op(PSX_OP_IDENTITY);
  if (!op) return fail(PSX_ERR_INVALID, "out of host memory");
  op->n = op->n_y = n;
  op->err_parts = pointwise_parts(n);
  *out = op;
  return PSX_OK;
}

PSX_API int psx_op_create_mask(int64_t n, const uint8_t* d_keep, psx_op** out) {
  PSX_REQUIRE(out && n > 0 && d_keep, "psx_op_create_mask: need n > 0 and a device keep-mask");
  psx_op* op = new_op(PSX_OP_MASK);
  if (!op) return fail(PSX_ERR_INVALID, "out of host memory");
  op->n = op->n_y = n;
  op->d_keep = d_keep;
  op->err_parts = pointwise_parts(n);
  *out = op;
  return PSX_OK;
}

PSX_API int psx_op_create_box(int C, int H, int W, int factor, psx_op** out) {
  PSX_REQUIRE(out && C > 0 && H > 0 && W > 0 && factor > 0, "psx_op_create_box: bad shape");
  PSX_REQUIRE(H % factor == 0 && W % factor == 0, "psx_op_create_box: H and W must be multiples of factor");
  psx_op* op = new_op(PSX_OP_BOX);
  if (!op) return fail(PSX_ERR_INVALID, "out of host memory");
  op->C = C; op->H = H; op->W = W; op->factor = factor;
  op->n = (int64_t)C * H * W;
  op->n_y = op->n / ((int64_t)factor * factor);
  op->err_parts = box_parts(op->n_y);
  *out = op;
  return PSX_OK;
}

PSX_API int psx_op_create_box_masked(int C, int H, int W, int factor, const uint8_t* d_keep, psx_op** out) {
  PSX_REQUIRE(d_keep != nullptr, "psx_op_create_box_masked: need a device keep-mask over the coarse grid");
  if (int rc = psx_op_create_box(C, H, W, factor, out)) return rc;
  (*out)->d_keep = d_keep;  // C * (H / factor) * (W / factor) bytes, caller-owned
  return PSX_OK;
}

PSX_API int psx_op_create_sepblur(int C, int H, int W, const float* h_taps_h, int kh, const float* h_taps_v,
                          int kv, psx_op** out) {
  PSX_REQUIRE(out && C > 0 && H > 0 && W > 0, "psx_op_create_sepblur: bad shape");
  PSX_REQUIRE(h_taps_h && h_taps_v, "psx_op_create_sepblur: null taps");
  PSX_REQUIRE(kh > 0 && kv > 0 && (kh & 1) && (kv & 1) && kh <= PSX_MAX_TAPS && kv <= PSX_MAX_TAPS,
              "psx_op_create_sepblur: tap counts must be odd and <= PSX_MAX_TAPS");
  psx_op* op = new_op(PSX_OP_SEPBLUR);
  if (!op) return fail(PSX_ERR_INVALID, "out of host memory");
  op->C = C; op->H = H; op->W = W;
  op->n = op->n_y = (int64_t)C * H * W;
  int rc = make_taps(h_taps_h, kh, false, &op->fh);
  if (!rc) rc = make_taps(h_taps_v, kv, false, &op->fv);
  if (!rc) rc = make_taps(h_taps_h, kh, true, &op->ah);
  if (!rc) rc = make_taps(h_taps_v, kv, true, &op->av);
  if (!rc) rc = sepblur_plan(op);
  if (rc) { delete op; return rc; }
  tcblur_plan(op);
  // Side streams / events for the split K1 (launch_pre_sepblur).  Without a device (host-only use of the
  // descriptor) or on any failure the descriptor simply has none and K1 runs on the caller's stream alone.
  op->aux_mu = new (std::nothrow) std::mutex();
  bool ok = op->aux_mu && cudaEventCreateWithFlags(&op->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < 3; ++i) {
    ok = cudaStreamCreateWithFlags(&op->aux_stream[i], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&op->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
    if (ok) op->aux_n = i + 1;
  }
  if (!ok) cudaGetLastError();
  *out = op;
  return PSX_OK;
}

// Row-segment form of a 2-D PSF for one direction: sign = +1 forward (out[p] = sum w in[p + off]),
// -1 adjoint (the flipped PSF).
// Row-segment / column-segment form of a 2-D PSF for one direction: sign = +1 forward (out[p] = sum w in[p + off]),
// -1 adjoint (the flipped PSF).  Row segments start at a column offset that is a multiple of 4 (vector loads along
// the row); column segments start at the exact first tap (the window slides down one row per load).  The cheaper
// form (chunks + segments) is kept: shallow motion lines are row-like, steep ones and camera shake column-like.
struct PsfBuild {
  std::vector<RowSeg> segs;
  std::vector<float> w;
  int dy_lo = 1 << 20, dy_hi = -(1 << 20), dx_lo = 1 << 20, dx_hi = -(1 << 20);
};

static void build_rows(const float* k, int kh, int kw, int sign, PsfBuild& b) {
  for (int row = 0; row < kh; ++row) {
    const int jy = sign > 0 ? row : kh - 1 - row;  // ascending dy in both directions
    int lo = kw, hi = -1;
    for (int jx = 0; jx < kw; ++jx)
      if (k[jy * kw + jx] != 0.f) {
        lo = jx < lo ? jx : lo;
        hi = jx > hi ? jx : hi;
      }
    if (hi < 0) continue;
    const int dy = sign * (jy - kh / 2);
    const int dxa = sign * ((sign > 0 ? lo : hi) - kw / 2), dxb = sign * ((sign > 0 ? hi : lo) - kw / 2);  // dxa <= dxb
    const int dx0 = dxa >= 0 ? (dxa / 4) * 4 : -(((-dxa) + 3) / 4) * 4;  // floor to a multiple of 4
    const int nch = (dxb - dx0 + 4) / 4;
    RowSeg sg;
    sg.dy = (int16_t)dy; sg.dx0 = (int16_t)dx0; sg.nch = (int16_t)nch; sg.w4_off = (uint16_t)(b.w.size() / 4);
    for (int i = 0; i < 4 * nch; ++i) {
      const int jx = sign * (dx0 + i) + kw / 2;     // tap at column offset dx  <->  kernel column sign*dx + kw/2
      b.w.push_back(jx >= 0 && jx < kw ? k[jy * kw + jx] : 0.f);
    }
    b.segs.push_back(sg);
    b.dy_lo = std::min(b.dy_lo, dy); b.dy_hi = std::max(b.dy_hi, dy);
    b.dx_lo = std::min(b.dx_lo, dx0); b.dx_hi = std::max(b.dx_hi, dx0 + 4 * nch);
  }
}

static void build_cols(const float* k, int kh, int kw, int sign, PsfBuild& b) {
  for (int col = 0; col < kw; ++col) {
    const int jx = sign > 0 ? col : kw - 1 - col;  // ascending dx in both directions
    int lo = kh, hi = -1;
    for (int jy = 0; jy < kh; ++jy)
      if (k[jy * kw + jx] != 0.f) {
        lo = jy < lo ? jy : lo;
        hi = jy > hi ? jy : hi;
      }
    if (hi < 0) continue;
    const int dx = sign * (jx - kw / 2);
    const int dya = sign * ((sign > 0 ? lo : hi) - kh / 2), dyb = sign * ((sign > 0 ? hi : lo) - kh / 2);  // dya <= dyb
    const int nch = (dyb - dya + 4) / 4;
    RowSeg sg;
    sg.dy = (int16_t)dya; sg.dx0 = (int16_t)dx; sg.nch = (int16_t)nch; sg.w4_off = (uint16_t)(b.w.size() / 4);
    for (int i = 0; i < 4 * nch; ++i) {
      const int jy = sign * (dya + i) + kh / 2;
      b.w.push_back(jy >= 0 && jy < kh ? k[jy * kw + jx] : 0.f);
    }
    b.segs.push_back(sg);
    b.dy_lo = std::min(b.dy_lo, dya); b.dy_hi = std::max(b.dy_hi, dya + 4 * nch);
    b.dx_lo = std::min(b.dx_lo, dx); b.dx_hi = std::max(b.dx_hi, dx);
  }
}

static int make_psf2d(const float* k, int kh, int kw, int sign, Psf2D* out) {
  PsfBuild rows, cols;
  build_rows(k, kh, kw, sign, rows);
  build_cols(k, kh, kw, sign, cols);
  if (rows.segs.empty()) return fail(PSX_ERR_INVALID, "psx_op_create_conv2d: kernel is all zeros");
  const char* force = getenv("PSX_PSF_FORM");  // "rows" / "cols": tests and measurements
  // cost model from the measured kernels (profiles/README.md): one segment costs about as much as one 4-tap chunk
  bool use_cols = cols.w.size() / 4 + cols.segs.size() < rows.w.size() / 4 + rows.segs.size();
  if (force) use_cols = force[0] == 'c';
  const PsfBuild& b = use_cols ? cols : rows;
  if (b.w.size() / 4 > 65535) return fail(PSX_ERR_UNSUPPORTED, "psx_op_create_conv2d: PSF too large");
  out->cols = use_cols ? 1 : 0;
  out->nseg = (int)b.segs.size();
  out->nw4 = (int)(b.w.size() / 4);
  out->dy_lo = b.dy_lo; out->dy_hi = b.dy_hi; out->dx_lo = b.dx_lo; out->dx_hi = b.dx_hi;
  out->h_v2 = nullptr;
  if (!use_cols && b.segs.size() <= (size_t)kC2MaxSeg && b.w.size() <= (size_t)kC2MaxTap) {
    C2Params* v2 = new (std::nothrow) C2Params();
    if (v2) {
      v2->nseg = (int)b.segs.size();
      v2->pad = 0;
      for (size_t i = 0; i < b.segs.size(); ++i) {
        const RowSeg& sg = b.segs[i];
        v2->seg[i] = make_int2((int)(uint16_t)sg.dy | ((int)sg.dx0 << 16), (int)(uint16_t)sg.nch | ((int)sg.w4_off * 4 << 16));
      }
      for (size_t i = 0; i < b.w.size(); ++i) v2->ww[i] = make_float2(b.w[i], b.w[i]);
      out->h_v2 = v2;
    }
  }
  int rc = check_cuda(cudaMalloc(&out->d_segs, b.segs.size() * sizeof(RowSeg)), "cudaMalloc PSF segments");
  if (!rc) rc = check_cuda(cudaMalloc(&out->d_w4, b.w.size() * sizeof(float)), "cudaMalloc PSF taps");
  if (!rc) rc = check_cuda(cudaMemcpy(out->d_segs, b.segs.data(), b.segs.size() * sizeof(RowSeg), cudaMemcpyHostToDevice), "copy PSF segments");
  if (!rc) rc = check_cuda(cudaMemcpy(out->d_w4, b.w.data(), b.w.size() * sizeof(float), cudaMemcpyHostToDevice), "copy PSF taps");
  return rc;
}

static void free_psf2d(Psf2D* p) {
  if (p->d_segs) cudaFree(p->d_segs);
  if (p->d_w4) cudaFree(p->d_w4);
  delete p->h_v2;
  p->h_v2 = nullptr;
  p->d_segs = nullptr;
  p->d_w4 = nullptr;
}

PSX_API int psx_op_create_conv2d(int C, int H, int W, const float* h_kernel, int kh, int kw, psx_op** out) {
  PSX_REQUIRE(out && C > 0 && H > 0 && W > 0 && h_kernel, "psx_op_create_conv2d: bad arguments");
  PSX_REQUIRE(kh > 0 && kw > 0 && (kh & 1) && (kw & 1) && kh <= PSX_MAX_TAPS && kw <= PSX_MAX_TAPS,
              "psx_op_create_conv2d: kernel sizes must be odd and <= PSX_MAX_TAPS");
  psx_op* op = new_op(PSX_OP_CONV2D);
  if (!op) return fail(PSX_ERR_INVALID, "out of host memory");
  op->C = C; op->H = H; op->W = W; op->kh = kh; op->kw = kw;
  op->n = op->n_y = (int64_t)C * H * W;
  int rc = make_psf2d(h_kernel, kh, kw, +1, &op->psf_f);
  if (!rc) rc = make_psf2d(h_kernel, kh, kw, -1, &op->psf_a);
  if (rc) { free_psf2d(&op->psf_f); free_psf2d(&op->psf_a); delete op; return rc; }
  op->err_parts = conv2d_err_parts(op);
  *out = op;
  return PSX_OK;
}

PSX_API int psx_op_destroy(psx_op* op) {
  if (!op) return PSX_OK;
  if (op->kind == PSX_OP_CONV2D) {
    free_psf2d(&op->psf_f);
    free_psf2d(&op->psf_a);
  }
  for (int i = 0; i < op->aux_n; ++i) {
    cudaStreamDestroy(op->aux_stream[i]);
    cudaEventDestroy(op->ev_join[i]);
  }
  if (op->ev_fork) cudaEventDestroy(op->ev_fork);
  tcblur_release(op);
  delete op->aux_mu;
  delete op;
  return PSX_OK;
}

PSX_API int psx_op_kind(const psx_op* op) { return op ? op->kind : -1; }
PSX_API int64_t psx_op_x_numel(const psx_op* op) { return op ? op->n : 0; }
PSX_API int64_t psx_op_y_numel(const psx_op* op) { return op ? op->n_y : 0; }
PSX_API int psx_op_err_parts(const psx_op* op) { return op ? op->err_parts : 0; }

PSX_API size_t psx_op_workspace_bytes(const psx_op* op, int64_t L) {
  if (!op || L <= 0) return 0;
  // separable blur: h1 and h2 of the CUDA-core K1 live in two regions -- the row-pair-interleaved h2 of a strip
  // covers cells of OTHER strips' h1 (offset 2 * column inside the double row), so writing it in place raced with
  // strips that had not been fetched yet once persistent CTAs drifted apart (L >= 32)
  if (op->kind == PSX_OP_SEPBLUR) return 2 * (size_t)L * op->n * sizeof(float);
  if (op->kind == PSX_OP_CONV2D) return (size_t)L * op->n * sizeof(float);
  return 0;
}

static int check_ws(const psx_op* op, int64_t L, void* ws, size_t bytes) {
  const size_t need = psx_op_workspace_bytes(op, L);
  if (need && (!ws || bytes < need)) return fail(PSX_ERR_INVALID, "workspace too small (see psx_op_workspace_bytes)");
  return PSX_OK;
}

PSX_API int psx_op_apply(const psx_op* op, const float* d_x, float* d_y, int64_t L, void* ws, size_t ws_bytes,
                 void* stream) {
  PSX_REQUIRE(op && d_x && d_y && L > 0, "psx_op_apply: null argument or L <= 0");
  if (int rc = check_ws(op, L, ws, ws_bytes)) return rc;
  return launch_op(op, false, d_x, d_y, L, (float*)ws, (cudaStream_t)stream);
}

PSX_API int psx_op_adjoint(const psx_op* op, const float* d_y, float* d_x, int64_t L, void* ws, size_t ws_bytes,
                   void* stream) {
  PSX_REQUIRE(op && d_x && d_y && L > 0, "psx_op_adjoint: null argument or L <= 0");
  if (int rc = check_ws(op, L, ws, ws_bytes)) return rc;
  return launch_op(op, true, d_y, d_x, L, (float*)ws, (cudaStream_t)stream);
}

PSX_API int psx_observe(const psx_op* op, const float* d_x, const float* d_noise, float noise_scale, float noise_shift,
                        float* d_y, int64_t L, void* ws, size_t ws_bytes, void* stream) {
  PSX_REQUIRE(op && d_x && d_y && L > 0, "psx_observe: null argument or L <= 0");
  PSX_REQUIRE(d_x != d_y, "psx_observe: d_y must not alias d_x");
  PSX_REQUIRE(std::isfinite(noise_scale) && std::isfinite(noise_shift), "psx_observe: non-finite noise parameter");
  if (int rc = check_ws(op, L, ws, ws_bytes)) return rc;
  if (int rc = launch_op(op, false, d_x, d_y, L, (float*)ws, (cudaStream_t)stream)) return rc;
  if (!d_noise) return PSX_OK;
  return launch_add_noise(d_y, d_noise, L * op->n_y, noise_scale, noise_shift, (cudaStream_t)stream);
}

PSX_API int psx_add_noise(float* d_y, const float* d_noise, int64_t numel, float noise_scale, float noise_shift,
                          void* stream) {
  PSX_REQUIRE(d_y && d_noise && numel > 0, "psx_add_noise: null pointer or empty tensor");
  PSX_REQUIRE(std::isfinite(noise_scale) && std::isfinite(noise_shift), "psx_add_noise: non-finite noise parameter");
  return launch_add_noise(d_y, d_noise, numel, noise_scale, noise_shift, (cudaStream_t)stream);
}

PSX_API int psx_image_to_u8(const float* d_chw, uint8_t* d_hwc, int64_t images, int C, int H, int W, void* stream) {
  PSX_REQUIRE(d_chw && d_hwc, "psx_image_to_u8: null pointer");
  PSX_REQUIRE(images > 0 && C > 0 && C <= 16 && H > 0 && W > 0, "psx_image_to_u8: bad sizes");
  return launch_image_to_u8(d_chw, d_hwc, images, C, (int64_t)H * W, (cudaStream_t)stream);
}

PSX_API int psx_image_from_u8(const uint8_t* d_hwc, float* d_chw, int64_t images, int C, int H, int W, void* stream) {
  PSX_REQUIRE(d_chw && d_hwc, "psx_image_from_u8: null pointer");
  PSX_REQUIRE(images > 0 && C > 0 && C <= 16 && H > 0 && W > 0, "psx_image_from_u8: bad sizes");
  return launch_image_from_u8(d_hwc, d_chw, images, C, (int64_t)H * W, (cudaStream_t)stream);
}

PSX_API int psx_gather(const float* d_in, const int64_t* d_idx, float* d_out, int64_t L, int64_t n, int64_t m,
                       void* stream) {
  PSX_REQUIRE(d_in && d_out && (d_idx || m == 0), "psx_gather: null pointer");
  PSX_REQUIRE(L > 0 && n > 0 && m >= 0 && m <= n, "psx_gather: bad sizes");
  return launch_gather(false, d_in, d_idx, d_out, L, n, m, (cudaStream_t)stream);
}

PSX_API int psx_scatter(const float* d_in, const int64_t* d_idx, float* d_out, int64_t L, int64_t n, int64_t m,
                        void* stream) {
  PSX_REQUIRE(d_out && ((d_in && d_idx) || m == 0), "psx_scatter: null pointer");
  PSX_REQUIRE(L > 0 && n > 0 && m >= 0 && m <= n, "psx_scatter: bad sizes");
  return launch_gather(true, d_in, d_idx, d_out, L, n, m, (cudaStream_t)stream);
}

// Shared body of psx_dps_pre / psx_dps_pre_dev: with d_row the kernels read their step scalars from device memory.
static int dps_pre_impl(const char* who, const psx_op* op, const float* d_x_t, const float* d_eps, const float* d_y,
                        int64_t L, int64_t obs_repeat, float sqrt_acp, float sqrt_1m_acp, float lik_weight,
                        const float* d_row, float* d_cot, float* d_err_part, float* d_x0_out, void* ws,
                        size_t ws_bytes, void* stream, float* d_mean_out = nullptr, float c_ell = 0.f,
                        float c_s = 0.f, const float* d_z = nullptr, float sd = 0.f) {
  if (!(op && d_x_t && d_eps && d_y && d_cot && d_err_part)) return fail(PSX_ERR_INVALID, std::string(who) + ": null pointer");
  if (!(L > 0 && L <= 65535)) return fail(PSX_ERR_INVALID, std::string(who) + ": L must be in [1, 65535]");
  if (!(obs_repeat > 0)) return fail(PSX_ERR_INVALID, std::string(who) + ": obs_repeat must be positive");
  if (int rc = check_ws(op, L, ws, ws_bytes)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (d_mean_out && !fuses_mean(op, L))
    return fail(PSX_ERR_UNSUPPORTED, std::string(who) + ": this operator's K1 does not emit the bridge mean at this "
                                                        "batch size (psx_op_fuses_mean is 0)");
  switch (op->kind) {
    case PSX_OP_IDENTITY:
    case PSX_OP_MASK:
      return launch_pre_pointwise(op, d_x_t, d_eps, d_y, L, obs_repeat, sqrt_acp, sqrt_1m_acp, lik_weight, d_row,
                                  d_cot, d_err_part, d_x0_out, st);
    case PSX_OP_BOX:
      return launch_pre_box(op, d_x_t, d_eps, d_y, L, obs_repeat, sqrt_acp, sqrt_1m_acp, lik_weight, d_row, d_cot,
                            d_err_part, d_x0_out, st);
    case PSX_OP_SEPBLUR:
      return launch_pre_sepblur(op, d_x_t, d_eps, d_y, L, obs_repeat, sqrt_acp, sqrt_1m_acp, lik_weight, d_row,
                                d_cot, d_err_part, d_x0_out, (float*)ws, st, false, d_mean_out, c_ell, c_s, d_z, sd);
    case PSX_OP_CONV2D:
      return launch_pre_conv2d(op, d_x_t, d_eps, d_y, L, obs_repeat, sqrt_acp, sqrt_1m_acp, lik_weight, d_row,
                               d_cot, d_err_part, d_x0_out, (float*)ws, st);
  }
  return fail(PSX_ERR_UNSUPPORTED, std::string(who) + ": unknown operator kind");
}

PSX_API int psx_dps_pre(const psx_op* op, const float* d_x_t, const float* d_eps, const float* d_y, int64_t L,
                int64_t obs_repeat, float sqrt_acp, float sqrt_1m_acp, float lik_weight, float* d_cot,
                float* d_err_part, float* d_x0_out, void* ws, size_t ws_bytes, void* stream) {
  PSX_REQUIRE(sqrt_acp > 0.f && std::isfinite(sqrt_acp) && std::isfinite(sqrt_1m_acp) && std::isfinite(lik_weight),
              "psx_dps_pre: non-finite or non-positive schedule scalar");
  return dps_pre_impl("psx_dps_pre", op, d_x_t, d_eps, d_y, L, obs_repeat, sqrt_acp, sqrt_1m_acp, lik_weight,
                      nullptr, d_cot, d_err_part, d_x0_out, ws, ws_bytes, stream);
}

PSX_API int psx_dps_pre_dev(const psx_op* op, const float* d_x_t, const float* d_eps, const float* d_y, int64_t L,
                            int64_t obs_repeat, const float* d_step_row, float* d_cot, float* d_err_part,
                            float* d_x0_out, void* ws, size_t ws_bytes, void* stream) {
  PSX_REQUIRE(d_step_row != nullptr, "psx_dps_pre_dev: null step row");
  return dps_pre_impl("psx_dps_pre_dev", op, d_x_t, d_eps, d_y, L, obs_repeat, 1.f, 0.f, 1.f, d_step_row, d_cot,
                      d_err_part, d_x0_out, ws, ws_bytes, stream);
}

PSX_API int psx_op_fuses_mean(const psx_op* op, int64_t L) {
  return op != nullptr && L > 0 && fuses_mean(op, L) ? 1 : 0;
}

PSX_API int psx_dps_pre_mean(const psx_op* op, const float* d_x_t, const float* d_eps, const float* d_y, int64_t L,
                             int64_t obs_repeat, float sqrt_acp, float sqrt_1m_acp, float lik_weight, float c_ell,
                             float c_s, const float* d_z, float std_, float* d_cot, float* d_err_part,
                             float* d_mean_out, void* ws, size_t ws_bytes, void* stream) {
  PSX_REQUIRE(d_mean_out != nullptr, "psx_dps_pre_mean: null d_mean_out");
  PSX_REQUIRE(sqrt_acp > 0.f && std::isfinite(sqrt_acp) && std::isfinite(sqrt_1m_acp) && std::isfinite(lik_weight) &&
                  std::isfinite(c_ell) && std::isfinite(c_s) && std::isfinite(std_),
              "psx_dps_pre_mean: non-finite or non-positive schedule scalar");
  return dps_pre_impl("psx_dps_pre_mean", op, d_x_t, d_eps, d_y, L, obs_repeat, sqrt_acp, sqrt_1m_acp, lik_weight,
                      nullptr, d_cot, d_err_part, nullptr, ws, ws_bytes, stream, d_mean_out, c_ell, c_s, d_z, std_);
}

PSX_API int psx_dps_pre_mean_dev(const psx_op* op, const float* d_x_t, const float* d_eps, const float* d_y,
                                 int64_t L, int64_t obs_repeat, const float* d_step_row, const float* d_z,
                                 float* d_cot, float* d_err_part, float* d_mean_out, void* ws, size_t ws_bytes,
                                 void* stream) {
  PSX_REQUIRE(d_step_row != nullptr && d_mean_out != nullptr, "psx_dps_pre_mean_dev: null step row / d_mean_out");
  return dps_pre_impl("psx_dps_pre_mean_dev", op, d_x_t, d_eps, d_y, L, obs_repeat, 1.f, 0.f, 1.f, d_step_row, d_cot,
                      d_err_part, nullptr, ws, ws_bytes, stream, d_mean_out, 0.f, 0.f, d_z, 0.f);
}

PSX_API int psx_dps_post_mean(const float* d_mean, const float* d_cot, const float* d_vjp, const float* d_z,
                              const float* d_err_part, int err_parts, int64_t L, int64_t n, float sqrt_1m_acp,
                              float std_, float gamma, float* d_x_next, float* d_err_out, void* stream) {
  PSX_REQUIRE(d_mean && d_cot && d_vjp && d_x_next, "psx_dps_post_mean: null pointer");
  PSX_REQUIRE(L > 0 && L <= 65535 && n > 0 && err_parts >= 0, "psx_dps_post_mean: bad sizes");
  PSX_REQUIRE((err_parts > 0) == (d_err_part != nullptr), "psx_dps_post_mean: d_err_part and err_parts must agree");
  PSX_REQUIRE(std::isfinite(sqrt_1m_acp) && std::isfinite(std_) && std::isfinite(gamma),
              "psx_dps_post_mean: non-finite scalar");
  return launch_post_mean(d_mean, d_cot, d_vjp, std_ == 0.f ? nullptr : d_z, d_err_part, err_parts, L, n,
                          sqrt_1m_acp, std_, gamma, nullptr, d_x_next, d_err_out, (cudaStream_t)stream);
}

PSX_API int psx_dps_post_mean_dev(const float* d_mean, const float* d_cot, const float* d_vjp, const float* d_z,
                                  const float* d_err_part, int err_parts, int64_t L, int64_t n,
                                  const float* d_step_row, float* d_x_next, float* d_err_out, void* stream) {
  PSX_REQUIRE(d_mean && d_cot && d_vjp && d_x_next && d_step_row, "psx_dps_post_mean_dev: null pointer");
  PSX_REQUIRE(L > 0 && L <= 65535 && n > 0 && err_parts >= 0, "psx_dps_post_mean_dev: bad sizes");
  PSX_REQUIRE((err_parts > 0) == (d_err_part != nullptr),
              "psx_dps_post_mean_dev: d_err_part and err_parts must agree");
  return launch_post_mean(d_mean, d_cot, d_vjp, d_z, d_err_part, err_parts, L, n, 0.f, 0.f, 0.f, d_step_row, d_x_next,
                          d_err_out, (cudaStream_t)stream);
}

PSX_API int psx_dps_post(const float* d_x_t, const float* d_eps, const float* d_cot, const float* d_vjp,
                 const float* d_z, const float* d_err_part, int err_parts, int64_t L, int64_t n,
                 float sqrt_acp, float sqrt_1m_acp, float c_ell, float c_s, float std_, float gamma,
                 float* d_x_next, float* d_err_out, void* stream) {
  PSX_REQUIRE(d_x_t && d_eps && d_cot && d_vjp && d_x_next, "psx_dps_post: null pointer");
  PSX_REQUIRE(L > 0 && L <= 65535 && n > 0 && err_parts >= 0, "psx_dps_post: bad sizes");
  PSX_REQUIRE((err_parts > 0) == (d_err_part != nullptr), "psx_dps_post: d_err_part and err_parts must agree");
  PSX_REQUIRE(d_z || std_ == 0.f, "psx_dps_post: d_z may be NULL only when std == 0");
  PSX_REQUIRE(sqrt_acp > 0.f && std::isfinite(sqrt_acp) && std::isfinite(c_ell) && std::isfinite(c_s) &&
                  std::isfinite(std_) && std::isfinite(gamma),
              "psx_dps_post: non-finite scalar");
  return launch_post(d_x_t, d_eps, d_cot, d_vjp, std_ == 0.f ? nullptr : d_z, d_err_part, err_parts, L, n,
                     sqrt_acp, sqrt_1m_acp, c_ell, c_s, std_, gamma, nullptr, d_x_next, d_err_out,
                     std_ == 0.f ? 0 : 1, 0, 0, nullptr, (cudaStream_t)stream);
}

PSX_API int psx_dps_post_dev(const float* d_x_t, const float* d_eps, const float* d_cot, const float* d_vjp,
                             const float* d_z, const float* d_err_part, int err_parts, int64_t L, int64_t n,
                             const float* d_step_row, float* d_x_next, float* d_err_out, void* stream) {
  PSX_REQUIRE(d_x_t && d_eps && d_cot && d_vjp && d_z && d_x_next && d_step_row, "psx_dps_post_dev: null pointer");
  PSX_REQUIRE(L > 0 && L <= 65535 && n > 0 && err_parts >= 0, "psx_dps_post_dev: bad sizes");
  PSX_REQUIRE((err_parts > 0) == (d_err_part != nullptr), "psx_dps_post_dev: d_err_part and err_parts must agree");
  return launch_post(d_x_t, d_eps, d_cot, d_vjp, d_z, d_err_part, err_parts, L, n, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f,
                     d_step_row, d_x_next, d_err_out, 1, 0, 0, nullptr, (cudaStream_t)stream);
}

PSX_API int psx_dps_post_philox(const float* d_x_t, const float* d_eps, const float* d_cot, const float* d_vjp,
                                const float* d_err_part, int err_parts, int64_t L, int64_t n, float sqrt_acp,
                                float sqrt_1m_acp, float c_ell, float c_s, float std_, float gamma, uint64_t seed,
                                uint64_t step, float* d_x_next, float* d_err_out, void* stream) {
  PSX_REQUIRE(d_x_t && d_eps && d_cot && d_vjp && d_x_next, "psx_dps_post_philox: null pointer");
  PSX_REQUIRE(L > 0 && L <= 65535 && n > 0 && err_parts >= 0, "psx_dps_post_philox: bad sizes");
  PSX_REQUIRE((err_parts > 0) == (d_err_part != nullptr), "psx_dps_post_philox: d_err_part and err_parts must agree");
  PSX_REQUIRE(sqrt_acp > 0.f && std::isfinite(sqrt_acp) && std::isfinite(c_ell) && std::isfinite(c_s) &&
                  std::isfinite(std_) && std::isfinite(gamma),
              "psx_dps_post_philox: non-finite scalar");
  return launch_post(d_x_t, d_eps, d_cot, d_vjp, nullptr, d_err_part, err_parts, L, n, sqrt_acp, sqrt_1m_acp, c_ell,
                     c_s, std_, gamma, nullptr, d_x_next, d_err_out, std_ == 0.f ? 0 : 2, seed, step, nullptr,
                     (cudaStream_t)stream);
}

PSX_API int psx_dps_post_philox_dev(const float* d_x_t, const float* d_eps, const float* d_cot, const float* d_vjp,
                                    const float* d_err_part, int err_parts, int64_t L, int64_t n,
                                    const float* d_step_row, const uint64_t* d_seed_step, float* d_x_next,
                                    float* d_err_out, void* stream) {
  PSX_REQUIRE(d_x_t && d_eps && d_cot && d_vjp && d_x_next && d_step_row && d_seed_step,
              "psx_dps_post_philox_dev: null pointer");
  PSX_REQUIRE(L > 0 && L <= 65535 && n > 0 && err_parts >= 0, "psx_dps_post_philox_dev: bad sizes");
  PSX_REQUIRE((err_parts > 0) == (d_err_part != nullptr),
              "psx_dps_post_philox_dev: d_err_part and err_parts must agree");
  return launch_post(d_x_t, d_eps, d_cot, d_vjp, nullptr, d_err_part, err_parts, L, n, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f,
                     d_step_row, d_x_next, d_err_out, 2, 0, 0, d_seed_step, (cudaStream_t)stream);
}

PSX_API int psx_dps_pre_bf16(const psx_op* op, const void* d_x_t, const void* d_eps, const float* d_y, int64_t L,
                             int64_t obs_repeat, float sqrt_acp, float sqrt_1m_acp, float lik_weight,
                             const float* d_step_row, void* d_cot, float* d_err_part, void* ws, size_t ws_bytes,
                             void* stream) {
  PSX_REQUIRE(op && d_x_t && d_eps && d_y && d_cot && d_err_part, "psx_dps_pre_bf16: null pointer");
  PSX_REQUIRE(L > 0 && L <= 65535 && obs_repeat > 0, "psx_dps_pre_bf16: bad sizes");
  PSX_REQUIRE(d_step_row || (sqrt_acp > 0.f && std::isfinite(sqrt_acp) && std::isfinite(sqrt_1m_acp) &&
                             std::isfinite(lik_weight)),
              "psx_dps_pre_bf16: non-finite or non-positive schedule scalar");
  const bool dev = d_step_row != nullptr;
  const float sa = dev ? 1.f : sqrt_acp, s1 = dev ? 0.f : sqrt_1m_acp, w = dev ? 1.f : lik_weight;
  if (op->kind == PSX_OP_BOX && op->d_keep)
    return fail(PSX_ERR_UNSUPPORTED, "psx_dps_pre_bf16: the masked box operator runs on an fp32 state only");
  if (op->kind == PSX_OP_BOX)
    return launch_pre_box_bf16(op, d_x_t, d_eps, d_y, L, obs_repeat, sa, s1, w, d_step_row, d_cot, d_err_part,
                               (cudaStream_t)stream);
  if (op->kind == PSX_OP_SEPBLUR) {
    if (int rc = check_ws(op, L, ws, ws_bytes)) return rc;
    return launch_pre_sepblur(op, (const float*)d_x_t, (const float*)d_eps, d_y, L, obs_repeat, sa, s1, w, d_step_row,
                              (float*)d_cot, d_err_part, nullptr, (float*)ws, (cudaStream_t)stream, true);
  }
  if (op->kind != PSX_OP_IDENTITY && op->kind != PSX_OP_MASK)
    return fail(PSX_ERR_UNSUPPORTED, "psx_dps_pre_bf16: the bf16 state path covers identity, mask, 4x box and separable-blur operators");
  return launch_pre_pointwise_bf16(op, d_x_t, d_eps, d_y, L, obs_repeat, sa, s1, w, d_step_row, d_cot, d_err_part,
                                   (cudaStream_t)stream);
}

PSX_API int psx_dps_post_bf16(const void* d_x_t, const void* d_eps, const void* d_cot, const void* d_vjp, const void* d_z,
                              const float* d_err_part, int err_parts, int64_t L, int64_t n, float sqrt_acp,
                              float sqrt_1m_acp, float c_ell, float c_s, float std_, float gamma,
                              const float* d_step_row, int use_philox, uint64_t seed, uint64_t step,
                              const uint64_t* d_seed_step, void* d_x_next, float* d_err_out, void* stream) {
  PSX_REQUIRE(d_x_t && d_eps && d_cot && d_vjp && d_x_next, "psx_dps_post_bf16: null pointer");
  PSX_REQUIRE(L > 0 && L <= 65535 && n > 0 && err_parts >= 0, "psx_dps_post_bf16: bad sizes");
  PSX_REQUIRE((err_parts > 0) == (d_err_part != nullptr), "psx_dps_post_bf16: d_err_part and err_parts must agree");
  const bool dev = d_step_row != nullptr;
  PSX_REQUIRE(dev || (sqrt_acp > 0.f && std::isfinite(sqrt_acp) && std::isfinite(c_ell) && std::isfinite(c_s) &&
                      std::isfinite(std_) && std::isfinite(gamma)),
              "psx_dps_post_bf16: non-finite scalar");
  PSX_REQUIRE(use_philox || d_z || (!dev && std_ == 0.f), "psx_dps_post_bf16: need d_z, Philox noise or std == 0");
  const int zmode = use_philox ? 2 : (d_z && (dev || std_ != 0.f) ? 1 : 0);
  return launch_post_bf16(d_x_t, d_eps, d_cot, d_vjp, d_z, d_err_part, err_parts, L, n, dev ? 1.f : sqrt_acp,
                          dev ? 0.f : sqrt_1m_acp, c_ell, c_s, std_, gamma, d_step_row, d_x_next, d_err_out, zmode, seed,
                          step, d_seed_step, (cudaStream_t)stream);
}

PSX_API int psx_philox_normal(float* d_out, int64_t numel, uint64_t seed, uint64_t step, void* stream) {
  PSX_REQUIRE(d_out && numel > 0, "psx_philox_normal: null pointer or empty tensor");
  return launch_philox_normal(d_out, numel, seed, step, (cudaStream_t)stream);
}

PSX_API int psx_bridge_update(const float* d_x, const float* d_eps, const float* d_z, const float* d_grad,
                              int64_t numel, float sqrt_acp, float sqrt_1m_acp, float c_ell, float c_s,
                              float std_, float grad_scale, float* d_x_next, void* stream) {
  PSX_REQUIRE(d_x && d_eps && d_x_next && numel > 0, "psx_bridge_update: null pointer or empty tensor");
  PSX_REQUIRE(d_z || std_ == 0.f, "psx_bridge_update: d_z may be NULL only when std == 0");
  PSX_REQUIRE(d_grad || grad_scale == 0.f, "psx_bridge_update: d_grad may be NULL only when grad_scale == 0");
  PSX_REQUIRE(sqrt_acp > 0.f && std::isfinite(sqrt_acp) && std::isfinite(c_ell) && std::isfinite(c_s) &&
                  std::isfinite(std_) && std::isfinite(grad_scale),
              "psx_bridge_update: non-finite scalar");
  return launch_bridge_update(d_x, d_eps, std_ == 0.f ? nullptr : d_z, grad_scale == 0.f ? nullptr : d_grad,
                              numel, sqrt_acp, sqrt_1m_acp, c_ell, c_s, std_, grad_scale, d_x_next,
                              (cudaStream_t)stream);
}

PSX_API int psx_lincomb3(const float* d_a, float ca, const float* d_b, float cb, const float* d_c, float cc,
                         float* d_out, int64_t numel, void* stream) {
  PSX_REQUIRE(d_a && d_b && d_out && numel > 0, "psx_lincomb3: null pointer or empty tensor");
  return launch_lincomb3(d_a, ca, d_b, cb, d_c, cc, d_out, numel, (cudaStream_t)stream);
}

PSX_API int psx_lincomb3_dev(const float* d_a, float ca, const float* d_b, float cb, const float* d_c, float cc,
                             const float* d_num, const float* d_den, float* d_out, int64_t numel, void* stream) {
  PSX_REQUIRE(d_a && d_b && d_c && d_num && d_out && numel > 0, "psx_lincomb3_dev: null pointer or empty tensor");
  return launch_lincomb3(d_a, ca, d_b, cb, d_c, cc, d_out, numel, (cudaStream_t)stream, d_num, d_den);
}

PSX_API int psx_ddim_eps_step(const float* d_x, const float* d_eps, const float* d_z, int64_t numel,
                              float sqrt_a_t, float sqrt_oma, float oma, float sqrt_a_p, float dir,
                              float sigma_t, float* d_x_prev, float* d_pred_x0, float* d_pseudo_x0,
                              void* stream) {
  PSX_REQUIRE(d_x && d_eps && d_x_prev && numel > 0, "psx_ddim_eps_step: null pointer or empty tensor");
  PSX_REQUIRE(d_z || sigma_t == 0.f, "psx_ddim_eps_step: d_z may be NULL only when sigma_t == 0");
  PSX_REQUIRE(sqrt_a_t > 0.f && std::isfinite(sqrt_a_t) && std::isfinite(sqrt_oma) && std::isfinite(oma) &&
                  std::isfinite(sqrt_a_p) && std::isfinite(dir) && std::isfinite(sigma_t),
              "psx_ddim_eps_step: non-finite scalar");
  return launch_ddim_eps(d_x, d_eps, sigma_t == 0.f ? nullptr : d_z, numel, sqrt_a_t, sqrt_oma, oma, sqrt_a_p, dir,
                         sigma_t, d_x_prev, d_pred_x0, d_pseudo_x0, (cudaStream_t)stream);
}

PSX_API int psx_stochastic_resample(const float* d_pseudo_x0, const float* d_x_t, const float* d_noise,
                                    int64_t numel, float c_p, float c_x, float den, float k_n, float* d_out,
                                    void* stream) {
  PSX_REQUIRE(d_pseudo_x0 && d_x_t && d_noise && d_out && numel > 0, "psx_stochastic_resample: null pointer");
  PSX_REQUIRE(den != 0.f && std::isfinite(den) && std::isfinite(c_p) && std::isfinite(c_x) && std::isfinite(k_n),
              "psx_stochastic_resample: bad scalar");
  return launch_stoch_resample(d_pseudo_x0, d_x_t, d_noise, numel, c_p, c_x, den, k_n, d_out, (cudaStream_t)stream);
}

PSX_API int psx_adamw_step(float* d_param, const float* d_grad, float* d_m, float* d_v, int64_t numel, float lr,
                           float beta1, float beta2, float adam_eps, float weight_decay, int step,
                           int* d_flags, int flag_in, const float* d_loss_parts, int64_t n_loss_parts,
                           float loss_scale, float loss_threshold, void* stream) {
  PSX_REQUIRE(d_param && d_grad && d_m && d_v && numel > 0, "psx_adamw_step: null pointer or empty tensor");
  PSX_REQUIRE(step >= 1 && lr > 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f,
              "psx_adamw_step: bad hyper-parameter");
  PSX_REQUIRE(!d_flags || ((flag_in == 0 || flag_in == 1) && d_loss_parts && n_loss_parts > 0),
              "psx_adamw_step: early-stop flags need loss partials");
  return launch_adamw(d_param, d_grad, d_m, d_v, numel, lr, beta1, beta2, adam_eps, weight_decay, step, d_flags,
                      flag_in, d_loss_parts, n_loss_parts, loss_scale, loss_threshold, (cudaStream_t)stream);
}

PSX_API int psx_tweedie(const float* d_x_t, const float* d_eps, int64_t L, int64_t n, float sqrt_acp,
                float sqrt_1m_acp, float* d_x0, float* d_sum, float* d_sumsq, void* stream) {
  PSX_REQUIRE(d_x_t && d_eps && d_x0, "psx_tweedie: null pointer");
  PSX_REQUIRE(L > 0 && n > 0, "psx_tweedie: bad sizes");
  PSX_REQUIRE(sqrt_acp > 0.f && std::isfinite(sqrt_acp), "psx_tweedie: bad sqrt_acp");
  return launch_tweedie(d_x_t, d_eps, L, n, sqrt_acp, sqrt_1m_acp, d_x0, d_sum, d_sumsq, (cudaStream_t)stream);
}

}  // extern "C"
