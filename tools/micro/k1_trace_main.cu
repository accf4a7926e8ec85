// %globaltimer phase timeline of the three default blur-K1 kernels (rows<Tweedie> -> cols16 -> rows_il) at config 2.
// Build (see README.md): -DPSX_TRACE -rdc=true together with the library sources.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "psx.h"
namespace psx { extern __device__ long long psx_trace3[3 * 1024 * 8]; }
int main() {
  const int C = 3, H = 256, W = 256, L = 16, k = 61;
  std::vector<float> taps(k);
  double s = 0; for (int i = 0; i < k; ++i) { double d = i - 30; taps[i] = (float)exp(-d * d / 18.0); s += taps[i]; }
  for (auto& t : taps) t = (float)(t / s);
  psx_op* op; if (psx_op_create_sepblur(C, H, W, taps.data(), k, taps.data(), k, &op)) { printf("%s\n", psx_last_error()); return 1; }
  const size_t n = (size_t)C * H * W, tot = n * L;
  float *x, *e, *y, *cot, *part, *ws, *flush;
  cudaMalloc(&x, tot * 4); cudaMalloc(&e, tot * 4); cudaMalloc(&y, n * 4); cudaMalloc(&cot, tot * 4);
  cudaMalloc(&part, L * psx_op_err_parts(op) * 4); cudaMalloc(&ws, psx_op_workspace_bytes(op, L));
  cudaMalloc(&flush, 256u << 20);
  cudaMemset(x, 0, tot * 4); cudaMemset(e, 0, tot * 4); cudaMemset(y, 0, n * 4);
  for (int it = 0; it < 4; ++it) {
    cudaMemsetAsync(flush, it, 256u << 20, 0);   // x / eps come from HBM, as after a UNet pass
    if (psx_dps_pre(op, x, e, y, L, L, 0.8f, 0.6f, 400.f, cot, part, nullptr, ws, psx_op_workspace_bytes(op, L), 0)) { printf("%s\n", psx_last_error()); return 1; }
  }
  cudaDeviceSynchronize();
  std::vector<long long> h(3 * 1024 * 8);
  cudaMemcpyFromSymbol(h.data(), psx::psx_trace3, h.size() * 8);
  const char* kn[3] = {"rows<Tweedie>", "cols16", "rows_il"};
  long long g0 = 0;
  for (int kid = 0; kid < 3; ++kid) {
    int nb = 0; while (nb < 1024 && h[(kid * 1024 + nb) * 8] != 0) ++nb;
    if (!nb) { printf("%s: no trace\n", kn[kid]); continue; }
    long long first = h[kid * 1024 * 8], last = 0;
    for (int b = 0; b < nb; ++b) { first = std::min(first, h[(kid * 1024 + b) * 8]); last = std::max(last, h[(kid * 1024 + b) * 8 + 5]); }
    if (kid == 0) g0 = first;
    double avg[6] = {0}, mx[6] = {0}; double start_avg = 0, start_max = 0;
    for (int b = 0; b < nb; ++b) {
      const long long* r = &h[(kid * 1024 + b) * 8];
      start_avg += (double)(r[0] - first); start_max = std::max(start_max, (double)(r[0] - first));
      for (int i = 1; i < 6; ++i) { double d = (double)(r[i] - r[0]); avg[i] += d; mx[i] = std::max(mx[i], d); }
    }
    printf("%-14s ctas %3d  span %.2f us (first CTA at %+.2f us of K1)  cta start skew avg %.2f max %.2f us\n", kn[kid], nb,
           (last - first) / 1e3, (first - g0) / 1e3, start_avg / nb / 1e3, start_max / 1e3);
    printf("   us since CTA start, avg (max): prologue %.2f (%.2f) | first data %.2f (%.2f) | tile0 mid %.2f (%.2f) | tile0 end %.2f (%.2f) | exit %.2f (%.2f)\n",
           avg[1] / nb / 1e3, mx[1] / 1e3, avg[2] / nb / 1e3, mx[2] / 1e3, avg[3] / nb / 1e3, mx[3] / 1e3, avg[4] / nb / 1e3, mx[4] / 1e3,
           avg[5] / nb / 1e3, mx[5] / 1e3);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
