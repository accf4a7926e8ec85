// Phase timeline of blur_k1_fused (compiled with -DPSX_TRACE -rdc=true together with the library sources).
#include <algorithm>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "psx.h"
namespace psx { extern __device__ long long psx_trace[4096 * 16]; }
int main() {
  const int C = 3, H = 256, W = 256, L = 16, k = 61;
  std::vector<float> taps(k);
  double s = 0; for (int i = 0; i < k; ++i) { double d = i - 30; taps[i] = (float)exp(-d * d / 18.0); s += taps[i]; }
  for (auto& t : taps) t = (float)(t / s);
  psx_op* op; if (psx_op_create_sepblur(C, H, W, taps.data(), k, taps.data(), k, &op)) { printf("%s\n", psx_last_error()); return 1; }
  const size_t n = (size_t)C * H * W, tot = n * L;
  float *x, *e, *y, *cot, *part, *ws;
  cudaMalloc(&x, tot * 4); cudaMalloc(&e, tot * 4); cudaMalloc(&y, n * 4); cudaMalloc(&cot, tot * 4);
  cudaMalloc(&part, L * psx_op_err_parts(op) * 4); cudaMalloc(&ws, psx_op_workspace_bytes(op, L));
  cudaMemset(x, 0, tot * 4); cudaMemset(e, 0, tot * 4); cudaMemset(y, 0, n * 4);
  for (int it = 0; it < 3; ++it)
    if (psx_dps_pre(op, x, e, y, L, L, 0.8f, 0.6f, 400.f, cot, part, nullptr, ws, psx_op_workspace_bytes(op, L), 0)) { printf("%s\n", psx_last_error()); return 1; }
  cudaDeviceSynchronize();
  const int nb = 48 * 8;
  std::vector<long long> h(nb * 16);
  cudaMemcpyFromSymbol(h.data(), psx::psx_trace, nb * 16 * 8);
  long long t0 = h[0]; for (int b = 0; b < nb; ++b) t0 = std::min(t0, h[b * 16]);
  const char* names[14] = {"start", "zeroed+tma", "sync#0", "conv done", "H done", "sync#1", "V done", "sync#2", "r written", "sync#3", "Vt done", "Ht done", "err", "exit"};
  // slots: 0 start,1 before sync0,2 after sync0 (P2 start),3 conv done,4 H done(before sync1),5 after sync1,6 V done(before sync2),7 after sync2,8 r written(before sync3),9 after sync3... print raw deltas
  double avg[14] = {0};
  for (int b = 0; b < nb; ++b) for (int i = 0; i < 14; ++i) avg[i] += (double)(h[b * 16 + i] - h[b * 16]);
  printf("avg ns since CTA start: "); for (int i = 0; i < 14; ++i) printf("[%d]%.0f ", i, avg[i] / nb); printf("\n");
  long long first = t0, last = 0; for (int b = 0; b < nb; ++b) last = std::max(last, h[b * 16 + 13]);
  printf("kernel span %.1f us\n", (last - first) / 1e3);
  for (int b : {0, 7, 100, 263, 264, 300, 383}) { printf("cta %3d start %+7.1f us: ", b, (h[b * 16] - t0) / 1e3); for (int i = 1; i < 14; ++i) printf("%.1f ", (h[b * 16 + i] - h[b * 16]) / 1e3); printf("\n"); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
