// %globaltimer phase timeline of the tensor-core blur K1 (blur_k1_tc, psx_tcblur.cu) at config 2 (L = 16) or L = 64.
// Build (see README.md): -DPSX_TRACE -rdc=true together with the library sources.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "psx.h"
namespace psx { extern __device__ long long psx_trace_tc[1024 * 2 * 16]; }
int main(int argc, char** argv) {
  const int C = 3, H = 256, W = 256, L = argc > 1 ? atoi(argv[1]) : 16, k = 61;
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
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  std::vector<long long> h(1024 * 2 * 16);
  cudaMemcpyFromSymbol(h.data(), psx::psx_trace_tc, h.size() * 8);
  const int nb = std::min(1024, L * C * 2);
  long long first = h[0], last = 0;
  for (int b = 0; b < nb; ++b) { first = std::min(first, h[(b * 2) * 16]); last = std::max(last, h[(b * 2) * 16 + 13]); }
  printf("blur_k1_tc L=%d ctas %d (traced %d): span %.2f us\n", L, L * C * 2, nb, (last - first) / 1e3);
  const char* en[14] = {"start", "setup done", "A1 loaded", "P1 done seen", "E1 done", "P2 done seen", "E2 done", "P3 done seen",
                        "cluster wait 1", "E3 done", "cluster wait 2", "P4 done seen", "E4 done", "exit"};
  const char* mn[12] = {"start", "setup done", "B block landed", "A1 full seen", "P1 issued", "P1 done", "A2 full seen", "P2 issued",
                        "A3 full seen", "P3 issued", "A4 full + cluster 2", "P4 issued"};
  for (int role = 0; role < 2; ++role) {
    const int ns = role ? 12 : 14;
    printf("%s: us since CTA start, avg (max) over CTAs; start skew avg/max below\n", role ? "UMMA warp" : "epilogue warp 0");
    for (int sl = 0; sl < ns; ++sl) {
      double avg = 0, mx = 0;
      for (int b = 0; b < nb; ++b) { double d = (double)(h[(b * 2 + role) * 16 + sl] - h[(b * 2 + role) * 16]); avg += d; mx = std::max(mx, d); }
      printf("   %-22s %7.2f (%7.2f)\n", role ? mn[sl] : en[sl], avg / nb / 1e3, mx / 1e3);
    }
  }
  double sk = 0, skm = 0;
  for (int b = 0; b < nb; ++b) { double d = (double)(h[b * 2 * 16] - first); sk += d; skm = std::max(skm, d); }
  printf("CTA start skew avg %.2f max %.2f us\n", sk / nb / 1e3, skm / 1e3);
  return 0;
}
