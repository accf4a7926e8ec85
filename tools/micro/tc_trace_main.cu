// %globaltimer phase timeline of the tensor-core blur K1 (blur_k1_tc, psx_tcblur.cu) at config 2 (L = 16) or L = 64.
// Build (see README.md): -DPSX_TRACE -rdc=true together with the library sources.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "psx.h"
namespace psx { extern __device__ long long psx_trace_tc[1024 * 4 * 32]; }
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
    if (argc <= 2) cudaMemsetAsync(flush, it, 256u << 20, 0);   // x / eps come from HBM, as after a UNet pass (any 2nd argument: warm L2)
    if (psx_dps_pre(op, x, e, y, L, L, 0.8f, 0.6f, 400.f, cot, part, nullptr, ws, psx_op_workspace_bytes(op, L), 0)) { printf("%s\n", psx_last_error()); return 1; }
  }
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  std::vector<long long> h(1024 * 4 * 32);
  cudaMemcpyFromSymbol(h.data(), psx::psx_trace_tc, h.size() * 8);
  const int nb = std::min(1024, L * C * 2);
  long long first = h[0], last = 0;
  for (int b = 0; b < nb; ++b) { first = std::min(first, h[(b * 4) * 32]); last = std::max(last, h[(b * 4) * 32 + 27]); }
  printf("blur_k1_tc L=%d ctas %d (traced %d): span %.2f us\n", L, L * C * 2, nb, (last - first) / 1e3);
  const char* en[28] = {"start", "setup done", "A1 loaded (all chunks)", "E1(0) P1 tile 0 seen", "E1(0) done", "E1(1) P1 done seen",
                        "E1(1) done", "E2(0,0) P2 seen", "E2(0,0) done", "E2(0,1) P2 seen", "E2(0,1) done", "E2(1,0) P2 seen",
                        "E2(1,0) done", "E2(1,1) P2 seen", "E2(1,1) done", "E3(0) P3 seen", "E3(0) done", "E3(1) P3 seen",
                        "E3(1) done", "E4(0) P4 seen", "E4(0) done", "E4(1) P4 seen", "E4(1) done", "E4(2) P4 seen",
                        "E4(2) done", "E4(3) P4 seen", "E4(3) done", "exit"};
  const char* mn[19] = {"start", "setup done", "P1 issued", "A2(0) + halo seen", "P2(0) issued", "A2(1) + halo seen", "P2(1) issued",
                        "A3(0) + halo seen", "P3(0) issued", "A3(1) + halo seen", "P3(1) issued", "A4(0) seen", "P4 steps 1-7 issued",
                        "A4(1) seen", "P4 issued", "A2(0) local seen", "A2(1) local seen", "A3(0) local seen", "A3(1) local seen"};
  printf("SM clock during CTA 0: %.0f MHz\n", (double)(h[31] - h[30]) / ((double)(h[27] - h[0]) / 1e3));
  const char* cn[9] = {"start", "A2(0) halo copy issued", "A2(1) halo copy issued", "A3(0) halo copy issued", "A3(1) halo copy issued",
                       "A2(0) source ready", "A2(1) source ready", "A3(0) source ready", "A3(1) source ready"};
  const char* sn[11] = {"start", "slot free: A2(0)", "slot free: A2(1)", "slot free: A3(0)", "slot free: A3(1)", "", "", "",
                        "E2(1,0) warp 0: TMEM read", "E2(1,0) warp 0: residual + split done", "E2(1,0) warp 0: stored"};
  const char* rn[4] = {"epilogue warp 0", "UMMA thread", "halo-copy thread", "slot-free signal thread"};
  for (int role = 0; role < 4; ++role) {
    const int ns = role == 0 ? 28 : role == 1 ? 19 : role == 2 ? 9 : 11;
    printf("%s: us since CTA start, avg over the rank-0 CTAs | rank-1 CTAs (max over all)\n", rn[role]);
    for (int sl = 0; sl < ns; ++sl) {
      double avg[2] = {0, 0}, mx = 0;
      for (int b = 0; b < nb; ++b) { double d = (double)(h[(b * 4 + role) * 32 + sl] - h[(b * 4) * 32]); avg[b & 1] += d; mx = std::max(mx, d); }
      const char* nm = role == 0 ? en[sl] : role == 1 ? mn[sl] : role == 2 ? cn[sl] : sn[sl];
      if (role >= 2 && (sl == 0 || (sl > 4 && sl < 8))) continue;
      printf("   %-24s %7.2f | %7.2f (%7.2f)\n", nm, avg[0] / (nb / 2) / 1e3, avg[1] / (nb / 2) / 1e3, mx / 1e3);
    }
  }
  double sk = 0, skm = 0;
  for (int b = 0; b < nb; ++b) { double d = (double)(h[b * 4 * 32] - first); sk += d; skm = std::max(skm, d); }
  printf("CTA start skew avg %.2f max %.2f us\n", sk / nb / 1e3, skm / 1e3);
  return 0;
}
