// FP32 FMA throughput probe for sm_100a: plain FFMA (3-register), FFMA with a constant-bank
// operand, and packed FFMA2 (fma.rn.f32x2).  One CTA of 1024 threads per SM x 4 waves.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, const float* wsrc, int iters) {
  float a[16];
  float2 b[16];
  const float w0 = wsrc[threadIdx.x & 7], w1 = wsrc[(threadIdx.x + 1) & 7];
  for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 0.001f + i; b[i] = make_float2(a[i], a[i] + 1.f); }
  const float2 ww = make_float2(w0, w1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a[i] = fmaf(a[i], w0, w1);
      if (MODE == 1) a[i] = fmaf(a[i], 0.999f, 0.001f);
      if (MODE == 2) b[i] = __ffma2_rn(b[i], ww, ww);
    }
  }
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += a[i] + b[i].x + b[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* out, float* w, int fma_per_instr) {
  const int iters = 4096, blocks = 148 * 2;
  k<MODE><<<blocks, 1024>>>(out, w, 16);
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  cudaEventRecord(s);
  k<MODE><<<blocks, 1024>>>(out, w, iters);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  double fma = (double)blocks * 1024 * iters * 16 * fma_per_instr;
  printf("%-10s %.3f ms  %.2f TFMA/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", name, ms, fma / ms / 1e9,
         fma / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
  float *out, *w;
  cudaMalloc(&out, 148 * 2 * 1024 * 4); cudaMalloc(&w, 64);
  float hw[8] = {0.999f, 0.001f, 0.998f, 0.002f, 0.997f, 0.003f, 0.996f, 0.004f};
  cudaMemcpy(w, hw, 32, cudaMemcpyHostToDevice);
  run<0>("FFMA", out, w, 1);
  run<1>("FFMA.imm", out, w, 1);
  run<2>("FFMA2", out, w, 2);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
