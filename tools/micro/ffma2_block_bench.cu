// Isolated throughput of the FFMA2 conv blocks used by the blur kernels (sm_100a).
//   A: row_block16<40> as in the kernels (tap pairs via uniform registers, window via LDS.128)
//   B: FFMA2 with a UR tap operand, window held in registers (no shared-memory traffic)
//   C: FFMA2 with all three operand pairs in vector registers
// Reports FMA/clk/SM assuming the kernel runs at the clock it reports through clock64().
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

struct Taps { int k; int lo; float2 ww[136]; };

template <int K>
__device__ __forceinline__ void row_block16(float2 (&acc)[16], const float2* __restrict__ row, const Taps& taps) {
  float2 win[24];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const float4 v = *reinterpret_cast<const float4*>(row + 2 * m);
    win[2 * m] = make_float2(v.x, v.y); win[2 * m + 1] = make_float2(v.z, v.w);
  }
#pragma unroll
  for (int c = 0; c < K; c += 8) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const float4 v = *reinterpret_cast<const float4*>(row + c + 16 + 2 * m);
      win[16 + 2 * m] = make_float2(v.x, v.y); win[17 + 2 * m] = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 w = taps.ww[c + i];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = __ffma2_rn(w, win[i + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) win[j] = win[j + 8];
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, int iters, int pitch2, const __grid_constant__ Taps taps,
                                            long long* cycles) {
  extern __shared__ __align__(128) float2 comp[];
  for (int i = threadIdx.x; i < 16 * pitch2; i += 256) comp[i] = make_float2(i * 1e-4f, 1.f);
  __syncthreads();
  const int rp = threadIdx.x & 15, cg = threadIdx.x >> 4;
  float2 acc[16];
  for (int j = 0; j < 16; ++j) acc[j] = make_float2(0.f, 0.f);
  const long long t0 = clock64();
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) row_block16<40>(acc, comp + rp * pitch2 + 16 * cg, taps);
  } else if (MODE == 1) {
    float2 win[24];
    for (int j = 0; j < 24; ++j) win[j] = comp[rp * pitch2 + j];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 40; c += 8)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 w = taps.ww[c + i];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = __ffma2_rn(w, win[i + j], acc[j]);
        }
    }
  } else {
    float2 win[24], wv[8];
    for (int j = 0; j < 24; ++j) win[j] = comp[rp * pitch2 + j];
    for (int j = 0; j < 8; ++j) wv[j] = comp[rp * pitch2 + 30 + j];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 40; c += 8)
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = __ffma2_rn(wv[i], win[i + j], acc[j]);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < 16; ++j) s += acc[j].x + acc[j].y;
  out[blockIdx.x * 256 + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ctas_per_sm) {
  const int pitch2 = 306, iters = 200, blocks = 148 * ctas_per_sm;
  float* out; long long* cyc;
  cudaMalloc(&out, blocks * 256 * 4); cudaMalloc(&cyc, blocks * 8);
  Taps t; t.k = 40; t.lo = -20;
  for (int i = 0; i < 136; ++i) t.ww[i] = make_float2(1e-3f * i, 1e-3f * i);
  const size_t smem = 16 * pitch2 * 8;
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  bench<MODE><<<blocks, 256, smem>>>(out, 2, pitch2, t, cyc);
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  cudaEventRecord(s);
  bench<MODE><<<blocks, 256, smem>>>(out, iters, pitch2, t, cyc);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  long long h[148 * 8]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
  const double fma_per_cta = 256.0 * iters * 640 * 2;  // 640 FFMA2 per block call, 2 FMA each
  printf("%-34s %d CTA/SM: %.3f ms, %.0f cycles/CTA -> %.1f FMA/clk/SM (event: %.2f TFMA/s)  %s\n", name, ctas_per_sm, ms,
         avg, fma_per_cta * ctas_per_sm / avg, fma_per_cta * blocks / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int c : {1, 2, 3}) run<0>("A row_block16 (UR taps + LDS.128)", c);
  for (int c : {1, 2}) run<1>("B UR taps, window in registers", c);
  for (int c : {1, 2}) run<2>("C all-register operands", c);
  return 0;
}
