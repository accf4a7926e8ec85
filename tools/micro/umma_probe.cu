// umma_probe.cu -- hardware probe for the tcgen05 building blocks of the tensor-core blur K1 (psx_tcblur.cu).
//
//   umma_probe layout <a_mn> <opt> <swap>   exactness of one M=128 x N=64 x K=112 product read from a K window inside
//                                           a larger SWIZZLE_NONE operand (a_mn: A is MN-major; opt: which of the two
//                                           core-matrix orders; swap: exchange the LBO / SBO descriptor fields)
//   umma_probe split                        fp16 hi/lo three-product scheme vs fp64 on a banded Toeplitz B
//   umma_probe time                         tensor-pipe cycles of one blur pass for N-tile widths 16 .. 256
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I samplers_b200/csrc -o tools/micro/umma_probe tools/micro/umma_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "psx_tc.cuh"

using namespace psx::tc;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

constexpr int M = 128;
constexpr int KE = 176;  // K extent of the A operand in shared memory
constexpr int KW = 112;  // K window consumed by the product (7 k-steps)
constexpr int K0 = 64;   // window start inside the A operand
constexpr int N = 64;

struct Cfg {
  int a_mn, opt, swap, terms;
};

// byte offset of element (mn, k) of an operand with `rows` M/N rows and `kext` K columns
__device__ __host__ inline uint32_t op_off(int mn, int k, int rows, int kext, int mn_major, int opt, uint32_t* lbo,
                                           uint32_t* sbo) {
  const uint32_t SBO = opt == 0 ? 128u : (uint32_t)(kext / 8) * 128u;
  const uint32_t LBO = opt == 0 ? (uint32_t)(rows / 8) * 128u : 128u;
  if (lbo) *lbo = LBO;
  if (sbo) *sbo = SBO;
  const uint32_t core = (uint32_t)(mn / 8) * SBO + (uint32_t)(k / 8) * LBO;
  return core + (mn_major ? (uint32_t)(k % 8) * 16u + (uint32_t)(mn % 8) * 2u
                          : (uint32_t)(mn % 8) * 16u + (uint32_t)(k % 8) * 2u);
}

__global__ void __launch_bounds__(128) probe_mma(const float* __restrict__ A, const float* __restrict__ B, float* D,
                                                 Cfg cfg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + M * KE * 2;
  uint8_t* b_hi = a_lo + M * KE * 2;
  uint8_t* b_lo = b_hi + N * KW * 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<64>(&tslot);
  uint32_t lboA, sboA, lboB, sboB;
  op_off(0, 0, M, KE, cfg.a_mn, cfg.opt, &lboA, &sboA);
  op_off(0, 0, N, KW, 0, cfg.opt, &lboB, &sboB);
  for (int i = tid; i < M * KE; i += 128) {
    const int m = i / KE, k = i % KE;
    const float v = A[i];
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const uint32_t o = op_off(m, k, M, KE, cfg.a_mn, cfg.opt, nullptr, nullptr);
    *reinterpret_cast<__half*>(a_hi + o) = h;
    *reinterpret_cast<__half*>(a_lo + o) = l;
  }
  for (int i = tid; i < N * KW; i += 128) {
    const int n = i / KW, k = i % KW;
    const float v = B[i];
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const uint32_t o = op_off(n, k, N, KW, 0, cfg.opt, nullptr, nullptr);
    *reinterpret_cast<__half*>(b_hi + o) = h;
    *reinterpret_cast<__half*>(b_lo + o) = l;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  if (tid == 0) {
    const uint32_t idesc = idesc_f16(M, N, cfg.a_mn, 0);
    const uint32_t fA1 = cfg.swap ? sboA : lboA, fA2 = cfg.swap ? lboA : sboA;
    const uint32_t fB1 = cfg.swap ? sboB : lboB, fB2 = cfg.swap ? lboB : sboB;
    const uint64_t dAh = smem_desc(smem_u32(a_hi) + (K0 / 8) * lboA, fA1, fA2);
    const uint64_t dAl = smem_desc(smem_u32(a_lo) + (K0 / 8) * lboA, fA1, fA2);
    const uint64_t dBh = smem_desc(smem_u32(b_hi), fB1, fB2);
    const uint64_t dBl = smem_desc(smem_u32(b_lo), fB1, fB2);
    for (int ks = 0; ks < KW / 16; ++ks) {
      const uint32_t adv_a = 2u * lboA * ks, adv_b = 2u * lboB * ks;
      umma_f16(tbase, desc_advance(dAh, adv_a), desc_advance(dBh, adv_b), idesc, ks > 0);
      if (cfg.terms == 3) {
        umma_f16(tbase, desc_advance(dAl, adv_a), desc_advance(dBh, adv_b), idesc, 1);
        umma_f16(tbase, desc_advance(dAh, adv_a), desc_advance(dBl, adv_b), idesc, 1);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + (tid & 31)) * N + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tbase);
}

// ------------------------------------------------------------------------------------------ timing
// One blur pass = (256 / NT) N-tiles x ksteps(NT) x 3 products of M = 128; operands are zeros.
template <int NT>
__global__ void __launch_bounds__(128) probe_time(long long* out, int passes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(&tslot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  constexpr int KS = (NT + 24 + 19 + 15) / 16;
  if (tid == 0) {
    const uint32_t idesc = idesc_f16(128, NT, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 160 * 1024;
    const long long t0 = clock64();
    for (int p = 0; p < passes; ++p)
      for (int t = 0; t < 256 / NT; ++t)
        for (int ks = 0; ks < KS; ++ks) {
          const uint64_t da = smem_desc(a0 + (t * (NT / 8) + 2 * ks) * 2048, 2048, 128);
          const uint64_t db = smem_desc(b0 + 2 * ks * (NT / 8) * 128, (NT / 8) * 128, 128);
          umma_f16(tbase + t * NT, da, db, idesc, ks > 0);
          umma_f16(tbase + t * NT, da + (80 * 1024 >> 4), db, idesc, 1);
          umma_f16(tbase + t * NT, da, db + (14 * 1024 >> 4), idesc, 1);
        }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tbase);
}

template <int NT>
static int run_time() {
  long long* d;
  CK(cudaMalloc(&d, 16));
  CK(cudaFuncSetAttribute(probe_time<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int passes : {1, 4}) {
    probe_time<NT><<<1, 128, 200 * 1024>>>(d, passes);
    CK(cudaDeviceSynchronize());
    long long h[2];
    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("time NT=%d ksteps=%d passes=%d issue_cycles=%lld total_cycles=%lld per_pass=%lld\n", NT,
           (NT + 24 + 19 + 15) / 16, passes, h[0], h[1], h[1] / passes);
  }
  // all SMs at once: does the per-SM rate hold chip-wide?
  probe_time<NT><<<148, 128, 200 * 1024>>>(d, 4);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe_time<NT><<<148, 128, 200 * 1024>>>(d, 16);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("time NT=%d 148 CTAs x 16 passes: %.2f us total, %.3f us per pass\n", NT, ms * 1e3, ms * 1e3 / 16);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    printf("usage: umma_probe layout <a_mn> <opt> <swap> | split | time\n");
    return 1;
  }
  if (!strcmp(argv[1], "time")) {
    int rc = run_time<16>();
    rc |= run_time<32>();
    rc |= run_time<64>();
    rc |= run_time<128>();
    rc |= run_time<256>();
    return rc;
  }
  Cfg cfg{0, 0, 0, 1};
  const bool split = !strcmp(argv[1], "split");
  if (!split) {
    if (argc < 5) return 1;
    cfg.a_mn = atoi(argv[2]);
    cfg.opt = atoi(argv[3]);
    cfg.swap = atoi(argv[4]);
  } else {
    cfg.terms = 3;
    if (argc >= 3) cfg.a_mn = atoi(argv[2]);
  }
  std::vector<float> A(M * KE), B(N * KW);
  std::vector<double> ref(M * N, 0.0);
  srand(1);
  if (!split) {
    for (auto& v : A) v = (float)((rand() % 9) - 4) * 0.25f;
    for (auto& v : B) v = (float)((rand() % 9) - 4) * 0.25f;
  } else {
    for (auto& v : A) {
      float u1 = (rand() + 1.0f) / (RAND_MAX + 2.0f), u2 = rand() / (float)RAND_MAX;
      v = sqrtf(-2.f * logf(u1)) * cosf(6.2831853f * u2);
    }
    double w[39], s = 0;
    for (int i = 0; i < 39; ++i) s += (w[i] = exp(-0.5 * (i - 19) * (i - 19) / 9.0));
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < KW; ++k) {
        const int t = (k - 24) - n + 19;
        B[n * KW + k] = (t >= 0 && t < 39) ? (float)(w[t] / s) : 0.f;
      }
  }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < KW; ++k) acc += (double)A[m * KE + K0 + k] * (double)B[n * KW + k];
      ref[m * N + n] = acc;
    }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, M * N * 4));
  const size_t smem = 2 * (M * KE * 2) + 2 * (N * KW * 2);
  CK(cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_mma<<<1, 128, smem>>>(dA, dB, dD, cfg);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(M * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0, fro_e = 0, fro_r = 0;
  int bad = 0;
  for (int i = 0; i < M * N; ++i) {
    const double e = fabs((double)D[i] - ref[i]);
    if (!(e <= 1e-3 * (1 + fabs(ref[i])))) ++bad;
    if (!(e <= max_err)) max_err = e;
    max_ref = fmax(max_ref, fabs(ref[i]));
    fro_e += e * e;
    fro_r += ref[i] * ref[i];
  }
  printf("%s a_mn=%d opt=%d swap=%d terms=%d: max_abs_err=%.3e max_ref=%.3e rel_fro=%.3e mismatches=%d/%d  %s\n", argv[1],
         cfg.a_mn, cfg.opt, cfg.swap, cfg.terms, max_err, max_ref, sqrt(fro_e / fro_r), bad, M * N,
         bad == 0 ? "PASS" : "FAIL");
  printf("  D[0][0..3] = %g %g %g %g   ref = %g %g %g %g\n", D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
  return 0;
}
