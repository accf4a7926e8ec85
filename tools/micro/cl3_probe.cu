// cl3_probe.cu -- two questions behind a THREE-CTA cluster per plane for blur_k1_tc at L = 16 (48 planes):
//  (A) do 48 clusters of 3 CTAs (608 threads, ~190 KB of dynamic shared memory: one CTA per SM) run as ONE wave on this
//      part?  cudaOccupancyMaxActiveClusters for cluster sizes 2 / 3 / 4, then an actual launch of 48 x 3 CTAs that
//      records %smid and the start time of every CTA.
//  (B) how long does the A1 load phase (x_t / eps -> registers -> fp16 hi / lo -> shared memory, LDG.256, four units in
//      flight per thread as in blur_k1_tc) take when U units of 32 rows x 128 columns are spread over G CTAs:
//      G = 96, U = 8 (today: a cluster pair per plane) against G = 144, U = 5 / 6 (a third CTA takes a third of the units).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I samplers_b200/csrc -I include -o tools/micro/cl3_probe tools/micro/cl3_probe.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "psx_tc.cuh"

using namespace psx::tc;

constexpr int N = 256;

__device__ __forceinline__ void ld256(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}

// ------------------------------------------------------------------------------------------------ (A)
__global__ void __launch_bounds__(608, 1) resident(long long* t_start, int* smid, int spin_us) {
  extern __shared__ uint8_t sm[];
  long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (threadIdx.x == 0) {
    int s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    t_start[blockIdx.x] = t0;
    smid[blockIdx.x] = s;
    sm[0] = 1;
  }
  cluster_arrive_release();
  cluster_wait_acquire();
  long long t;
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < 1000LL * spin_us);
  cluster_arrive_release();
  cluster_wait_acquire();
}

static void probe_resident(int csize, int nclusters) {
  const int smem = 190 * 1024;
  cudaFuncSetAttribute(resident, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(resident, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * csize);
  cfg.blockDim = dim3(608);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csize;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int maxc = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, resident, &cfg);
  printf("cluster size %d: cudaOccupancyMaxActiveClusters = %d (%s)\n", csize, maxc, cudaGetErrorString(e));
  long long* ts;
  int* sm;
  const int g = nclusters * csize;
  cudaMalloc(&ts, g * 8);
  cudaMalloc(&sm, g * 4);
  for (int rep = 0; rep < 3; ++rep) {
    e = cudaLaunchKernelEx(&cfg, resident, ts, sm, 30);
    cudaError_t e2 = cudaDeviceSynchronize();
    std::vector<long long> h(g);
    std::vector<int> s(g);
    cudaMemcpy(h.data(), ts, g * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(s.data(), sm, g * 4, cudaMemcpyDeviceToHost);
    long long mn = *std::min_element(h.begin(), h.end()), mx = *std::max_element(h.begin(), h.end());
    std::vector<int> u(s);
    std::sort(u.begin(), u.end());
    const int distinct = (int)(std::unique(u.begin(), u.end()) - u.begin());
    int late = 0;
    for (auto v : h) late += (v - mn > 10000);
    printf("  launch of %d clusters x %d: %s / %s, start spread %.2f us, distinct SMs %d, CTAs starting > 10 us late: %d\n",
           nclusters, csize, cudaGetErrorString(e), cudaGetErrorString(e2), (mx - mn) / 1e3, distinct, late);
  }
  cudaFree(ts);
  cudaFree(sm);
}

// ------------------------------------------------------------------------------------------------ (B)
// unit w = (plane, half, chunk): rows 32 chunk .. + 31, columns 128 half .. + 127 of x and eps
template <int U>
__global__ void __launch_bounds__(608, 1) ldunits(const float* __restrict__ x, const float* __restrict__ eps, float* out,
                                                  long long* tim, int units_total) {
  extern __shared__ __align__(1024) uint8_t op[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (warp < 16) {
    const int r = lane & 7, cg = 4 * (warp & 3) + (lane >> 3), kb = warp >> 2;
    float4 xa[4], xb[4], ea[4], eb[4];
    auto addr = [&](int i) -> size_t {
      int w = (blockIdx.x * U + i) % units_total;
      const int plane = w >> 4, half = (w >> 3) & 1, chunk = w & 7;
      return (size_t)plane * N * N + (size_t)(32 * chunk + 8 * kb + r) * N + 128 * half + 8 * cg;
    };
#pragma unroll
    for (int u = 0; u < 4 && u < U; ++u) {
      ld256(x + addr(u), xa[u], xb[u]);
      ld256(eps + addr(u), ea[u], eb[u]);
    }
    uint8_t* d0 = op + kb * 4096 + cg * 128 + r * 16;
    const float s1 = 0.6f;
#pragma unroll
    for (int c = 0; c < U; ++c) {
      const int u = c & 3;
      uint4 hi, lo;
      split2(fmaf(-s1, ea[u].x, xa[u].x), fmaf(-s1, ea[u].y, xa[u].y), hi.x, lo.x);
      split2(fmaf(-s1, ea[u].z, xa[u].z), fmaf(-s1, ea[u].w, xa[u].w), hi.y, lo.y);
      split2(fmaf(-s1, eb[u].x, xb[u].x), fmaf(-s1, eb[u].y, xb[u].y), hi.z, lo.z);
      split2(fmaf(-s1, eb[u].z, xb[u].z), fmaf(-s1, eb[u].w, xb[u].w), hi.w, lo.w);
      if (c + 4 < U) {
        ld256(x + addr(c + 4), xa[u], xb[u]);
        ld256(eps + addr(c + 4), ea[u], eb[u]);
      }
      *reinterpret_cast<uint4*>(d0 + c * 16384) = hi;
      *reinterpret_cast<uint4*>(d0 + c * 16384 + 2048) = lo;
    }
  }
  __syncthreads();
  long long t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (tid == 0) tim[blockIdx.x] = t1 - t0;
  if (op[tid * 16] == 77 && op[tid * 16 + 2048] == 78) out[blockIdx.x * 608 + tid] = 1.f;
}

template <int U>
static void run_units(const float* x, const float* e, float* out, long long* tim, float* flush, int grid, int units_total,
                      bool warm) {
  const int smem = U * 16384 + 1024;
  cudaFuncSetAttribute(ldunits<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<float> ts;
  std::vector<long long> h(grid);
  double cta = 0, ctamax = 0;
  for (int it = 0; it < 6; ++it) {
    if (!warm) cudaMemsetAsync(flush, it, 512u << 20, 0);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    ldunits<U><<<grid, 608, 200 * 1024>>>(x, e, out, tim, units_total);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (it) ts.push_back(ms * 1e3f);
    cudaMemcpy(h.data(), tim, grid * 8, cudaMemcpyDeviceToHost);
    cta = 0;
    ctamax = 0;
    for (auto v : h) {
      cta += (double)v;
      ctamax = std::max(ctamax, (double)v);
    }
    cta /= grid;
  }
  (void)smem;
  std::sort(ts.begin(), ts.end());
  printf("grid %3d x %d units (%3d KB in / CTA, %5.1f MB total) %s L2: launch %.2f us, per-CTA %.2f us avg / %.2f max -> %.1f GB/s per SM  [%s]\n",
         grid, U, U * 32, grid * U * 32 / 1024.0, warm ? "warm" : "cold", ts[ts.size() / 2], cta / 1e3, ctamax / 1e3,
         U * 32768.0 / cta, cudaGetErrorString(cudaGetLastError()));
}


// ------------------------------------------------------------------------------------------------ (C)
// hybrid: units 0..3 of the CTA through LDG (as above), units 4..7 concurrently by TMA (one 2-D box of 128 columns x 32
// rows per tensor and unit) into a 128 KB staging area; then every warp reads the staged data once (LDS.128 pass).
// Do the LDG path and the TMA path have separate per-SM limits (does the sum exceed the ~32 GB/s of LDG alone)?
#include <cuda.h>
template <int NL, int NT>
__global__ void __launch_bounds__(608, 1) ldhybrid(const float* __restrict__ x, const float* __restrict__ eps, float* out,
                                                   const __grid_constant__ CUtensorMap mx,
                                                   const __grid_constant__ CUtensorMap me, long long* tim) {
  extern __shared__ __align__(1024) uint8_t op[];
  __shared__ uint64_t bar;
  uint8_t* stage = op + 65536;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long t0, t1 = 0, t2 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int plane = blockIdx.x >> 1, half = blockIdx.x & 1;
  if (warp == 16 && lane == 0 && NT > 0) {
    mbar_expect_tx(&bar, NT * 2 * 16384);
    for (int u = 0; u < NT; ++u) {
      const int chunk = NL + u;
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                       smem_u32(stage + u * 32768)),
                   "l"(&mx), "r"(128 * half), "r"(plane * N + 32 * chunk), "r"(smem_u32(&bar))
                   : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                       smem_u32(stage + u * 32768 + 16384)),
                   "l"(&me), "r"(128 * half), "r"(plane * N + 32 * chunk), "r"(smem_u32(&bar))
                   : "memory");
    }
  }
  float acc = 0.f;
  if (warp < 16) {
    const int r = lane & 7, cg = 4 * (warp & 3) + (lane >> 3), kb = warp >> 2;
    float4 xa[4], xb[4], ea[4], eb[4];
    auto addr = [&](int i) -> size_t {
      return (size_t)plane * N * N + (size_t)(32 * i + 8 * kb + r) * N + 128 * half + 8 * cg;
    };
#pragma unroll
    for (int u = 0; u < 4 && u < NL; ++u) {
      ld256(x + addr(u), xa[u], xb[u]);
      ld256(eps + addr(u), ea[u], eb[u]);
    }
    uint8_t* d0 = op + kb * 4096 + cg * 128 + r * 16;
    const float s1 = 0.6f;
#pragma unroll
    for (int c = 0; c < NL; ++c) {
      const int u = c & 3;
      uint4 hi, lo;
      split2(fmaf(-s1, ea[u].x, xa[u].x), fmaf(-s1, ea[u].y, xa[u].y), hi.x, lo.x);
      split2(fmaf(-s1, ea[u].z, xa[u].z), fmaf(-s1, ea[u].w, xa[u].w), hi.y, lo.y);
      split2(fmaf(-s1, eb[u].x, xb[u].x), fmaf(-s1, eb[u].y, xb[u].y), hi.z, lo.z);
      split2(fmaf(-s1, eb[u].z, xb[u].z), fmaf(-s1, eb[u].w, xb[u].w), hi.w, lo.w);
      if (c + 4 < NL) {
        ld256(x + addr(c + 4), xa[u], xb[u]);
        ld256(eps + addr(c + 4), ea[u], eb[u]);
      }
      *reinterpret_cast<uint4*>(d0 + (c & 3) * 16384) = hi;
      *reinterpret_cast<uint4*>(d0 + (c & 3) * 16384 + 2048) = lo;
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (NT > 0) {
      mbar_wait(&bar, 0);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
      for (int i = tid; i < NT * 32768 / 16; i += 512) {
        const float4 v = reinterpret_cast<const float4*>(stage)[i];
        acc += v.x + v.y + v.z + v.w;
      }
    }
  }
  __syncthreads();
  long long t3;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t3));
  if (tid == 0) {
    tim[blockIdx.x * 4] = t3 - t0;
    tim[blockIdx.x * 4 + 1] = t1 - t0;
    tim[blockIdx.x * 4 + 2] = t2 - t0;
  }
  if (acc == 1234.5f || (op[tid * 16] == 77 && op[tid * 16 + 2048] == 78)) out[blockIdx.x * 608 + tid] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int NL, int NT>
static void run_hybrid(const float* x, const float* e, float* out, long long* tim, float* flush, int grid, CUtensorMap mx,
                       CUtensorMap me, bool warm) {
  cudaFuncSetAttribute(ldhybrid<NL, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<float> ts;
  std::vector<long long> h(grid * 4);
  double a[3] = {0, 0, 0};
  for (int it = 0; it < 6; ++it) {
    if (!warm) cudaMemsetAsync(flush, it, 512u << 20, 0);
    cudaEvent_t ea, eb;
    cudaEventCreate(&ea);
    cudaEventCreate(&eb);
    cudaEventRecord(ea);
    ldhybrid<NL, NT><<<grid, 608, 200 * 1024>>>(x, e, out, mx, me, tim);
    cudaEventRecord(eb);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, ea, eb);
    if (it) ts.push_back(ms * 1e3f);
    cudaMemcpy(h.data(), tim, grid * 32, cudaMemcpyDeviceToHost);
    for (int k = 0; k < 3; ++k) {
      a[k] = 0;
      for (int b = 0; b < grid; ++b) a[k] += (double)h[b * 4 + k];
      a[k] /= grid;
    }
  }
  std::sort(ts.begin(), ts.end());
  printf("hybrid grid %3d: %d units LDG + %d units TMA, %s L2: launch %.2f us, per-CTA total %.2f us (LDG part done %.2f, TMA landed %.2f) -> %.1f GB/s per SM  [%s]\n",
         grid, NL, NT, warm ? "warm" : "cold", ts[ts.size() / 2], a[0] / 1e3, a[1] / 1e3, a[2] / 1e3,
         (NL + NT) * 32768.0 / a[0], cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char**) {
  if (argc < 2) {
    for (int cs : {2, 3, 4}) probe_resident(cs, 48);
    probe_resident(3, 49);
    probe_resident(2, 74);
  }

  const int L = 16, C = 3;
  const size_t tot = (size_t)L * C * N * N;
  const int units_total = L * C * 16;
  float *x, *e, *out, *flush;
  long long* tim;
  cudaMalloc(&x, tot * 4);
  cudaMalloc(&e, tot * 4);
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&tim, 64 * 1024);
  cudaMalloc(&flush, 512u << 20);
  cudaMemset(x, 0, tot * 4);
  cudaMemset(e, 0, tot * 4);
  void* pfn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &pfn, cudaEnableDefault, &qr);
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(pfn);
  CUtensorMap mx, me;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)L * C * N};
    const cuuint64_t strides[1] = {(cuuint64_t)N * 4};
    const cuuint32_t box[2] = {128, 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r1 = fn(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = fn(&me, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, e, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("tensor maps: %d %d\n", (int)r1, (int)r2);
  }
  for (int warm = 0; warm < 2; ++warm) {
    run_hybrid<8, 0>(x, e, out, tim, flush, 96, mx, me, warm);
    run_hybrid<4, 4>(x, e, out, tim, flush, 96, mx, me, warm);
    run_hybrid<5, 3>(x, e, out, tim, flush, 96, mx, me, warm);
    run_hybrid<3, 4>(x, e, out, tim, flush, 96, mx, me, warm);
    run_hybrid<0, 4>(x, e, out, tim, flush, 96, mx, me, warm);
    run_hybrid<4, 0>(x, e, out, tim, flush, 96, mx, me, warm);
    run_units<8>(x, e, out, tim, flush, 96, units_total, warm);
    run_units<8>(x, e, out, tim, flush, 48, units_total, warm);
    run_units<5>(x, e, out, tim, flush, 144, units_total, warm);
    run_units<6>(x, e, out, tim, flush, 144, units_total, warm);
    run_units<6>(x, e, out, tim, flush, 128, units_total, warm);
    run_units<5>(x, e, out, tim, flush, 148, units_total, warm);
    run_units<4>(x, e, out, tim, flush, 96, units_total, warm);
    run_units<2>(x, e, out, tim, flush, 96, units_total, warm);
  }
  return 0;
}
