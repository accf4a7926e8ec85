// ldphase_probe.cu -- how fast can one CTA per SM pull its 256 x 176 fp32 windows of x_t and eps (the A1 operand of
// the tensor-core blur K1) out of HBM?  Variants of the access pattern, 96 / 148 CTAs of 512 threads, ~200 KB of
// dynamic shared memory per CTA as in blur_k1_tc, cold L2 before every launch.
//   0  two LDG.128 (L1::no_allocate) per tensor, lane = (row in 8-row group, 32-byte chunk), 4 items in flight
//   1  one LDG.256 per tensor, same mapping, 4 items in flight
//   2  as 0 but allocating in L1
//   3  one LDG.256 per tensor, all 12 items in flight
//   4  row-coalesced LDG.128 (lane = 16-byte piece of one row), linear shared-memory stores (wrong layout, timing only)
//   5  TMA: eight 2-D boxes of 176 x 32 fp32 per tensor, all of x in flight, then all of eps
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I samplers_b200/csrc -I include -o tools/micro/ldphase_probe tools/micro/ldphase_probe.cu -lcuda
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "psx_tc.cuh"

using namespace psx::tc;

constexpr int PAD = 24, KC = 22, LBO = 8192, LO = 4096, N = 256;

__device__ __forceinline__ float4 ldna(const float* p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldca(const float* p) {
  float4 v;
  asm("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void ld256(const float* p, float4& a, float4& b) {
  asm("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(p));
}
__device__ __forceinline__ void conv_store(uint8_t* op, int kc, int rg, int r, bool in, float s1, float4 xa, float4 xb,
                                           float4 ea, float4 eb) {
  uint4 hi, lo;
  split2(fmaf(-s1, ea.x, xa.x), fmaf(-s1, ea.y, xa.y), hi.x, lo.x);
  split2(fmaf(-s1, ea.z, xa.z), fmaf(-s1, ea.w, xa.w), hi.y, lo.y);
  split2(fmaf(-s1, eb.x, xb.x), fmaf(-s1, eb.y, xb.y), hi.z, lo.z);
  split2(fmaf(-s1, eb.z, xb.z), fmaf(-s1, eb.w, xb.w), hi.w, lo.w);
  if (!in) hi = lo = make_uint4(0, 0, 0, 0);
  if (kc < KC) {
    uint8_t* d = op + kc * LBO + rg * 128 + r * 16;
    *reinterpret_cast<uint4*>(d) = hi;
    *reinterpret_cast<uint4*>(d + LO) = lo;
  }
}

template <int V, int BATCH>
__device__ __forceinline__ void load_items(uint8_t* op, const float* xp, const float* ep, int j0, float s1, int warp,
                                           int lane) {
  const int r = lane & 7, c = lane >> 3;
  for (int b0 = 0; b0 < 12; b0 += BATCH) {
    float4 xa[BATCH], xb[BATCH], ea[BATCH], eb[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int it = warp + (b0 + u) * 16, rg = it / 6, kc = (it % 6) * 4 + c;
      const int j = j0 - PAD + 8 * kc;
      const bool in = kc < KC && j >= 0 && j < N;
      const int off = (8 * rg + r) * N + (in ? j : j0);
      if (V == 0) { xa[u] = ldna(xp + off); xb[u] = ldna(xp + off + 4); ea[u] = ldna(ep + off); eb[u] = ldna(ep + off + 4); }
      if (V == 2) { xa[u] = ldca(xp + off); xb[u] = ldca(xp + off + 4); ea[u] = ldca(ep + off); eb[u] = ldca(ep + off + 4); }
      if (V == 1 || V == 3) { ld256(xp + off, xa[u], xb[u]); ld256(ep + off, ea[u], eb[u]); }
    }
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int it = warp + (b0 + u) * 16, rg = it / 6, kc = (it % 6) * 4 + c;
      const int j = j0 - PAD + 8 * kc;
      conv_store(op, kc, rg, r, j >= 0 && j < N, s1, xa[u], xb[u], ea[u], eb[u]);
    }
  }
}

template <int V>
__global__ void __launch_bounds__(512, 1) ldphase(const float* __restrict__ x, const float* __restrict__ eps, float* out,
                                                  const __grid_constant__ CUtensorMap mx,
                                                  const __grid_constant__ CUtensorMap me, long long* tim) {
  extern __shared__ __align__(1024) uint8_t op[];
  __shared__ uint64_t bar[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int plane = blockIdx.x >> 1, rank = blockIdx.x & 1, j0 = rank * 128;
  const float* xp = x + (size_t)plane * N * N;
  const float* ep = eps + (size_t)plane * N * N;
  long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (V <= 3) {
    if (V == 3) load_items<3, 12>(op, xp, ep, j0, 0.6f, warp, lane);
    else load_items<V, 4>(op, xp, ep, j0, 0.6f, warp, lane);
  } else if (V == 4) {
    // 256 rows x 44 pieces of 16 B (columns j0 - 24 ..): 22 piece-iterations of 512 threads, two batches of 11
    for (int b0 = 0; b0 < 22; b0 += 11) {
      float4 a[11], e[11];
#pragma unroll
      for (int u = 0; u < 11; ++u) {
        const int p = tid + (b0 + u) * 512, row = p / 44, pc = p % 44, j = j0 - PAD + 4 * pc;
        const int off = row * N + ((j >= 0 && j < N) ? j : j0);
        a[u] = ldna(xp + off);
        e[u] = ldna(ep + off);
      }
#pragma unroll
      for (int u = 0; u < 11; ++u) {
        const int p = tid + (b0 + u) * 512;
        uint2 hi, lo;
        split2(fmaf(-0.6f, e[u].x, a[u].x), fmaf(-0.6f, e[u].y, a[u].y), hi.x, lo.x);
        split2(fmaf(-0.6f, e[u].z, a[u].z), fmaf(-0.6f, e[u].w, a[u].w), hi.y, lo.y);
        *reinterpret_cast<uint2*>(op + (size_t)p * 8) = hi;
        *reinterpret_cast<uint2*>(op + 90112 + (size_t)p * 8) = lo;
      }
    }
  } else {
    if (tid == 0) {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      fence_mbar_init();
    }
    __syncthreads();
    float acc = 0.f;
    for (int ten = 0; ten < 2; ++ten) {
      if (tid == 0) {
        mbar_expect_tx(&bar[ten], 8 * 176 * 32 * 4);
        for (int b = 0; b < 8; ++b)
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                  smem_u32(op + b * 176 * 32 * 4)),
              "l"(ten ? &me : &mx), "r"(j0 - PAD), "r"(plane * N + b * 32), "r"(smem_u32(&bar[ten]))
              : "memory");
      }
      mbar_wait(&bar[ten], 0);
      for (int i = tid; i < 176 * 256 / 4; i += 512) {
        const float4 v = reinterpret_cast<const float4*>(op)[i];
        acc += v.x + v.y + v.z + v.w;
      }
      __syncthreads();
    }
    if (acc == 1234.5f) out[tid] = acc;
  }
  __syncthreads();
  long long t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (tid == 0) tim[blockIdx.x] = t1 - t0;
  if (op[tid * 16] == 77 && op[tid * 16 + 90112] == 78) out[blockIdx.x * 512 + tid] = 1.f;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int V>
static void run(const float* x, const float* e, float* out, CUtensorMap mx, CUtensorMap me, long long* tim, float* flush,
                int grid) {
  cudaFuncSetAttribute(ldphase<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<float> ts;
  std::vector<long long> h(grid);
  double cta = 0;
  for (int it = 0; it < 6; ++it) {
    cudaMemsetAsync(flush, it, 256u << 20, 0);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    ldphase<V><<<grid, 512, 200 * 1024>>>(x, e, out, mx, me, tim);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (it) ts.push_back(ms * 1e3f);
    cudaMemcpy(h.data(), tim, grid * 8, cudaMemcpyDeviceToHost);
    cta = 0;
    for (auto v : h) cta += (double)v;
    cta /= grid;
  }
  std::sort(ts.begin(), ts.end());
  printf("variant %d grid %3d: launch %.2f us (median of 5, cold L2), per-CTA phase %.2f us avg  [%s]\n", V, grid,
         ts[ts.size() / 2], cta / 1e3, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int L = 32, C = 3;
  const size_t tot = (size_t)L * C * N * N;
  float *x, *e, *out, *flush;
  long long* tim;
  cudaMalloc(&x, tot * 4);
  cudaMalloc(&e, tot * 4);
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&tim, 8 * 1024);
  cudaMalloc(&flush, 256u << 20);
  cudaMemset(x, 0, tot * 4);
  cudaMemset(e, 0, tot * 4);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(p);
  CUtensorMap mx, me;
  const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)L * C * N};
  const cuuint64_t strides[1] = {(cuuint64_t)N * 4};
  const cuuint32_t box[2] = {176, 32};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r1 = fn(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = fn(&me, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, e, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("tensor maps: %d %d\n", (int)r1, (int)r2);
  for (int grid : {96, 148, 192}) {
    run<0>(x, e, out, mx, me, tim, flush, grid);
    run<1>(x, e, out, mx, me, tim, flush, grid);
    run<2>(x, e, out, mx, me, tim, flush, grid);
    run<3>(x, e, out, mx, me, tim, flush, grid);
    run<4>(x, e, out, mx, me, tim, flush, grid);
    run<5>(x, e, out, mx, me, tim, flush, grid);
  }
  return 0;
}
