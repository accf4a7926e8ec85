// dsmem_probe.cu -- latency of the cluster / distributed-shared-memory primitives a CTA pair can use to hand a 12 KB
// halo to its neighbour (psx_tcblur.cu).  Rank 0 sends, rank 1 polls a local mbarrier; both read %globaltimer.
//   0  mbarrier.arrive.release.cluster on the peer's barrier (no data)
//   1  mbarrier.arrive.relaxed.cluster (no data)
//   2  cp.async.bulk shared::cta -> shared::cluster, BYTES, complete_tx on the peer's barrier
//   3  4 warps x st.shared::cluster.v4 (BYTES in total), fence.proxy.async, one release arrive per warp
//   4  4 warps x st.shared::cluster.v4, then barrier.cluster.arrive.release / wait.acquire by every thread
//   5  4 warps x st.async (16 B, complete_tx on the peer's barrier)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I samplers_b200/csrc -o tools/micro/dsmem_probe tools/micro/dsmem_probe.cu
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "psx_tc.cuh"

using namespace psx::tc;

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int V>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1) probe(long long* out, int bytes, int busy) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);          // [0] data barrier
  uint8_t* src = sm + 1024;
  uint8_t* dst = sm + 1024 + 32768;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;
  if (tid == 0) {
    mbar_init(bar, V == 3 ? 4 : 1);
    fence_mbar_init();
    if (V == 2 || V == 5) mbar_expect_tx(bar, bytes);
  }
  for (int i = tid; i < 32768 / 16; i += blockDim.x) reinterpret_cast<uint4*>(src)[i] = make_uint4(i, i, i, i);
  fence_async_smem();
  __syncthreads();
  cluster_arrive_release();
  cluster_wait_acquire();
  // optional smem traffic on the receiving SM (warps 0-3 of rank 1 hammer shared memory while waiting)
  long long t0 = 0, t1 = 0;
  if (rank == 0) {
    if (warp < 4) {
      // let the receiver reach its polling loop
      const long long s = gtime();
      while (gtime() - s < 3000) {
      }
      __syncwarp();
      t0 = gtime();
      const uint32_t cbar = mapa(smem_u32(bar), peer), cdst = mapa(smem_u32(dst), peer);
      if (V == 0) {
        if (tid == 0) mbar_arrive_remote(cbar);
      } else if (V == 1) {
        if (tid == 0) asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cbar) : "memory");
      } else if (V == 2) {
        if (tid == 0) bulk_s2peer(cdst, src, bytes, cbar);
      } else if (V == 3 || V == 4) {
        for (int i = tid; i < bytes / 16; i += 128) st_cluster_v4(cdst + i * 16, make_uint4(i, 1, 2, 3));
        if (V == 3) {
          fence_async_all();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(cbar);
        }
      } else if (V == 5) {
        for (int i = tid; i < bytes / 16; i += 128)
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(
                           cdst + i * 16),
                       "r"(i), "r"(1), "r"(2), "r"(3), "r"(cbar)
                       : "memory");
      }
      t1 = gtime();
    }
    if (V == 4) {
      cluster_arrive_release();
      cluster_wait_acquire();
    }
    if (tid == 0) {
      out[blockIdx.x * 4 + 0] = t0;
      out[blockIdx.x * 4 + 1] = t1;
    }
  } else {
    if (V == 4) {
      cluster_arrive_release();
      cluster_wait_acquire();
      t1 = gtime();
    } else if (warp == 4) {
      if (lane == 0) {
        mbar_spin(bar, 0);
        t1 = gtime();
      }
    } else if (busy) {
      // shared-memory traffic on the receiving SM
      uint4 acc = make_uint4(0, 0, 0, 0);
      for (int it = 0; it < 400; ++it) {
        const uint4 v = reinterpret_cast<uint4*>(src)[(tid + it * 128) & 2047];
        acc.x += v.x;
        reinterpret_cast<uint4*>(src)[(tid + it * 128 + 64) & 2047] = acc;
      }
      if (acc.x == 12345) out[0] = 1;
    }
    if ((V == 4 && tid == 0) || (V != 4 && tid == 128)) out[blockIdx.x * 4 + 1] = t1;
  }
  cluster_arrive_release();
  cluster_wait_acquire();
}

template <int V>
static void run(long long* d_out, int bytes, int clusters, int busy) {
  const int smem = 1024 + 65536;
  cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<double> lat, issue;
  std::vector<long long> h(clusters * 8);
  for (int it = 0; it < 6; ++it) {
    probe<V><<<clusters * 2, 160, smem>>>(d_out, bytes, busy);
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), d_out, h.size() * 8, cudaMemcpyDeviceToHost);
    if (!it) continue;
    for (int c = 0; c < clusters; ++c) {
      lat.push_back((double)(h[(2 * c + 1) * 4 + 1] - h[(2 * c) * 4 + 0]));
      issue.push_back((double)(h[(2 * c) * 4 + 1] - h[(2 * c) * 4 + 0]));
    }
  }
  std::sort(lat.begin(), lat.end());
  std::sort(issue.begin(), issue.end());
  printf("variant %d bytes %5d clusters %3d busy %d: send -> seen by the peer %6.0f ns median (%6.0f max), sender busy %6.0f ns  [%s]\n",
         V, bytes, clusters, busy, lat[lat.size() / 2], lat.back(), issue[issue.size() / 2],
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8 * 4 * 512);
  for (int busy : {0, 1})
    for (int clusters : {1, 48}) {
      run<0>(d_out, 0, clusters, busy);
      run<1>(d_out, 0, clusters, busy);
      for (int bytes : {4096, 12288, 24576}) run<2>(d_out, bytes, clusters, busy);
      for (int bytes : {4096, 12288, 24576}) run<3>(d_out, bytes, clusters, busy);
      for (int bytes : {4096, 12288, 24576}) run<4>(d_out, bytes, clusters, busy);
      for (int bytes : {4096, 12288, 24576}) run<5>(d_out, bytes, clusters, busy);
    }
  return 0;
}
