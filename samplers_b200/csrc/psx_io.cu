// psx_io.cu -- the data formats either side of the sampling loop:
//   * observation simulation  y = A(x) + eps          (samplers/inverse_problem.py:35-67, noise.py:81-92,125-138)
//   * [-1, 1] float CHW  <->  uint8 HWC image tensors (samplers/utils/image.py:9-64 + torchvision's
//     to_pil_image / to_tensor arithmetic), so that only bytes cross PCIe.
// Every expression keeps torch's rounding sequence (one rounding per tensor op); results are bit-exact.
#include "psx_common.cuh"

namespace psx {

constexpr int kIoThreads = 256;

static int io_blocks(int64_t work) {
  int64_t b = (work + kIoThreads - 1) / kIoThreads;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

// y += RN(RN(a * noise) + b)      Gaussian: a = sigma, b = 0 (noise ~ N(0,1));  Poisson: a = 1, b = -rate (noise = k)
__global__ void __launch_bounds__(kIoThreads)
k_add_noise(float* __restrict__ y, const float* __restrict__ noise, int64_t total, float a, float b) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __fadd_rn(y[i], __fadd_rn(__fmul_rn(a, noise[i]), b));
}

int launch_add_noise(float* y, const float* noise, int64_t total, float a, float b, cudaStream_t st) {
  k_add_noise<<<io_blocks(total), kIoThreads, 0, st>>>(y, noise, total, a, b);
  return check_cuda(cudaGetLastError(), "k_add_noise launch");
}

// One thread per pixel: C coalesced plane reads, C adjacent byte writes.
//   u8 = (uint8) trunc( RN( RN(RN(clamp(x, -1, 1) + 1) * 0.5) * 255 ) )
// image.py:24-27 (clamp, +1, *0.5) then torchvision to_pil_image (mul(255).byte()).
__global__ void __launch_bounds__(kIoThreads)
k_image_to_u8(const float* __restrict__ chw, uint8_t* __restrict__ hwc, int64_t images, int C, int64_t HW) {
  const int64_t total = images * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t img = i / HW, p = i - img * HW;
    const float* src = chw + img * C * HW + p;
    uint8_t* dst = hwc + i * C;
    for (int c = 0; c < C; ++c) {
      float v = src[(int64_t)c * HW];
      v = fminf(fmaxf(v, -1.f), 1.f);
      v = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.f), 0.5f), 255.f);
      dst[c] = (uint8_t)(int)v;  // v in [0, 255]: C truncation, as Tensor.byte()
    }
  }
}

//   x = RN( RN(RN(u8 / 255) * 2) - 1 )        torchvision to_tensor (div(255)) then image.py:60 (*2 - 1)
__global__ void __launch_bounds__(kIoThreads)
k_image_from_u8(const uint8_t* __restrict__ hwc, float* __restrict__ chw, int64_t images, int C, int64_t HW) {
  const int64_t total = images * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t img = i / HW, p = i - img * HW;
    const uint8_t* src = hwc + i * C;
    float* dst = chw + img * C * HW + p;
    for (int c = 0; c < C; ++c) {
      const float f = __fdiv_rn((float)src[c], 255.f);
      dst[(int64_t)c * HW] = __fsub_rn(__fmul_rn(f, 2.f), 1.f);
    }
  }
}

int launch_image_to_u8(const float* chw, uint8_t* hwc, int64_t images, int C, int64_t HW, cudaStream_t st) {
  k_image_to_u8<<<io_blocks(images * HW), kIoThreads, 0, st>>>(chw, hwc, images, C, HW);
  return check_cuda(cudaGetLastError(), "k_image_to_u8 launch");
}

int launch_image_from_u8(const uint8_t* hwc, float* chw, int64_t images, int C, int64_t HW, cudaStream_t st) {
  k_image_from_u8<<<io_blocks(images * HW), kIoThreads, 0, st>>>(hwc, chw, images, C, HW);
  return check_cuda(cudaGetLastError(), "k_image_from_u8 launch");
}

}  // namespace psx
