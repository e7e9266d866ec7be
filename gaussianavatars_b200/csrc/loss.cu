// loss.cu -- L1 photometric loss against a uint8 ground-truth image, forward and gradient in ONE pass over the
// rendered image (the first half of train.py:128-131: `gt_image.cuda()`, `l1_loss(image, gt)`).
// The reference uploads the ground truth as float32 (4 B/px/channel) and runs ~6 eager kernels for L1 + its autograd;
// here the uint8 image is uploaded (1 B), converted in-register, and dL/dimage = sign(render - gt) / n is written in the
// same kernel that accumulates the loss.
#include "common.cuh"
#include "kernels.cuh"

namespace gab {

__global__ void __launch_bounds__(256) l1_loss_u8_kernel(int64_t n, const float* __restrict__ img,
                                                         const uint8_t* __restrict__ gt, float inv_n,
                                                         float* __restrict__ grad, float* __restrict__ loss_sum) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  float acc = 0.f;
  if (i4 + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(img + i4);
    const uchar4 g = *reinterpret_cast<const uchar4*>(gt + i4);
    const float d0 = v.x - g.x * (1.f / 255.f), d1 = v.y - g.y * (1.f / 255.f);
    const float d2 = v.z - g.z * (1.f / 255.f), d3 = v.w - g.w * (1.f / 255.f);
    acc = fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
    auto sgn = [](float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); };
    *reinterpret_cast<float4*>(grad + i4) = make_float4(sgn(d0) * inv_n, sgn(d1) * inv_n, sgn(d2) * inv_n, sgn(d3) * inv_n);
  } else {
    for (int64_t i = i4; i < n; i++) {
      const float d = img[i] - gt[i] * (1.f / 255.f);
      acc += fabsf(d);
      grad[i] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * inv_n;
    }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) s += part[w];
    atomicAdd(loss_sum, s * inv_n);
  }
}

void launch_l1_loss_u8(int64_t n, const float* img, const uint8_t* gt, float* grad, float* loss, cudaStream_t stream) {
  if (n == 0) return;
  const int64_t threads = (n + 3) / 4;
  l1_loss_u8_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(n, img, gt, 1.0f / (float)n, grad, loss);
  count_launch();
}

}  // namespace gab
