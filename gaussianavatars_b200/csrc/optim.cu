// optim.cu -- Adam on the splat arrays as ONE multi-tensor launch.
//
// The reference steps `torch.optim.Adam(l, lr=0.0, eps=1e-15)` over six parameter groups with per-group learning rates
// (scene/gaussian_model.py:213-232, train.py:207-209): ~10 eager kernels per group per step in the default
// implementation.  The update is pure HBM streaming -- read p, g, m, v (16 B), write p, m, v (12 B) per element --
// so all groups go through one grid-stride kernel with 128-bit accesses; `blockIdx.y` picks the group.
//
// Arithmetic follows torch's single-tensor Adam (torch/optim/adam.py, amsgrad=False, weight_decay=0, maximize=False):
//   m <- m + (g - m) (1 - beta1)                 (Tensor.lerp_)
//   v <- beta2 v + (1 - beta2) g g               (mul_ / addcmul_)
//   p <- p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// The scalars (1 - beta, lr / (1 - beta1^t), ...) are formed on the host in double from double hyper-parameters, as
// torch forms them from Python floats, and rounded to float once: 1.0f - 0.999f would already be off by 1.3e-5.
#include "common.cuh"
#include "kernels.cuh"

namespace gab {

struct AdamBatch {
  float* p[GAB_ADAM_MAX_SEGMENTS];
  const float* g[GAB_ADAM_MAX_SEGMENTS];
  float* m[GAB_ADAM_MAX_SEGMENTS];
  float* v[GAB_ADAM_MAX_SEGMENTS];
  int64_t n[GAB_ADAM_MAX_SEGMENTS];
  float step_size[GAB_ADAM_MAX_SEGMENTS];  // lr / bias_correction1
  int32_t vec4[GAB_ADAM_MAX_SEGMENTS];     // bit 0: p, m, v 16-byte aligned (128-bit path); bit 1: g aligned too
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float w1, float beta2, float w2,
                                         float inv_bc2_sqrt, float eps, float step_size) {
  m = m + (g - m) * w1;
  v = v * beta2 + (w2 * g) * g;
  const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_kernel(AdamBatch b, float w1, float beta2, float w2, float inv_bc2_sqrt,
                                                   float eps) {
  const int s = blockIdx.y;
  const int64_t n = b.n[s];
  float* __restrict__ p = b.p[s];
  const float* __restrict__ g = b.g[s];
  float* __restrict__ m = b.m[s];
  float* __restrict__ v = b.v[s];
  const float step_size = b.step_size[s];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t tail = 0;
  if (b.vec4[s] & 1) {
    // the gradient is usually a view into the fused backward's flat buffer: its offset need not be 16-byte aligned
    const bool g_vec = (b.vec4[s] & 2) != 0;
    const int64_t n4 = n >> 2;
    auto load_g = [&](int64_t i) {
      return g_vec ? reinterpret_cast<const float4*>(g)[i] : make_float4(g[4 * i], g[4 * i + 1], g[4 * i + 2], g[4 * i + 3]);
    };
    auto update = [&](float4& P, const float4& G, float4& M, float4& V) {
      adam_one(P.x, G.x, M.x, V.x, w1, beta2, w2, inv_bc2_sqrt, eps, step_size);
      adam_one(P.y, G.y, M.y, V.y, w1, beta2, w2, inv_bc2_sqrt, eps, step_size);
      adam_one(P.z, G.z, M.z, V.z, w1, beta2, w2, inv_bc2_sqrt, eps, step_size);
      adam_one(P.w, G.w, M.w, V.w, w1, beta2, w2, inv_bc2_sqrt, eps, step_size);
    };
    int64_t i = t0;
    for (; i + stride < n4; i += 2 * stride) {  // two independent 4 x 16 B load groups in flight per thread
      const int64_t j = i + stride;
      float4 P0 = reinterpret_cast<float4*>(p)[i], P1 = reinterpret_cast<float4*>(p)[j];
      const float4 G0 = load_g(i), G1 = load_g(j);
      float4 M0 = reinterpret_cast<float4*>(m)[i], M1 = reinterpret_cast<float4*>(m)[j];
      float4 V0 = reinterpret_cast<float4*>(v)[i], V1 = reinterpret_cast<float4*>(v)[j];
      update(P0, G0, M0, V0);
      update(P1, G1, M1, V1);
      reinterpret_cast<float4*>(p)[i] = P0; reinterpret_cast<float4*>(p)[j] = P1;
      reinterpret_cast<float4*>(m)[i] = M0; reinterpret_cast<float4*>(m)[j] = M1;
      reinterpret_cast<float4*>(v)[i] = V0; reinterpret_cast<float4*>(v)[j] = V1;
    }
    if (i < n4) {
      float4 P = reinterpret_cast<float4*>(p)[i];
      const float4 G = load_g(i);
      float4 M = reinterpret_cast<float4*>(m)[i];
      float4 V = reinterpret_cast<float4*>(v)[i];
      update(P, G, M, V);
      reinterpret_cast<float4*>(p)[i] = P;
      reinterpret_cast<float4*>(m)[i] = M;
      reinterpret_cast<float4*>(v)[i] = V;
    }
    tail = n4 << 2;
  }
  for (int64_t i = tail + t0; i < n; i += stride) {
    float P = p[i], M = m[i], V = v[i];
    adam_one(P, g[i], M, V, w1, beta2, w2, inv_bc2_sqrt, eps, step_size);
    p[i] = P;
    m[i] = M;
    v[i] = V;
  }
}

void launch_adam(int num_segments, const gab200_adam_segment* segs, int64_t step, double beta1, double beta2, double eps,
                 cudaStream_t stream) {
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  int next = 0;  // first input segment not yet consumed (empty segments are skipped without filling a slot)
  while (next < num_segments) {
    AdamBatch b;
    int cnt = 0;
    int64_t longest = 0;
    for (; next < num_segments && cnt < GAB_ADAM_MAX_SEGMENTS; next++) {
      const gab200_adam_segment& s = segs[next];
      if (s.n <= 0) continue;
      b.p[cnt] = s.param;
      b.g[cnt] = s.grad;
      b.m[cnt] = s.exp_avg;
      b.v[cnt] = s.exp_avg_sq;
      b.n[cnt] = s.n;
      b.step_size[cnt] = (float)(s.lr / bc1);
      const uintptr_t bits = (uintptr_t)s.param | (uintptr_t)s.exp_avg | (uintptr_t)s.exp_avg_sq;
      b.vec4[cnt] = (bits & 15) == 0 ? (((uintptr_t)s.grad & 15) == 0 ? 3 : 1) : 0;
      longest = s.n > longest ? s.n : longest;
      cnt++;
    }
    if (cnt == 0) continue;
    for (int i = cnt; i < GAB_ADAM_MAX_SEGMENTS; i++) {
      b.p[i] = nullptr; b.g[i] = nullptr; b.m[i] = nullptr; b.v[i] = nullptr;
      b.n[i] = 0; b.step_size[i] = 0.f; b.vec4[i] = 0;
    }
    // two float4 per thread for the longest group, capped at 8 waves of 148 SMs x 8 resident CTAs
    int64_t blocks = (longest / 8 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8 * 8) blocks = 148 * 8 * 8;
    adam_kernel<<<dim3((unsigned)blocks, (unsigned)cnt), 256, 0, stream>>>(
        b, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), inv_bc2_sqrt, (float)eps);
    count_launch();
  }
}

}  // namespace gab
