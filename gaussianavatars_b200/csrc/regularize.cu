// regularize.cu -- the two splat regularisers of the mesh-bound training step, forward and backward in one launch each.
//
// Replaces train.py:134-146 (with arguments/__init__.py:100-105: lambda_xyz 1e-2, threshold_xyz 1, lambda_scale 1,
// threshold_scale 0.6, metric_* False):
//     losses['xyz']   = relu(_xyz[vis].norm(dim=1) - t_xyz).mean() * l_xyz                       (metric_xyz False)
//                       relu((_xyz * face_scaling[binding])[vis] - t_xyz).norm(dim=1).mean() * l_xyz   (True)
//     losses['scale'] = relu(exp(_scaling[vis]) - t_s).norm(dim=1).mean() * l_s                    (metric_scale False)
//                       relu(get_scaling[vis] - t_s).norm(dim=1).mean() * l_s                      (True)
// with vis = radii > 0 of the frame just rendered.  In the reference these lines call get_scaling / index the
// parameters with a boolean mask every step (a dozen eager launches, several (P,3) temporaries, and the only reason a
// training step still needs the eager getters once render() is fused).  Here: one pass over 28 B/splat each way.
#include "common.cuh"
#include "kernels.cuh"

namespace gab {

struct RegTerm {
  float l;        // ||relu(v - t)||   (or relu(||x|| - t) for the non-metric position term)
  float d[3];     // d l / d v_k
};
__device__ __forceinline__ RegTerm relu_norm(const float v[3], float t) {
  RegTerm r;
  float s[3], q = 0.f;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    s[k] = fmaxf(v[k] - t, 0.f);
    q += s[k] * s[k];
  }
  r.l = sqrtf(q);
#pragma unroll
  for (int k = 0; k < 3; k++) r.d[k] = r.l > 0.f ? s[k] / r.l : 0.f;  // torch's norm backward: 0 at the origin
  return r;
}

// sums[0] = sum of the position terms, [1] = sum of the scale terms, [2] = number of visible splats
template <bool BACKWARD>
__global__ void __launch_bounds__(256) regularize_kernel(gab200_regularize_args a, double* __restrict__ sums,
                                                         const float* __restrict__ g_out /* [2] upstream */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float tx = 0.f, ts = 0.f, cnt = 0.f;
  float cx = 0.f, cs = 0.f;
  if (BACKWARD) {
    const double n = sums[2];
    cx = (float)((double)g_out[0] * (double)a.lambda_xyz / n);  // mean over the visible splats; n = 0 -> nan like torch
    cs = (float)((double)g_out[1] * (double)a.lambda_scale / n);
  }
  if (i < a.P) {
    const bool vis = a.radii[i] > 0;
    const bool bound = a.binding != nullptr;
    const int f = bound ? a.binding[i] : 0;
    const float fs = bound ? a.face_scaling[f] : 1.f;
    float x[3], e[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      x[k] = a.xyz[3 * (size_t)i + k];
      e[k] = expf(a.scaling[3 * (size_t)i + k]);
    }
    float gx[3] = {0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f}, gfs = 0.f;
    if (vis) {
      cnt = 1.f;
      if (a.metric_xyz) {
        const float v[3] = {x[0] * fs, x[1] * fs, x[2] * fs};
        const RegTerm r = relu_norm(v, a.threshold_xyz);
        tx = r.l;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          gx[k] = cx * r.d[k] * fs;
          gfs += cx * r.d[k] * x[k];
        }
      } else {
        const float nrm = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        tx = fmaxf(nrm - a.threshold_xyz, 0.f);
        const float w = (nrm > a.threshold_xyz && nrm > 0.f) ? cx / nrm : 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) gx[k] = w * x[k];
      }
      if (a.lambda_scale != 0.f) {
        const float m = a.metric_scale ? fs : 1.f;
        const float v[3] = {e[0] * m, e[1] * m, e[2] * m};
        const RegTerm r = relu_norm(v, a.threshold_scale);
        ts = r.l;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          gs[k] = cs * r.d[k] * v[k];                      // d exp(s) m / d s = exp(s) m
          if (a.metric_scale) gfs += cs * r.d[k] * e[k];
        }
      }
    }
    if (BACKWARD) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        a.grad_xyz[3 * (size_t)i + k] = gx[k];
        a.grad_scaling[3 * (size_t)i + k] = gs[k];
      }
      if (bound && a.grad_face_scaling != nullptr && gfs != 0.f) atomicAdd(a.grad_face_scaling + f, gfs);
    }
  }
  if (!BACKWARD) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      tx += __shfl_xor_sync(0xffffffffu, tx, m);
      ts += __shfl_xor_sync(0xffffffffu, ts, m);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, m);
    }
    __shared__ float part[3][8];
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { part[0][w] = tx; part[1][w] = ts; part[2][w] = cnt; }
    __syncthreads();
    if (threadIdx.x < 3) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 8; k++) s += (double)part[threadIdx.x][k];
      if (s != 0.0) atomicAdd(sums + threadIdx.x, s);
    }
  }
}

__global__ void regularize_finish_kernel(gab200_regularize_args a, const double* __restrict__ sums) {
  const double n = sums[2];
  a.loss[0] = (float)(sums[0] / n * (double)a.lambda_xyz);
  a.loss[1] = a.lambda_scale != 0.f ? (float)(sums[1] / n * (double)a.lambda_scale) : 0.f;
  a.loss[2] = (float)n;
}

cudaError_t launch_regularize_forward(const gab200_regularize_args& a, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(a.sums, 0, 3 * sizeof(double), stream);
  if (e != cudaSuccess) return e;
  if (a.P > 0) {
    regularize_kernel<false><<<(a.P + 255) / 256, 256, 0, stream>>>(a, a.sums, nullptr);
    count_launch();
  }
  regularize_finish_kernel<<<1, 1, 0, stream>>>(a, a.sums);
  count_launch();
  return cudaSuccess;
}
cudaError_t launch_regularize_backward(const gab200_regularize_args& a, const float* g_out, cudaStream_t stream) {
  if (a.P == 0) return cudaSuccess;
  regularize_kernel<true><<<(a.P + 255) / 256, 256, 0, stream>>>(a, a.sums, g_out);
  count_launch();
  return cudaSuccess;
}

}  // namespace gab
