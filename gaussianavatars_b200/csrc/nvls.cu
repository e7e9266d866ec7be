// nvls.cu -- all-reduce of the flat splat-gradient buffer through NVSwitch multicast (NVLS), two-shot, in place.
//
// Every rank holds a replica of the buffer in NVLink symmetric memory; `mc` is the MULTICAST address of that
// allocation (one address that names all replicas).  Rank r owns the r-th slice:
//     multimem.ld_reduce.add  [mc + i]   -- the switch fetches element i from every replica and returns the SUM
//     multimem.st             [mc + i]   -- the switch writes the sum into every replica
// so each element crosses NVLink once on the way in and once on the way out per GPU ((N-1)/N of the buffer each way),
// against ~2x that for a ring and N x for the push-style reduction (`multimem.red` from preprocess_bwd, dist.py).
// The caller brackets the launch with two group barriers (all replicas written / all slices stored): dist.py does it
// with the symmetric-memory signal pads, stream-ordered, no host involvement -- the whole thing is captured inside the
// step's CUDA graph.  Replaces one ncclAllReduce of 23.6 MB per step.
#include "common.cuh"
#include "kernels.cuh"

namespace gab {

__device__ __forceinline__ float4 mc_ld_reduce4(const float* mc_addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void mc_st4(float* mc_addr, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float mc_ld_reduce1(const float* mc_addr) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc_addr) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st1(float* mc_addr, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc_addr), "f"(v) : "memory");
}

// quads [q0, q1) of the buffer belong to this rank; the last rank also takes the n % 4 trailing floats
__global__ void __launch_bounds__(256) nvls_allreduce_kernel(float* __restrict__ mc, int64_t q0, int64_t q1,
                                                             int64_t tail0, int64_t tail1) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t q = q0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // four switch round trips in flight per thread: a multimem.ld_reduce takes a few microseconds to come back
  for (; q + 3 * stride < q1; q += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = mc_ld_reduce4(mc + 4 * (q + u * stride));
#pragma unroll
    for (int u = 0; u < 4; ++u) mc_st4(mc + 4 * (q + u * stride), v[u]);
  }
  for (; q < q1; q += stride) {
    float* a = mc + 4 * q;
    mc_st4(a, mc_ld_reduce4(a));
  }
  if (blockIdx.x == 0)
    for (int64_t i = tail0 + threadIdx.x; i < tail1; i += blockDim.x) mc_st1(mc + i, mc_ld_reduce1(mc + i));
  __threadfence_system();   // the stores are performed at every replica before the closing group barrier signals
}

void launch_nvls_allreduce(float* mc, int64_t n, int rank, int world, cudaStream_t stream) {
  if (n <= 0 || world <= 0) return;
  const int64_t quads = n / 4;
  const int64_t per = (quads + world - 1) / world;
  const int64_t q0 = per * rank < quads ? per * rank : quads;
  const int64_t q1 = q0 + per < quads ? q0 + per : quads;
  const bool last = rank == world - 1;
  const int64_t tail0 = last ? quads * 4 : 0, tail1 = last ? n : 0;
  int64_t blocks = (q1 - q0 + 255) / 256;
  if (blocks < 1) blocks = 1;
  const int knob = tune_get(GAB200_TUNE_NVLS_CTAS);
  // measured on 2 B200s (profiles/r02/nvls_probe_n2.json): 32 .. 1184 CTAs all take 70-80 us for 23.6 MB -- the switch,
  // not the issue rate, is the limit -- so the kernel stays small and leaves the SMs to the frame it runs beside
  const int64_t cap = knob > 0 ? knob : 64;
  if (blocks > cap) blocks = cap;
  nvls_allreduce_kernel<<<(unsigned)blocks, 256, 0, stream>>>(mc, q0, q1, tail0, tail1);
  count_launch();
}

}  // namespace gab
