// blend.cu -- per-tile front-to-back alpha blend (forward) and reverse-walk gradient (backward) for sm_100a.
// Replaces renderCUDA fwd/bwd of the reference module.
//
// B200-first mapping (NOT the reference's 1 thread = 1 pixel, 256-thread block):
//   * one CTA per 16x16 tile with 256/K threads; every thread owns a COLUMN STRIP of K pixels.  The x-offset to a
//     splat (dx) is then shared by the K pixels, so the exponent is 3 flops per pixel,
//         power(dy) = p0 + dy * (q + h * dy),   p0 = -A dx^2/2, q = -B dx, h = -C/2   (pre-scaled by log2 e -> ex2),
//     and the shared-memory broadcast reads of the splat record are amortised K times.
//   * splat records (48 B, three 16-B quads) are GATHERED straight into shared memory with cp.async (LDGSTS),
//     double buffered one chunk ahead, ids one further chunk ahead: no register staging, no exposed L2 latency.
//   * backward: per-lane partial sums over the K pixels collapse to three moments (S0,S1,S2) because dx is
//     shared; nine per-splat gradient components are then reduced across the warp with a 14-shuffle
//     multi-value butterfly (instead of 45 shuffles or 9*32 atomics) and leave the SM as ONE 9-lane RED.ADD.F32
//     per (warp, splat); warps whose pixels all rejected the splat skip the reduction entirely (vote).
// Tensor cores are not used: there is no dense contraction on this path (north_star).
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace gab {

#define LOG2E 1.4426950408889634f
#define ALPHA_MIN (1.0f / 255.0f)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void gather_rec(SplatRec* dst, const SplatRec* src) {
  cp_async16(&dst->q0, &src->q0);
  cp_async16(&dst->q1, &src->q1);
  cp_async16(&dst->q2, &src->q2);
}

// =====================================================================================================
// Forward
// =====================================================================================================
template <int K>
__global__ void __launch_bounds__(256 / K) blend_forward_kernel(int W, int H, int gx, const uint2* __restrict__ ranges,
                                                                const uint32_t* __restrict__ point_list,
                                                                const SplatRec* __restrict__ rec,
                                                                const float* __restrict__ bg,
                                                                float* __restrict__ out_color,
                                                                float* __restrict__ final_T,
                                                                uint32_t* __restrict__ n_contrib) {
  constexpr int NT = 256 / K;
  __shared__ SplatRec buf[2][NT];
  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int t = threadIdx.x;
  const int pixx = tx * GAB_TILE + (t & 15);
  const int pixy0 = ty * GAB_TILE + (t >> 4) * K;
  const float fx = (float)pixx, fy0 = (float)pixy0;
  const uint2 range = ranges[tile];
  const int n = (int)(range.y - range.x);
  const uint32_t* ids = point_list + range.x;

  float T[K], Cr[K], Cg[K], Cb[K];
  uint32_t last[K];
  uint32_t done = 0;
  constexpr uint32_t ALL = (1u << K) - 1u;
#pragma unroll
  for (int i = 0; i < K; i++) {
    T[i] = 1.f; Cr[i] = Cg[i] = Cb[i] = 0.f; last[i] = 0;
    if (pixx >= W || pixy0 + i >= H) done |= 1u << i;
  }

  const int nchunks = (n + NT - 1) / NT;
  // prologue: ids of chunk 0 -> gather chunk 0; ids of chunk 1 in flight
  uint32_t id_next = (t < n) ? ids[t] : 0xffffffffu;
  if (id_next != 0xffffffffu) gather_rec(&buf[0][t], rec + id_next);
  cp_async_commit();
  id_next = (NT + t < n) ? ids[NT + t] : 0xffffffffu;

  for (int c = 0; c < nchunks; c++) {
    if (c + 1 < nchunks && id_next != 0xffffffffu) gather_rec(&buf[(c + 1) & 1][t], rec + id_next);
    cp_async_commit();
    {
      const int p = (c + 2) * NT + t;
      id_next = (p < n) ? ids[p] : 0xffffffffu;
    }
    cp_async_wait<1>();
    if (__syncthreads_and(done == ALL)) break;  // also publishes chunk c to the CTA
    const SplatRec* cur = buf[c & 1];
    const int cnt = min(NT, n - c * NT);
    const uint32_t pos0 = (uint32_t)(c * NT);
    for (int j = 0; j < cnt; j++) {
      if (done == ALL) break;
      const float4 q0 = cur[j].q0;
      const float4 q1 = cur[j].q1;
      const float cb = cur[j].q2.x;
      const float dx = q0.x - fx, dy0 = q0.y - fy0;
      const float p0 = (-0.5f * LOG2E) * q0.z * dx * dx;
      const float qq = (-LOG2E) * q0.w * dx;
      const float hh = (-0.5f * LOG2E) * q1.x;
      const float op = q1.y;
#pragma unroll
      for (int i = 0; i < K; i++) {
        const float dy = dy0 - (float)i;
        const float pw = fmaf(dy, fmaf(hh, dy, qq), p0);
        const float alpha = fminf(0.99f, op * ex2_approx(pw));
        if (!((done >> i) & 1u) && pw <= 0.f && alpha >= ALPHA_MIN) {
          const float test_T = T[i] * (1.f - alpha);
          if (test_T < 0.0001f) {
            done |= 1u << i;
          } else {
            const float w = alpha * T[i];
            Cr[i] = fmaf(q1.z, w, Cr[i]);
            Cg[i] = fmaf(q1.w, w, Cg[i]);
            Cb[i] = fmaf(cb, w, Cb[i]);
            T[i] = test_T;
            last[i] = pos0 + (uint32_t)j + 1u;
          }
        }
      }
    }
    __syncthreads();  // everyone is done with buf[c&1] before chunk c+2 is gathered into it
  }
  cp_async_wait<0>();

  const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
  const size_t HW = (size_t)H * W;
#pragma unroll
  for (int i = 0; i < K; i++) {
    const int y = pixy0 + i;
    if (pixx < W && y < H) {
      const size_t pix = (size_t)y * W + pixx;
      out_color[pix] = fmaf(T[i], bg0, Cr[i]);
      out_color[HW + pix] = fmaf(T[i], bg1, Cg[i]);
      out_color[2 * HW + pix] = fmaf(T[i], bg2, Cb[i]);
      if (final_T != nullptr) {
        final_T[pix] = T[i];
        n_contrib[pix] = last[i];
      }
    }
  }
}

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

void launch_blend_forward(int W, int H, const uint2* ranges, const uint32_t* point_list, const SplatRec* rec,
                          const float* bg, float* out_color, float* final_T, uint32_t* n_contrib,
                          cudaStream_t stream) {
  const int gx = (W + GAB_TILE - 1) / GAB_TILE, gy = (H + GAB_TILE - 1) / GAB_TILE;
  const int tiles = gx * gy;
  if (tiles == 0) return;
  static const int K = env_int("GAB200_FWD_K", 4);
  switch (K) {
    case 1: blend_forward_kernel<1><<<tiles, 256, 0, stream>>>(W, H, gx, ranges, point_list, rec, bg, out_color, final_T, n_contrib); break;
    case 2: blend_forward_kernel<2><<<tiles, 128, 0, stream>>>(W, H, gx, ranges, point_list, rec, bg, out_color, final_T, n_contrib); break;
    case 8: blend_forward_kernel<8><<<tiles, 32, 0, stream>>>(W, H, gx, ranges, point_list, rec, bg, out_color, final_T, n_contrib); break;
    default: blend_forward_kernel<4><<<tiles, 64, 0, stream>>>(W, H, gx, ranges, point_list, rec, bg, out_color, final_T, n_contrib); break;
  }
  count_launch();
}

// =====================================================================================================
// Backward
// =====================================================================================================
// Multi-value butterfly: reduces v[0..7] across the 32 lanes with 4+2+1+1+1 = 9 shuffles.  On return every lane
// holds the warp total of component (lane >> 2).
__device__ __forceinline__ float warp_reduce8(const float v[8], int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  float w[4], u[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float send = h16 ? v[i] : v[i + 4];
    const float keep = h16 ? v[i + 4] : v[i];
    w[i] = keep + __shfl_xor_sync(FULL, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const float send = h8 ? w[i] : w[i + 2];
    const float keep = h8 ? w[i + 2] : w[i];
    u[i] = keep + __shfl_xor_sync(FULL, send, 8);
  }
  const float send = h4 ? u[0] : u[1];
  const float keep = h4 ? u[1] : u[0];
  float r = keep + __shfl_xor_sync(FULL, send, 4);
  r += __shfl_xor_sync(FULL, r, 2);
  r += __shfl_xor_sync(FULL, r, 1);
  return r;
}
__device__ __forceinline__ float warp_reduce1(float v) {
  constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
  return v;
}

template <int K>
__global__ void __launch_bounds__(256 / K) blend_backward_kernel(int W, int H, int gx, const uint2* __restrict__ ranges,
                                                                 const uint32_t* __restrict__ point_list,
                                                                 const SplatRec* __restrict__ rec,
                                                                 const float* __restrict__ bg,
                                                                 const float* __restrict__ final_T,
                                                                 const uint32_t* __restrict__ n_contrib,
                                                                 const float* __restrict__ dL_dpix,
                                                                 float* __restrict__ g2d) {
  constexpr int NT = 256 / K;
  __shared__ SplatRec buf[2][NT];
  __shared__ uint32_t buf_id[2][NT];
  __shared__ int s_max;
  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int t = threadIdx.x, lane = t & 31;
  const int pixx = tx * GAB_TILE + (t & 15);
  const int pixy0 = ty * GAB_TILE + (t >> 4) * K;
  const float fx = (float)pixx, fy0 = (float)pixy0;
  const uint2 range = ranges[tile];
  const uint32_t* ids = point_list + range.x;
  const size_t HW = (size_t)H * W;
  const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];

  float T[K], ar[K], ag[K], ab[K], dr[K], dg[K], db[K], bgT[K];
  int nc[K];
  int my_max = 0;
#pragma unroll
  for (int i = 0; i < K; i++) {
    const int y = pixy0 + i;
    ar[i] = ag[i] = ab[i] = 0.f;
    if (pixx < W && y < H) {
      const size_t pix = (size_t)y * W + pixx;
      T[i] = final_T[pix];
      nc[i] = (int)n_contrib[pix];
      dr[i] = dL_dpix[pix];
      dg[i] = dL_dpix[HW + pix];
      db[i] = dL_dpix[2 * HW + pix];
    } else {
      T[i] = 0.f; nc[i] = 0; dr[i] = dg[i] = db[i] = 0.f;
    }
    bgT[i] = T[i] * (bg0 * dr[i] + bg1 * dg[i] + bg2 * db[i]);
    my_max = max(my_max, nc[i]);
  }
  // the tile only needs instances [0, max n_contrib): nothing behind the last contributor of any pixel matters
  if (t == 0) s_max = 0;
  __syncthreads();
  my_max = __reduce_max_sync(0xffffffffu, my_max);
  if (lane == 0 && my_max > 0) atomicMax(&s_max, my_max);
  __syncthreads();
  const int n = s_max;
  if (n == 0) return;
  const int nchunks = (n + NT - 1) / NT;
  const float half_W = 0.5f * (float)W, half_H = 0.5f * (float)H;

  // reverse walk: chunk c covers positions n-1-c*NT-j (j = 0..NT-1)
  uint32_t id_next = (t < n) ? ids[n - 1 - t] : 0xffffffffu;
  if (id_next != 0xffffffffu) gather_rec(&buf[0][t], rec + id_next);
  buf_id[0][t] = id_next;
  cp_async_commit();
  id_next = (NT + t < n) ? ids[n - 1 - NT - t] : 0xffffffffu;

  for (int c = 0; c < nchunks; c++) {
    if (c + 1 < nchunks) {
      if (id_next != 0xffffffffu) gather_rec(&buf[(c + 1) & 1][t], rec + id_next);
      buf_id[(c + 1) & 1][t] = id_next;
    }
    cp_async_commit();
    {
      const int p = (c + 2) * NT + t;
      id_next = (p < n) ? ids[n - 1 - p] : 0xffffffffu;
    }
    cp_async_wait<1>();
    __syncthreads();
    const SplatRec* cur = buf[c & 1];
    const uint32_t* cur_id = buf_id[c & 1];
    const int cnt = min(NT, n - c * NT);
    for (int j = 0; j < cnt; j++) {
      const int pos = n - 1 - c * NT - j;  // 0-based position in the tile's list; contributes to pixel iff pos < nc
      const float4 q0 = cur[j].q0;
      const float4 q1 = cur[j].q1;
      const float cbl = cur[j].q2.x;
      const float dx = q0.x - fx, dy0 = q0.y - fy0;
      const float A = q0.z, B = q0.w, C = q1.x, op = q1.y;
      const float p0 = (-0.5f * LOG2E) * A * dx * dx;
      const float qq = (-LOG2E) * B * dx;
      const float hh = (-0.5f * LOG2E) * C;
      float S0 = 0.f, S1 = 0.f, S2 = 0.f, go = 0.f, gr = 0.f, gg = 0.f, gb = 0.f;
      bool any = false;
#pragma unroll
      for (int i = 0; i < K; i++) {
        const float dy = dy0 - (float)i;
        const float pw = fmaf(dy, fmaf(hh, dy, qq), p0);
        const float G = ex2_approx(pw);
        const float alpha = fminf(0.99f, op * G);
        if (pos < nc[i] && pw <= 0.f && alpha >= ALPHA_MIN) {
          any = true;
          const float ra = rcp_approx(1.f - alpha);
          T[i] *= ra;  // transmittance in FRONT of this splat
          const float w = alpha * T[i];
          gr = fmaf(w, dr[i], gr);
          gg = fmaf(w, dg[i], gg);
          gb = fmaf(w, db[i], gb);
          // dL/dalpha = T * sum_ch (c - colour behind) dpix  -  T_final/(1-alpha) * (bg . dpix)
          float dLda = (q1.z - ar[i]) * dr[i];
          dLda = fmaf(q1.w - ag[i], dg[i], dLda);
          dLda = fmaf(cbl - ab[i], db[i], dLda);
          dLda = fmaf(dLda, T[i], -bgT[i] * ra);
          // colour behind the NEXT (nearer) splat: this one composited over what was behind it
          ar[i] = fmaf(alpha, q1.z - ar[i], ar[i]);
          ag[i] = fmaf(alpha, q1.w - ag[i], ag[i]);
          ab[i] = fmaf(alpha, cbl - ab[i], ab[i]);
          go = fmaf(G, dLda, go);
          const float s = G * op * dLda;  // G * dL/dG
          const float sd = s * dy;
          S0 += s;
          S1 += sd;
          S2 = fmaf(sd, dy, S2);
        }
      }
      if (!__any_sync(0xffffffffu, any)) continue;
      float v[8];
      v[0] = (-A * dx * S0 - B * S1) * half_W;  // dL/dmean2D.x (NDC units)
      v[1] = (-C * S1 - B * dx * S0) * half_H;  // dL/dmean2D.y
      v[2] = -0.5f * dx * dx * S0;              // dL/dconic.xx
      v[3] = -0.5f * dx * S1;                   // dL/dconic.xy (stored once)
      v[4] = -0.5f * S2;                        // dL/dconic.yy
      v[5] = go;                                // dL/dopacity
      v[6] = gr;
      v[7] = gg;
      const float r8 = warp_reduce8(v, lane);
      const float r1 = warp_reduce1(gb);
      const uint32_t id = cur_id[j];
      if ((lane & 3) == 0)
        atomicAdd(g2d + (size_t)id * GAB_G2D_STRIDE + (lane >> 2), r8);
      else if (lane == 1)
        atomicAdd(g2d + (size_t)id * GAB_G2D_STRIDE + 8, r1);
    }
    __syncthreads();
  }
  cp_async_wait<0>();
}

void launch_blend_backward(int W, int H, const uint2* ranges, const uint32_t* point_list, const SplatRec* rec,
                           const float* bg, const float* final_T, const uint32_t* n_contrib, const float* dL_dpix,
                           float* g2d, cudaStream_t stream) {
  const int gx = (W + GAB_TILE - 1) / GAB_TILE, gy = (H + GAB_TILE - 1) / GAB_TILE;
  const int tiles = gx * gy;
  if (tiles == 0) return;
  static const int K = env_int("GAB200_BWD_K", 4);
  switch (K) {
    case 1: blend_backward_kernel<1><<<tiles, 256, 0, stream>>>(W, H, gx, ranges, point_list, rec, bg, final_T, n_contrib, dL_dpix, g2d); break;
    case 2: blend_backward_kernel<2><<<tiles, 128, 0, stream>>>(W, H, gx, ranges, point_list, rec, bg, final_T, n_contrib, dL_dpix, g2d); break;
    case 8: blend_backward_kernel<8><<<tiles, 32, 0, stream>>>(W, H, gx, ranges, point_list, rec, bg, final_T, n_contrib, dL_dpix, g2d); break;
    default: blend_backward_kernel<4><<<tiles, 64, 0, stream>>>(W, H, gx, ranges, point_list, rec, bg, final_T, n_contrib, dL_dpix, g2d); break;
  }
  count_launch();
}

}  // namespace gab
