// preprocess.cu -- per-splat forward stage with the FLAME mesh binding fused in, plus tile|depth key emission
// and tile-range detection.  COMPILED WITH --fmad=false: every float expression rounds exactly as written
// (IEEE rn, no contraction), and the association of every sum is the documented left-to-right order, so the
// depth bits, pixel centres, radii and tile rectangles -- hence the tile|depth keys -- are reproducible bit for
// bit by the CPU oracle (oracle/splat_oracle.c, built with -ffp-contract=off).
//
// Replaces, in ONE kernel (SURVEY.md 2.4 K1 + the eager getters of 2.4(b)):
//   scene/gaussian_model.py:113-160   get_xyz / get_rotation / get_scaling / get_opacity / get_features
//   diff_gaussian_rasterization preprocessCUDA (absent submodule; behaviour: SURVEY.md Appendix B.1)
#include "common.cuh"
#include "kernels.cuh"
#include "splat_math.cuh"

namespace gab {

__device__ __forceinline__ void stage_camera(const gab200_forward_args& a, Camera& cam) {
  int t = threadIdx.x;
  if (t < 16) cam.V[t] = a.viewmatrix[t];
  else if (t < 32) cam.Pm[t - 16] = a.projmatrix[t - 16];
  else if (t < 35) cam.campos[t - 32] = a.campos[t - 32];
  __syncthreads();
}

// =====================================================================================================
// K1: fused bind + activate + project + EWA + SH->RGB.  One thread per splat.
// =====================================================================================================
template <bool BOUND>
__global__ void __launch_bounds__(256) preprocess_kernel(gab200_forward_args a, SplatRec* __restrict__ rec,
                                                         uint32_t* __restrict__ tiles_touched,
                                                         uint8_t* __restrict__ clamped, int exact_binning) {
  __shared__ Camera cam;
  stage_camera(a, cam);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.P) return;
  const int W = a.image_width, H = a.image_height;
  const int gx = (W + GAB_TILE - 1) / GAB_TILE, gy = (H + GAB_TILE - 1) / GAB_TILE;

  SplatRec out;
  out.q0 = make_float4(0.f, 0.f, 0.f, 0.f);
  out.q1 = make_float4(0.f, 0.f, 0.f, 0.f);
  out.q2 = make_float4(0.f, 0.f, 0.f, 0.f);
  int radius_out = 0;
  uint32_t tiles_out = 0;
  uint8_t clamp_bits = 0;

  float3 p;
  float opacity;
  float c3[6];
  bool have_cov = false;
  if (BOUND) {
    Activated act;
    bind_activate(a, i, act);
    p = act.mean;
    opacity = act.opacity;
    float s[3] = {a.scale_modifier * act.s[0], a.scale_modifier * act.s[1], a.scale_modifier * act.s[2]};
    cov3d_from_R(act.R, s, c3);
    have_cov = true;
  } else {
    p = make_float3(a.means3D[3 * i], a.means3D[3 * i + 1], a.means3D[3 * i + 2]);
    opacity = a.opacities[i];
  }

  const float3 t = xform4x3(cam.V, p);
  if (t.z > 0.2f) {
    const float hx = cam.Pm[0] * p.x + cam.Pm[4] * p.y + cam.Pm[8] * p.z + cam.Pm[12];
    const float hy = cam.Pm[1] * p.x + cam.Pm[5] * p.y + cam.Pm[9] * p.z + cam.Pm[13];
    const float hw = cam.Pm[3] * p.x + cam.Pm[7] * p.y + cam.Pm[11] * p.z + cam.Pm[15];
    const float p_w = 1.0f / (hw + 0.0000001f);
    const float ndc_x = hx * p_w, ndc_y = hy * p_w;

    if (!have_cov) {
      if (a.cov3D_precomp != nullptr) {
#pragma unroll
        for (int k = 0; k < 6; k++) c3[k] = a.cov3D_precomp[6 * (size_t)i + k];
      } else {
        float R[9];
        quat_to_R(a.rotations[4 * i], a.rotations[4 * i + 1], a.rotations[4 * i + 2], a.rotations[4 * i + 3], R);
        float s[3] = {a.scale_modifier * a.scales[3 * i], a.scale_modifier * a.scales[3 * i + 1],
                      a.scale_modifier * a.scales[3 * i + 2]};
        cov3d_from_R(R, s, c3);
      }
    }

    const float focal_x = (float)W / (2.0f * a.tanfovx), focal_y = (float)H / (2.0f * a.tanfovy);
    const float limx = 1.3f * a.tanfovx, limy = 1.3f * a.tanfovy;
    const float txtz = t.x / t.z, tytz = t.y / t.z;
    const float tcx = fminf(limx, fmaxf(-limx, txtz)) * t.z;
    const float tcy = fminf(limy, fmaxf(-limy, tytz)) * t.z;
    const float j00 = focal_x / t.z, j02 = -(focal_x * tcx) / (t.z * t.z);
    const float j11 = focal_y / t.z, j12 = -(focal_y * tcy) / (t.z * t.z);
    float T0[3], T1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      T0[c] = j00 * cam.V[4 * c + 0] + j02 * cam.V[4 * c + 2];
      T1[c] = j11 * cam.V[4 * c + 1] + j12 * cam.V[4 * c + 2];
    }
    const float S[9] = {c3[0], c3[1], c3[2], c3[1], c3[3], c3[4], c3[2], c3[4], c3[5]};
    float u[3], v[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      u[r] = S[3 * r + 0] * T0[0] + S[3 * r + 1] * T0[1] + S[3 * r + 2] * T0[2];
      v[r] = S[3 * r + 0] * T1[0] + S[3 * r + 1] * T1[1] + S[3 * r + 2] * T1[2];
    }
    float ca = T0[0] * u[0] + T0[1] * u[1] + T0[2] * u[2];
    float cb = T0[0] * v[0] + T0[1] * v[1] + T0[2] * v[2];
    float cc = T1[0] * v[0] + T1[1] * v[1] + T1[2] * v[2];
    ca += 0.3f;
    cc += 0.3f;
    const float det = ca * cc - cb * cb;
    if (det != 0.0f) {
      const float det_inv = 1.f / det;
      const float conic_x = cc * det_inv, conic_y = -cb * det_inv, conic_z = ca * det_inv;
      const float mid = 0.5f * (ca + cc);
      const float sq = sqrtf(fmaxf(0.1f, mid * mid - det));
      const float lambda1 = mid + sq, lambda2 = mid - sq;
      const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
      const float px = ndc2pix(ndc_x, W), py = ndc2pix(ndc_y, H);
      int x0, y0, x1, y1;
      tile_rect(px, py, (int)my_radius, gx, gy, x0, y0, x1, y1);
      if ((x1 - x0) * (y1 - y0) != 0) {
        float rgb[3];
        if (a.colors_precomp != nullptr) {
          rgb[0] = a.colors_precomp[3 * (size_t)i];
          rgb[1] = a.colors_precomp[3 * (size_t)i + 1];
          rgb[2] = a.colors_precomp[3 * (size_t)i + 2];
        } else {
          float3 d = make_float3(p.x - cam.campos[0], p.y - cam.campos[1], p.z - cam.campos[2]);
          const float len = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
          d.x = d.x / len; d.y = d.y / len; d.z = d.z / len;
          float B[16];
          sh_basis(a.sh_degree, d, B);
          const int nb = (a.sh_degree + 1) * (a.sh_degree + 1);
          float acc[3] = {0.f, 0.f, 0.f};
          if (BOUND) {
            const float* dc = a.sh_dc + 3 * (size_t)i;
            acc[0] = acc[0] + B[0] * dc[0];
            acc[1] = acc[1] + B[0] * dc[1];
            acc[2] = acc[2] + B[0] * dc[2];
            const float* rest = a.sh_rest + (size_t)i * (a.sh_coeffs - 1) * 3;
            for (int k = 1; k < nb; k++) {
              acc[0] = acc[0] + B[k] * rest[3 * (k - 1) + 0];
              acc[1] = acc[1] + B[k] * rest[3 * (k - 1) + 1];
              acc[2] = acc[2] + B[k] * rest[3 * (k - 1) + 2];
            }
          } else {
            const float* sh = a.shs + (size_t)i * a.sh_coeffs * 3;
            for (int k = 0; k < nb; k++) {
              acc[0] = acc[0] + B[k] * sh[3 * k + 0];
              acc[1] = acc[1] + B[k] * sh[3 * k + 1];
              acc[2] = acc[2] + B[k] * sh[3 * k + 2];
            }
          }
#pragma unroll
          for (int ch = 0; ch < 3; ch++) {
            float r = acc[ch] + 0.5f;
            if (r < 0.f) clamp_bits |= (1u << ch);
            rgb[ch] = fmaxf(r, 0.f);
          }
        }
        radius_out = (int)my_radius;
        out.q0 = make_float4(px, py, conic_x, conic_y);
        out.q1 = make_float4(conic_z, opacity, rgb[0], rgb[1]);
        if (exact_binning) {
          tiles_out = (uint32_t)((y1 - y0) * (x1 - x0));
        } else {
          TileSpan span(px, py, conic_x, conic_y, conic_z, opacity, x0, x1);
          uint32_t cnt = 0;
          if (span.any)
            for (int ty = y0; ty < y1; ty++) {
              int cx0, cx1;
              span.row(ty, cx0, cx1);
              cnt += (uint32_t)(cx1 - cx0);
            }
          tiles_out = cnt;
        }
        out.q2 = make_float4(rgb[2], t.z, __int_as_float(radius_out), __int_as_float((int)tiles_out));
      }
    }
  }
  rec[i] = out;
  a.radii[i] = radius_out;
  tiles_touched[i] = tiles_out;
  if (clamped != nullptr) clamped[i] = clamp_bits;
}

void launch_preprocess(const gab200_forward_args& a, SplatRec* rec, uint32_t* tiles_touched, uint8_t* clamped,
                       cudaStream_t stream) {
  const int threads = 256, blocks = (a.P + threads - 1) / threads;
  if (blocks == 0) return;
  if (a.input_mode == GAB200_INPUT_BOUND_RAW)
    preprocess_kernel<true><<<blocks, threads, 0, stream>>>(a, rec, tiles_touched, clamped, a.exact_binning);
  else
    preprocess_kernel<false><<<blocks, threads, 0, stream>>>(a, rec, tiles_touched, clamped, a.exact_binning);
  count_launch();
}

// =====================================================================================================
// Export of the binding + activation only (gab200_bind_activate)
// =====================================================================================================
__global__ void __launch_bounds__(256) bind_activate_kernel(gab200_forward_args a, float* __restrict__ means3D,
                                                            float* __restrict__ opacities,
                                                            float* __restrict__ scales, float* __restrict__ cov3D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.P) return;
  Activated act;
  bind_activate(a, i, act);
  if (means3D) {
    means3D[3 * (size_t)i] = act.mean.x;
    means3D[3 * (size_t)i + 1] = act.mean.y;
    means3D[3 * (size_t)i + 2] = act.mean.z;
  }
  if (opacities) opacities[i] = act.opacity;
  if (scales) {
    scales[3 * (size_t)i] = act.s[0];
    scales[3 * (size_t)i + 1] = act.s[1];
    scales[3 * (size_t)i + 2] = act.s[2];
  }
  if (cov3D) {
    float s[3] = {a.scale_modifier * act.s[0], a.scale_modifier * act.s[1], a.scale_modifier * act.s[2]};
    float c3[6];
    cov3d_from_R(act.R, s, c3);
#pragma unroll
    for (int k = 0; k < 6; k++) cov3D[6 * (size_t)i + k] = c3[k];
  }
}

void launch_bind_activate(const gab200_forward_args& a, float* means3D, float* opacities, float* scales, float* cov3D,
                          cudaStream_t stream) {
  const int threads = 256, blocks = (a.P + threads - 1) / threads;
  if (blocks == 0) return;
  bind_activate_kernel<<<blocks, threads, 0, stream>>>(a, means3D, opacities, scales, cov3D);
  count_launch();
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ V,
                                    uint8_t* __restrict__ present) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float3 p = make_float3(means3D[3 * (size_t)i], means3D[3 * (size_t)i + 1], means3D[3 * (size_t)i + 2]);
  float z = V[2] * p.x + V[6] * p.y + V[10] * p.z + V[14];
  present[i] = z > 0.2f ? 1 : 0;
}

void launch_mark_visible(int P, const float* means3D, const float* V, uint8_t* present, cudaStream_t stream) {
  if (P == 0) return;
  mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, V, present);
  count_launch();
}

// =====================================================================================================
// K3: tile|depth key emission.  One WARP per 32 splats; the lanes cooperate on each splat's tile list so that
// stores are coalesced and a 256-tile splat does not serialise one thread.  Emission order inside a splat is
// row-major (y, then x) from offsets[i-1] -- identical to the reference, so the stable sort's tie order is too.
// =====================================================================================================
__global__ void __launch_bounds__(256) emit_keys_kernel(int P, int gx, int gy, const SplatRec* __restrict__ rec,
                                                        const uint32_t* __restrict__ offsets,
                                                        uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                        int exact_binning) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int i = warp_global * 32 + lane;
  float px = 0.f, py = 0.f, cA = 0.f, cB = 0.f, cC = 0.f, op = 0.f, depth = 0.f;
  int radius = 0;
  uint32_t ntiles = 0, off = 0;
  if (i < P) {
    const float4 q2 = rec[i].q2;
    ntiles = (uint32_t)__float_as_int(q2.w);
    if (ntiles) {
      const float4 q0 = rec[i].q0;
      const float4 q1 = rec[i].q1;
      px = q0.x; py = q0.y; cA = q0.z; cB = q0.w; cC = q1.x; op = q1.y;
      depth = q2.y;
      radius = __float_as_int(q2.z);
      off = (i == 0) ? 0u : offsets[i - 1];
    }
  }
  uint32_t todo = __ballot_sync(0xffffffffu, ntiles != 0);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const float spx = __shfl_sync(0xffffffffu, px, src), spy = __shfl_sync(0xffffffffu, py, src);
    const int srad = __shfl_sync(0xffffffffu, radius, src);
    const uint32_t soff = __shfl_sync(0xffffffffu, off, src);
    const uint32_t sdepth = __float_as_uint(__shfl_sync(0xffffffffu, depth, src));
    const uint32_t sid = (uint32_t)(warp_global * 32 + src);
    int x0, y0, x1, y1;
    tile_rect(spx, spy, srad, gx, gy, x0, y0, x1, y1);
    if (exact_binning) {
      const int w = x1 - x0, cnt = w * (y1 - y0);
      for (int t = lane; t < cnt; t += 32) {
        const int y = y0 + t / w, x = x0 + t % w;
        keys[soff + t] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | sdepth;
        vals[soff + t] = sid;
      }
    } else {
      const float sA = __shfl_sync(0xffffffffu, cA, src), sB = __shfl_sync(0xffffffffu, cB, src);
      const float sC = __shfl_sync(0xffffffffu, cC, src), sop = __shfl_sync(0xffffffffu, op, src);
      TileSpan span(spx, spy, sA, sB, sC, sop, x0, x1);
      // rows are walked by all lanes together; each row's span is written by lanes 0..len-1 (len <= grid width)
      uint32_t o = soff;
      for (int ty = y0; ty < y1; ty++) {
        int cx0, cx1;
        span.row(ty, cx0, cx1);
        const int len = cx1 - cx0;
        for (int t = lane; t < len; t += 32) {
          keys[o + t] = ((uint64_t)(uint32_t)(ty * gx + cx0 + t) << 32) | sdepth;
          vals[o + t] = sid;
        }
        o += (uint32_t)len;
      }
    }
  }
}

void launch_emit_keys(int P, int gx, int gy, const SplatRec* rec, const uint32_t* offsets, uint64_t* keys,
                      uint32_t* vals, int exact_binning, cudaStream_t stream) {
  const int warps = (P + 31) / 32;
  const int threads = 256, blocks = (warps * 32 + threads - 1) / threads;
  if (blocks == 0) return;
  emit_keys_kernel<<<blocks, threads, 0, stream>>>(P, gx, gy, rec, offsets, keys, vals, exact_binning);
  count_launch();
}

// =====================================================================================================
// K5: tile ranges from key transitions in the sorted stream (ranges pre-zeroed by the caller).
// =====================================================================================================
__global__ void __launch_bounds__(256) tile_ranges_kernel(int64_t N, const uint64_t* __restrict__ keys,
                                                          uint2* __restrict__ ranges) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N) return;
  const uint32_t cur = (uint32_t)(keys[idx] >> 32);
  if (idx == 0)
    ranges[cur].x = 0;
  else {
    const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
    if (cur != prev) {
      ranges[prev].y = (uint32_t)idx;
      ranges[cur].x = (uint32_t)idx;
    }
  }
  if (idx == N - 1) ranges[cur].y = (uint32_t)N;
}

void launch_tile_ranges(int64_t N, const uint64_t* keys, uint2* ranges, cudaStream_t stream) {
  if (N == 0) return;
  const int threads = 256;
  const int64_t blocks = (N + threads - 1) / threads;
  tile_ranges_kernel<<<(unsigned)blocks, threads, 0, stream>>>(N, keys, ranges);
  count_launch();
}

}  // namespace gab
