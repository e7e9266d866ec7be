// Micro-benchmark behind a design decision (north_star "warp-shuffle prefix scans for
// transmittance"): the forward blend of one 16x16 tile written two ways over the same synthetic tile lists,
//   A  pixel-parallel  one thread per pixel walks the list sequentially (T is a running product in a register;
//                      the reference's semantics: skip alpha < 1/255, stop when T(1-alpha) < 1e-4)
//   B  splat-parallel  a warp takes one pixel at a time, lane = splat: 32 alphas at once, the transmittance in front
//                      of each splat by a 5-step warp-shuffle prefix product of (1 - alpha), the stop rule from a
//                      ballot on the scanned products, colour accumulated per lane and reduced once per pixel
// and checks that both produce the same image (the scan multiplies in a different association: ~1e-6 differences).
// Standalone (no library): nvcc -gencode arch=compute_100a,code=sm_100a -O3 examples/scan_blend_probe.cu -o scan_probe
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct Rec { float px, py, A, B, C, op, r, g, b, pad[3]; };  // 48 B like the library's SplatRec (conic in natural units)

__global__ void __launch_bounds__(256) blend_pixel_parallel(const Rec* __restrict__ rec, const int* __restrict__ start,
                                                            float* __restrict__ out) {
  __shared__ Rec buf[256];
  const int tile = blockIdx.x, t = threadIdx.x;
  const int lo = start[tile], n = start[tile + 1] - lo;
  const float fx = (float)(t & 15), fy = (float)(t >> 4);
  float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f;
  bool done = false;
  for (int base = 0; base < n; base += 256) {
    __syncthreads();
    if (base + t < n) buf[t] = rec[lo + base + t];
    __syncthreads();
    if (__syncthreads_and(done)) break;
    const int cnt = min(256, n - base);
    for (int j = 0; j < cnt && !done; j++) {
      const Rec& s = buf[j];
      const float dx = s.px - fx, dy = s.py - fy;
      const float pw = -0.5f * (s.A * dx * dx + s.C * dy * dy) - s.B * dx * dy;
      if (pw > 0.f) continue;
      const float a = fminf(0.99f, s.op * __expf(pw));
      if (a < 1.f / 255.f) continue;
      const float Tn = T * (1.f - a);
      if (Tn < 1e-4f) { done = true; break; }
      const float w = a * T;
      cr += s.r * w; cg += s.g * w; cb += s.b * w;
      T = Tn;
    }
  }
  float* o = out + ((size_t)tile * 256 + t) * 4;
  o[0] = cr; o[1] = cg; o[2] = cb; o[3] = T;
}

__global__ void __launch_bounds__(256) blend_splat_parallel(const Rec* __restrict__ rec, const int* __restrict__ start,
                                                            float* __restrict__ out) {
  __shared__ Rec buf[256];
  const int tile = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int lo = start[tile], n = start[tile + 1] - lo;
  // warp w owns pixels 32 w .. 32 w + 31 of the tile; per pixel state of the chunk loop lives in lane p's registers
  float T_pix = 1.f, cr_pix = 0.f, cg_pix = 0.f, cb_pix = 0.f;
  bool done_pix = false;
  for (int base = 0; base < n; base += 256) {
    __syncthreads();
    if (base + t < n) buf[t] = rec[lo + base + t];
    __syncthreads();
    if (__syncthreads_and(done_pix)) break;
    const int cnt = min(256, n - base);
    for (int p = 0; p < 32; p++) {  // one pixel at a time, lanes = splats
      if (__shfl_sync(0xffffffffu, (int)done_pix, p)) continue;
      const int pix = warp * 32 + p;
      const float fx = (float)(pix & 15), fy = (float)(pix >> 4);
      float T = __shfl_sync(0xffffffffu, T_pix, p);
      float cr = 0.f, cg = 0.f, cb = 0.f;
      bool stop = false;
      for (int j0 = 0; j0 < cnt && !stop; j0 += 32) {
        const int j = j0 + lane;
        float a = 0.f, r = 0.f, g = 0.f, b = 0.f;
        if (j < cnt) {
          const Rec& s = buf[j];
          const float dx = s.px - fx, dy = s.py - fy;
          const float pw = -0.5f * (s.A * dx * dx + s.C * dy * dy) - s.B * dx * dy;
          const float al = fminf(0.99f, s.op * __expf(pw));
          if (pw <= 0.f && al >= 1.f / 255.f) { a = al; r = s.r; g = s.g; b = s.b; }
        }
        // inclusive prefix product of (1 - alpha) over the lanes
        float q = 1.f - a;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const float up = __shfl_up_sync(0xffffffffu, q, d);
          if (lane >= d) q *= up;
        }
        const float Tin = T * q;                     // transmittance BEHIND this splat
        const unsigned dead = __ballot_sync(0xffffffffu, a > 0.f && Tin < 1e-4f);
        const int first = dead ? __ffs(dead) - 1 : 32;  // first splat that would drop T below 1e-4: it and all later ones are cut
        if (lane < first && a > 0.f) {
          const float Tfront = Tin / (1.f - a);      // = T * prod_{k<lane}(1 - a_k)
          const float w = a * Tfront;
          cr += r * w; cg += g * w; cb += b * w;
        }
        const int last = min(first, 32) - 1;
        if (last >= 0) T = __shfl_sync(0xffffffffu, Tin, last);
        stop = first < 32;
      }
      // one reduction per (pixel, chunk)
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        cr += __shfl_xor_sync(0xffffffffu, cr, m);
        cg += __shfl_xor_sync(0xffffffffu, cg, m);
        cb += __shfl_xor_sync(0xffffffffu, cb, m);
      }
      if (lane == p) { T_pix = T; cr_pix += cr; cg_pix += cg; cb_pix += cb; done_pix = stop; }
    }
  }
  float* o = out + ((size_t)tile * 256 + t) * 4;
  o[0] = cr_pix; o[1] = cg_pix; o[2] = cb_pix; o[3] = T_pix;
}

int main() {
  const int tiles = 2048;
  std::srand(7);
  auto rnd = [] { return (float)std::rand() / (float)RAND_MAX; };
  std::vector<int> start(tiles + 1, 0);
  for (int t = 0; t < tiles; t++) start[t + 1] = start[t] + 64 + (int)(700.f * rnd() * rnd());  // mean ~ 240, tail to 760
  std::vector<Rec> rec((size_t)start[tiles]);
  for (auto& s : rec) {
    s.px = -8.f + 32.f * rnd(); s.py = -8.f + 32.f * rnd();
    const float sig = 2.f + 6.f * rnd();             // 3-sigma radius 6..24 px
    s.A = s.C = 1.f / (sig * sig); s.B = 0.3f * s.A * (rnd() - 0.5f);
    s.op = rnd() < 0.1f ? 0.002f : 0.05f + 0.9f * rnd();
    s.r = rnd(); s.g = rnd(); s.b = rnd();
  }
  Rec* d_rec; int* d_start; float *d_a, *d_b;
  cudaMalloc(&d_rec, rec.size() * sizeof(Rec)); cudaMalloc(&d_start, start.size() * sizeof(int));
  cudaMalloc(&d_a, (size_t)tiles * 256 * 4 * sizeof(float)); cudaMalloc(&d_b, (size_t)tiles * 256 * 4 * sizeof(float));
  cudaMemcpy(d_rec, rec.data(), rec.size() * sizeof(Rec), cudaMemcpyHostToDevice);
  cudaMemcpy(d_start, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms[2] = {0, 0};
  for (int which = 0; which < 2; which++) {
    for (int it = 0; it < 3; it++) {
      if (which == 0) blend_pixel_parallel<<<tiles, 256>>>(d_rec, d_start, d_a);
      else blend_splat_parallel<<<tiles, 256>>>(d_rec, d_start, d_b);
    }
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int it = 0; it < 20; it++) {
      if (which == 0) blend_pixel_parallel<<<tiles, 256>>>(d_rec, d_start, d_a);
      else blend_splat_parallel<<<tiles, 256>>>(d_rec, d_start, d_b);
    }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms[which], e0, e1);
    ms[which] /= 20.f;
  }
  if (cudaGetLastError() != cudaSuccess) { std::fprintf(stderr, "CUDA error\n"); return 1; }
  std::vector<float> a((size_t)tiles * 1024), b(a.size());
  cudaMemcpy(a.data(), d_a, a.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(b.data(), d_b, b.size() * 4, cudaMemcpyDeviceToHost);
  double maxd = 0; size_t big = 0;
  for (size_t i = 0; i < a.size(); i++) { const double d = std::fabs((double)a[i] - b[i]); maxd = d > maxd ? d : maxd; big += d > 1e-4; }
  std::printf("{\"tiles\": %d, \"instances\": %d, \"pixel_parallel_ms\": %.4f, \"splat_parallel_scan_ms\": %.4f, "
              "\"ratio\": %.2f, \"max_abs_diff\": %.3e, \"values_over_1e-4\": %zu, \"values\": %zu}\n",
              tiles, start[tiles], ms[0], ms[1], ms[1] / ms[0], maxd, big, a.size());
  return 0;
}
