// A non-Python caller of the C ABI (include/gab200_rasterizer.h): renders P random splats (a first frame with the
// mid-frame sync, a second one sync-free on the hints the first left behind -- identical image), runs the backward,
// and with a path argument dumps inputs and results for tests/test_gpu_dropin.py to compare with the Python surface.
// No torch: device memory comes from cudaMalloc, the three scratch buffers from the allocation callbacks -- what
// INTEGRATION.md section 3 describes.  Build (from the repo root, after `python -m gaussianavatars_b200.build`):
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -Iinclude examples/abi_forward_backward.cu \
//        -Lgaussianavatars_b200 -lgaussianavatars_b200 -Xlinker -rpath=$PWD/gaussianavatars_b200 -o abi_example
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gab200_rasterizer.h"

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } \
  } while (0)

// scratch: the reference's geomBuffer / binningBuffer / imgBuffer "resize" lambdas.  The library asks once per buffer
// per frame; the caller owns the memory and keeps it alive until the frame's backward has run.
struct Arena {
  std::vector<void*> blocks;
  ~Arena() { for (void* p : blocks) cudaFree(p); }
};
static void* arena_alloc(void* user, size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
  static_cast<Arena*>(user)->blocks.push_back(p);
  return p;
}

template <typename T>
static T* upload(const std::vector<T>& h) {
  T* d = nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) return nullptr;
  cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  return d;
}

template <typename T>
static bool dump(std::FILE* f, const T* dev, size_t n) {
  std::vector<T> h(n);
  if (cudaMemcpy(h.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess) return false;
  return std::fwrite(h.data(), sizeof(T), n, f) == n;
}

int main(int argc, char** argv) {
  const int P = 20000, W = 640, H = 360;
  std::srand(1);
  auto rnd = [] { return (float)std::rand() / (float)RAND_MAX; };
  std::vector<float> means(3 * P), scales(3 * P), rots(4 * P), opac(P), rgb(3 * P);
  for (int i = 0; i < P; i++) {
    means[3 * i] = 4.f * rnd() - 2.f; means[3 * i + 1] = 2.f * rnd() - 1.f; means[3 * i + 2] = 3.f + 3.f * rnd();
    for (int k = 0; k < 3; k++) { scales[3 * i + k] = 0.01f + 0.04f * rnd(); rgb[3 * i + k] = rnd(); }
    rots[4 * i] = 1.f; rots[4 * i + 1] = rots[4 * i + 2] = rots[4 * i + 3] = 0.f;  // unit quaternion, wxyz
    opac[i] = 0.1f + 0.8f * rnd();
  }
  // camera at the origin looking down +z, row-vector convention (scene/cameras.py:44-47): identity view matrix,
  // projection = getProjectionMatrix(znear 0.01, zfar 100, fov 60 deg) transposed
  const float tanfov = std::tan(0.5f * 60.f * 3.14159265f / 180.f), aspect = (float)W / (float)H;
  const float tanx = tanfov * aspect, tany = tanfov, zn = 0.01f, zf = 100.f;
  std::vector<float> view = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  std::vector<float> proj = {1.f / tanx, 0, 0, 0, 0, 1.f / tany, 0, 0, 0, 0, zf / (zf - zn), 1, 0, 0, -(zf * zn) / (zf - zn), 0};
  std::vector<float> bg = {0, 0, 0}, campos = {0, 0, 0};

  gab200_forward_args a{};
  a.abi_version = GAB200_ABI_VERSION;
  a.input_mode = GAB200_INPUT_ACTIVATED;
  a.P = P; a.sh_degree = 0; a.sh_coeffs = 0;
  a.image_width = W; a.image_height = H; a.tanfovx = tanx; a.tanfovy = tany; a.scale_modifier = 1.f;
  a.need_backward = 1;
  a.bg = upload(bg); a.viewmatrix = upload(view); a.projmatrix = upload(proj); a.campos = upload(campos);
  a.means3D = upload(means); a.opacities = upload(opac); a.scales = upload(scales); a.rotations = upload(rots);
  a.colors_precomp = upload(rgb);
  float* out = nullptr; int32_t* radii = nullptr;
  CK(cudaMalloc(&out, sizeof(float) * 3 * W * H));
  CK(cudaMalloc(&radii, sizeof(int32_t) * P));
  a.out_color = out; a.radii = radii;
  Arena arena;
  a.alloc_geom = a.alloc_binning = a.alloc_image = arena_alloc;
  a.alloc_user = &arena;

  cudaStream_t stream;
  CK(cudaStreamCreate(&stream));
  gab200_frame_state st{};
  const int64_t n = gab200_forward(&a, &st, stream);
  if (n < 0) { std::fprintf(stderr, "gab200_forward: %s\n", gab200_status_string((int32_t)n)); return 1; }
  std::printf("forward: %lld (splat, tile) instances, depth-sort path %d, key range [%08x, %08x]\n", (long long)n,
              st.depth_sort_path, st.depth_key_min, st.depth_key_max);

  // second frame: the first one's instance count and depth-key range are the hints; nothing waits for the GPU until
  // the call's final check (GAB200_SYNC_LATE).  Same inputs -> the image must be bit-identical.
  std::vector<float> img1((size_t)3 * W * H), img2((size_t)3 * W * H);
  CK(cudaStreamSynchronize(stream));
  CK(cudaMemcpy(img1.data(), out, img1.size() * sizeof(float), cudaMemcpyDeviceToHost));
  a.sync_mode = GAB200_SYNC_LATE;
  a.binning_hint = (int32_t)(n + n / 4 + 4096);
  a.depth_hint_lo = st.depth_key_min;
  a.depth_hint_hi = st.depth_key_max;
  a.frame_seq = 2;
  const int64_t n2 = gab200_forward(&a, &st, stream);
  if (n2 < 0) { std::fprintf(stderr, "gab200_forward (2): %s\n", gab200_status_string((int32_t)n2)); return 1; }
  CK(cudaStreamSynchronize(stream));
  CK(cudaMemcpy(img2.data(), out, img2.size() * sizeof(float), cudaMemcpyDeviceToHost));
  size_t differing = 0;
  for (size_t i = 0; i < img1.size(); i++) differing += img1[i] != img2[i] ? 1 : 0;
  std::printf("second frame: %lld instances, depth-sort path %d, attempts %d, capacity %lld, pixels differing from frame 1: %zu\n",
              (long long)n2, st.depth_sort_path, st.attempts, (long long)st.binning_capacity, differing);
  if (n2 != n || differing != 0) { std::fprintf(stderr, "sync-free frame differs from the first\n"); return 1; }

  // backward of sum(image): dL/dimage = 1
  std::vector<float> ones((size_t)3 * W * H, 1.f);
  gab200_backward_args b{};
  b.abi_version = GAB200_ABI_VERSION;
  b.fwd = &a; b.state = &st;
  b.dL_dout_color = upload(ones);
  float *g_means = nullptr, *g_m2d = nullptr, *g_op = nullptr, *g_col = nullptr, *g_sc = nullptr, *g_rot = nullptr;
  CK(cudaMalloc(&g_means, sizeof(float) * 3 * P)); CK(cudaMalloc(&g_m2d, sizeof(float) * 3 * P));
  CK(cudaMalloc(&g_op, sizeof(float) * P));        CK(cudaMalloc(&g_col, sizeof(float) * 3 * P));
  CK(cudaMalloc(&g_sc, sizeof(float) * 3 * P));    CK(cudaMalloc(&g_rot, sizeof(float) * 4 * P));
  b.dL_dmeans3D = g_means; b.dL_dmeans2D = g_m2d; b.dL_dopacity = g_op; b.dL_dcolors = g_col;
  b.dL_dscales = g_sc; b.dL_drotations = g_rot;
  const int32_t rc = gab200_backward(&b, stream);
  if (rc < 0) { std::fprintf(stderr, "gab200_backward: %s\n", gab200_status_string(rc)); return 1; }
  CK(cudaStreamSynchronize(stream));
  std::vector<float> h_op(P);
  CK(cudaMemcpy(h_op.data(), g_op, sizeof(float) * P, cudaMemcpyDeviceToHost));
  double s = 0;
  for (float v : h_op) s += v;
  std::printf("backward: sum dL/dopacity = %.6f; %lld library launches so far\n", s, (long long)gab200_launch_count());
  if (argc > 1) {  // inputs + results, raw little-endian: the Python surface must reproduce them
    std::FILE* f = std::fopen(argv[1], "wb");
    if (!f) { std::perror(argv[1]); return 1; }
    const int32_t head[4] = {P, W, H, (int32_t)n2};
    const float cam[2] = {tanx, tany};
    bool ok = std::fwrite(head, sizeof(head), 1, f) == 1 && std::fwrite(cam, sizeof(cam), 1, f) == 1;
    ok = ok && std::fwrite(means.data(), 4, means.size(), f) == means.size() &&
         std::fwrite(scales.data(), 4, scales.size(), f) == scales.size() && std::fwrite(rots.data(), 4, rots.size(), f) == rots.size() &&
         std::fwrite(opac.data(), 4, opac.size(), f) == opac.size() && std::fwrite(rgb.data(), 4, rgb.size(), f) == rgb.size() &&
         std::fwrite(view.data(), 4, 16, f) == 16 && std::fwrite(proj.data(), 4, 16, f) == 16;
    ok = ok && dump(f, out, (size_t)3 * W * H) && dump(f, radii, (size_t)P) && dump(f, g_means, (size_t)3 * P) &&
         dump(f, g_op, (size_t)P) && dump(f, g_col, (size_t)3 * P) && dump(f, g_sc, (size_t)3 * P);
    std::fclose(f);
    if (!ok) { std::fprintf(stderr, "dump failed\n"); return 1; }
  }
  return 0;
}
