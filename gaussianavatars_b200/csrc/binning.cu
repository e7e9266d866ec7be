// binning.cu -- the two library primitives of the path (north_star: cub::DeviceRadixSort for the tile|depth keys).
//   K2: cub::DeviceScan::InclusiveSum over tiles_touched          
//   K4: cub::DeviceRadixSort::SortPairs on (u64 key, u32 splat id), bits [0, 32 + bits(tile id))   (K4)
// The sort is stable, so equal (tile, depth-bits) keys keep emission order = ascending splat id: the sorted stream is
// bit-identical to the reference's for identical keys.
#include <cub/cub.cuh>

#include "common.cuh"
#include "kernels.cuh"

namespace gab {

size_t scan_temp_bytes(int P) {
  size_t bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, P);
  return bytes;
}

cudaError_t run_scan(void* temp, size_t temp_bytes, const uint32_t* in, uint32_t* out, int P, cudaStream_t stream) {
  count_launch();
  return cub::DeviceScan::InclusiveSum(temp, temp_bytes, in, out, P, stream);
}

size_t sort_temp_bytes(int64_t N, int end_bit) {
  size_t bytes = 0;
  cub::DoubleBuffer<uint64_t> k(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, N, 0, end_bit);
  return bytes;
}

cudaError_t run_sort(void* temp, size_t temp_bytes, uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a,
                     uint32_t* vals_b, int64_t N, int end_bit, int* selector_out, cudaStream_t stream) {
  cub::DoubleBuffer<uint64_t> k(keys_a, keys_b);
  cub::DoubleBuffer<uint32_t> v(vals_a, vals_b);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k, v, N, 0, end_bit, stream);
  // onesweep: 1 histogram + 1 scan + one pass per 8-bit digit
  for (int i = 0; i < 2 + (end_bit + 7) / 8; i++) count_launch();
  *selector_out = k.selector;
  return e;
}

}  // namespace gab
