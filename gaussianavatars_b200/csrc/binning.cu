// binning.cu -- the library primitives of the path (north_star: cub::DeviceRadixSort for the tile|depth keys).
//
// The reference sorts N (tile << 32 | depth) 64-bit keys in one LSD radix sort.  An LSD radix sort
// is a sequence of stable passes from the low digits to the high digits, and the low 32 bits (the depth) are the
// same for every instance of a splat.  So the low-digit passes are done ONCE PER SPLAT, before duplication:
//   stage A  cub::DeviceRadixSort::SortPairs (u32 depth bits, u32 splat id) over the P splats          [4 passes x 8 B x P]
//   scan     cub::DeviceScan::InclusiveSum of tiles_touched in depth order -> emission offsets
//   emit     instances are written in depth order with key = tile id only
//   stage B  cub::DeviceRadixSort::SortPairs (u32 tile id, u32 splat id), bits [0, bits(tiles))        [2 passes x 8 B x N]
// Stable passes compose: the result is exactly the reference's order (ties: ascending splat id), tested bit for bit,
// while the N-sized traffic drops from 6 passes x 12 B to 2 passes x 8 B.
//
// Stage A + scan at P ~ 1e5 are latency, not bandwidth: cub's onesweep is 6 launches of ~12 CTAs chained by
// decoupled look-back, the scan 2 more (77 us together at 100k splats).  When the caller passes the previous frame's
// depth-key range (gab200_forward_args.depth_hint_*), stage A + scan run as a BUCKET SORT instead (preprocess.cu):
//   preprocess_kernel    bucket = floor((key - lo) * nb / (hi - lo + 1)) clamped to [0, nb) -- monotonic in the key;
//                        atomicAdd on the bucket's splat counter returns the splat's arrival rank; a second counter
//                        sums the bucket's instances.  (32..64 splats per bucket, nb a power of two <= 8192.)
//   depth_scatter_kernel every CTA rebuilds the exclusive prefix of the nb counters in shared memory and drops
//                        (key, id) at start[bucket] + rank; CTA 0 publishes the prefixes, N and M.
//   depth_bucket_kernel  one CTA per bucket ranks its (key << 32 | id) pairs by counting in shared memory and writes
//                        ids in (key, id) order -- exactly what a stable sort by key of ids 0..P-1 gives -- plus the
//                        running instance counts (the scan), offset by the bucket's base.
// Keys outside [lo, hi] share the end buckets (order still right); a bucket that outgrows shared memory (2048 splats)
// raises a flag that the host reads with N, and the frame is redone with stage A + scan above.  Same output either
// way, bit for bit (tests/test_gpu_depth_sort.py).
#include <cub/cub.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "common.cuh"
#include "kernels.cuh"

namespace gab {

struct GatherTiles {
  const uint32_t* tiles;
  __host__ __device__ __forceinline__ uint32_t operator()(const uint32_t& id) const { return tiles[id]; }
};
using GatherIt = cub::TransformInputIterator<uint32_t, GatherTiles, const uint32_t*>;

size_t scan_temp_bytes(int P) {
  size_t bytes = 0;
  GatherIt it((const uint32_t*)nullptr, GatherTiles{nullptr});
  cub::DeviceScan::InclusiveSum(nullptr, bytes, it, (uint32_t*)nullptr, P);
  return bytes;
}

// offsets[j] = sum_{k<=j} tiles_touched[order[k]]
cudaError_t run_scan(void* temp, size_t temp_bytes, const uint32_t* order, const uint32_t* tiles_touched, uint32_t* out,
                     int P, cudaStream_t stream) {
  count_launch();
  GatherIt it(order, GatherTiles{tiles_touched});
  return cub::DeviceScan::InclusiveSum(temp, temp_bytes, it, out, P, stream);
}

size_t sort_temp_bytes(int64_t N, int end_bit) {
  size_t bytes = 0;
  cub::DoubleBuffer<uint32_t> k(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, N, 0, end_bit);
  return bytes;
}

cudaError_t run_sort(void* temp, size_t temp_bytes, uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a,
                     uint32_t* vals_b, int64_t N, int end_bit, int* selector_out, cudaStream_t stream) {
  cub::DoubleBuffer<uint32_t> k(keys_a, keys_b);
  cub::DoubleBuffer<uint32_t> v(vals_a, vals_b);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k, v, N, 0, end_bit, stream);
  // onesweep: 1 histogram + 1 scan + one pass per 8-bit digit
  for (int i = 0; i < 2 + (end_bit + 7) / 8; i++) count_launch();
  *selector_out = k.selector;
  return e;
}

}  // namespace gab
