// Integer binning pipeline: prefix sum -> duplicate-with-keys -> (tile|depth) radix sort ->
// tile ranges + sorted-instance record gather. Everything here is bit-exact by construction.
//
// Replaces cub InclusiveSum (reference rasterizer_impl.cu:270-273), duplicateWithKeys (:64-101),
// cub SortPairs (:298-309), cudaMemset + identifyTileRanges (:311-316, :106-125).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "gsr_common.cuh"

namespace gsr {

uint32_t higher_msb(uint32_t n) {
  // bisection for (index of the highest set bit) + 1, as the reference's getHigherMsb
  uint32_t msb = sizeof(n) * 4, step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb)
      msb += step;
    else
      msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

size_t scan_temp_bytes(int P) {
  size_t bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, P);
  return bytes;
}

size_t sort_temp_bytes(int R) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, R);
  return bytes;
}

cudaError_t run_scan(const GeomState& g, int P, cudaStream_t s) {
  size_t bytes = g.scan_temp_bytes;
  return cub::DeviceScan::InclusiveSum(g.scan_temp, bytes, g.tiles_touched, g.point_offsets, P, s);
}

cudaError_t run_sort(const BinningState& b, int R, int end_bit, cudaStream_t s) {
  size_t bytes = b.sort_temp_bytes;
  return cub::DeviceRadixSort::SortPairs(b.sort_temp, bytes, b.keys_unsorted, b.keys, b.vals_unsorted, b.vals, R, 0,
                                         end_bit, s);
}

namespace {

// One thread per Gaussian; rows outermost, instances of a Gaussian contiguous from offsets[idx-1]
// (reference rasterizer_impl.cu:90-99). The tile rectangle is recomputed from the stored pixel mean and
// the integer radius with the same float ops as in preprocess (getRect, auxiliary.h:39-46).
__global__ void __launch_bounds__(256) duplicate_with_keys_kernel(int P, const SplatRecord* __restrict__ rec,
                                                                 const uint32_t* __restrict__ offsets,
                                                                 const int* __restrict__ radii,
                                                                 uint64_t* __restrict__ keys,
                                                                 uint32_t* __restrict__ vals, int grid_x, int grid_y) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const int radius = radii[idx];
  if (radius <= 0) return;
  uint32_t off = idx == 0 ? 0u : offsets[idx - 1];
  const float4 q0 = rec[idx].q0;
  const float depth = rec[idx].q2.w;
  const float rf = (float)radius;
  const int rminx = min(grid_x, max(0, __float2int_rz(__fmul_rn(__fadd_rn(q0.x, -rf), 0.0625f))));
  const int rminy = min(grid_y, max(0, __float2int_rz(__fmul_rn(__fadd_rn(q0.y, -rf), 0.0625f))));
  const int rmaxx =
      min(grid_x, max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(q0.x, rf), 16.0f), -1.0f), 0.0625f))));
  const int rmaxy =
      min(grid_y, max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(q0.y, rf), 16.0f), -1.0f), 0.0625f))));
  const uint64_t depth_bits = (uint64_t)__float_as_uint(depth);
  for (int y = rminy; y < rmaxy; ++y) {
    for (int x = rminx; x < rmaxx; ++x) {
      keys[off] = ((uint64_t)(uint32_t)(y * grid_x + x) << 32) | depth_bits;
      vals[off] = (uint32_t)idx;
      ++off;
    }
  }
}

// One thread per sorted instance: tile range boundaries (identifyTileRanges) fused with the gather of
// the Gaussian's 48-byte record into sorted order (so the blend kernels read contiguous batches).
__global__ void __launch_bounds__(256) finalize_binning_kernel(int R, const uint64_t* __restrict__ keys,
                                                              const uint32_t* __restrict__ vals,
                                                              const SplatRecord* __restrict__ rec,
                                                              SplatRecord* __restrict__ inst,
                                                              uint2* __restrict__ ranges) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const uint32_t tile = (uint32_t)(keys[i] >> 32);
  if (i == 0) {
    ranges[tile].x = 0;
  } else {
    const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
    if (tile != prev) {
      ranges[prev].y = (uint32_t)i;
      ranges[tile].x = (uint32_t)i;
    }
  }
  if (i == R - 1) ranges[tile].y = (uint32_t)R;

  const uint32_t id = vals[i];
  const float4* src = reinterpret_cast<const float4*>(rec + id);
  float4* dst = reinterpret_cast<float4*>(inst + i);
  const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
  dst[0] = a, dst[1] = b, dst[2] = c;
}

// Launch order of the blend CTAs: tiles bucketed by floor(log2(list length)) and emitted longest
// bucket first (coarse LPT scheduling; the order inside a bucket is irrelevant). One CTA.
__global__ void __launch_bounds__(1024) order_tiles_kernel(int tiles, const uint2* __restrict__ ranges,
                                                          uint32_t* __restrict__ order,
                                                          uint32_t* __restrict__ tile_last) {
  __shared__ uint32_t s_count[33], s_base[33];
  if (threadIdx.x < 33) s_count[threadIdx.x] = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < tiles; t += blockDim.x) {
    const uint2 r = ranges[t];
    atomicAdd(&s_count[32 - __clz(r.y - r.x)], 1u);
    tile_last[t] = 0;  // blend_fwd accumulates the tile's last contributor with atomicMax
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    for (int b = 32; b >= 0; --b) {
      s_base[b] = acc;
      acc += s_count[b];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < tiles; t += blockDim.x) {
    const uint2 r = ranges[t];
    order[atomicAdd(&s_base[32 - __clz(r.y - r.x)], 1u)] = (uint32_t)t;
  }
}

}  // namespace

cudaError_t launch_order_tiles(int tiles, const uint2* ranges, uint32_t* tile_order, uint32_t* tile_last,
                               cudaStream_t s) {
  order_tiles_kernel<<<1, 1024, 0, s>>>(tiles, ranges, tile_order, tile_last);
  return cudaGetLastError();
}

cudaError_t launch_duplicate_with_keys(int P, const GeomState& g, const int* radii, const BinningState& b, int grid_x,
                                       int grid_y, cudaStream_t s) {
  duplicate_with_keys_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.rec, g.point_offsets, radii, b.keys_unsorted,
                                                            b.vals_unsorted, grid_x, grid_y);
  return cudaGetLastError();
}

cudaError_t launch_finalize_binning(int R, const BinningState& b, const GeomState& g, int grid_x, uint2* ranges,
                                    cudaStream_t s) {
  if (R <= 0) return cudaSuccess;
  finalize_binning_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, b.keys, b.vals, g.rec, b.inst, ranges);
  return cudaGetLastError();
}

}  // namespace gsr
