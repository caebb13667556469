// Shared declarations of libgsr_b200: scratch-state layout, launch prototypes, small PTX helpers.
//
// The three byte buffers (geometry / image / binning) play the role of the reference's
// GeometryState / ImageState / BinningState (include/gs/cuda_rasterizer/rasterizer_impl.h:20-67,
// src/cuda_rasterizer/rasterizer_impl.cu:137-177), but the layout inside them is B200-first:
// per-Gaussian data needed by the blend is one 48-byte record (three float4) so that the blend
// kernels can stage contiguous batches of sorted instances with one TMA bulk copy per batch.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "gsr.h"

namespace gsr {

constexpr int kTile = 16;           // BLOCK_X == BLOCK_Y (config.h:16-17)
constexpr int kTilePixels = 256;    // threads per blend CTA
constexpr int kBatch = 256;         // sorted instances staged per TMA bulk copy
constexpr int kAlign = 128;         // sub-array alignment, same as the reference's obtain(..., 128)

// One Gaussian as the blend sees it (48 B). Written by preprocess per Gaussian, then gathered into
// sorted-instance order by finalize_binning.
//   q0 = { mean2D.x, mean2D.y, extent.x, extent.y }   extent = conservative half-size of {alpha >= 1/255}
//   q1 = { conic.x, conic.y, conic.z, opacity }       (reference conic_opacity, forward.cu:284)
//   q2 = { r, g, b, depth }                           (reference rgb + depths, forward.cu:274-280)
struct alignas(16) SplatRecord {
  float4 q0, q1, q2;
};
static_assert(sizeof(SplatRecord) == 48, "record must be 3 x 16 B for bulk copies");

struct GeomState {
  SplatRecord* rec;         // [P]
  uint32_t* tiles_touched;  // [P]
  uint32_t* point_offsets;  // [P] inclusive prefix sum of tiles_touched
  int* internal_radii;      // [P] only used when the caller passes radii == nullptr
  uint8_t* clamped;         // [P] bit c set = colour channel c was clamped at 0 (forward.cu:72-74)
  char* scan_temp;          // CUB DeviceScan temp
  size_t scan_temp_bytes;
};

struct ImageState {
  float* final_T;       // [W*H]   reference accum_alpha
  uint32_t* n_contrib;  // [W*H]
  uint2* ranges;        // [tiles]
  uint32_t* tile_last;  // [tiles] max n_contrib in the tile: where the backward starts
  uint32_t* tile_order; // [tiles] tile ids, longest instance lists first (launch order of the blend CTAs)
  uint32_t* tile_count; // [tiles] instances per tile (preprocess), then the tile's scatter cursor (tile_sort.cu)
  uint32_t* bucket_base; // [tiles] first depth-bucket descriptor of the tile's list, if it is a long one (tile_sort.cu)
  uint32_t* summary;    // [12] {R, longest list, lists longer than split_len, overflow code (sync-free forward),
                        //       depth buckets of the long lists, -, then two 64-bit work counters of the blend forward:
                        //       [6,7] (pixel, candidate) pairs inside the candidates' boxes = alphas evaluated,
                        //       [8,9] pairs blended = what the backward visits; [10,11] -}
  uint32_t* block_count;  // [tiles*8] candidates per 8x4 pixel block (block_lists.cu)
  uint4* bdesc;           // [tiles*8] by (position in tile_order, block): {range start, list length, candidates, tile} --
                          //           one load tells a blend warp everything it needs to find its list
  uint32_t* last_group;   // [W*H] 1 + the group (of 32 candidates of its block's list) of the pixel's last contributor
  uint32_t* tail_hits;    // [tiles*8][32] hit masks of each block's last, partial group (see BinningState::bhits)
};

// tile_sort.cu: a segment (a whole short list, or one depth bucket of a long list) of up to kSegCap instances is
// sorted by a 256-thread CTA in shared memory. Lists longer than Tuning::split_len (2^k - 1 <= kSegCap - 1: the class
// boundary then coincides with tile_order's buckets) are first cut into depth buckets of ~Tuning::bucket_target
// instances, at most kMaxBuckets per list. No list-length ceiling: work is linear in the length.
constexpr int kSegCap = 2048;
constexpr int kMaxBuckets = 256;
constexpr int kMinBucketTarget = 256;
// depth buckets a list of n instances is cut into (0: sorted whole); shared by tile_scan and partition_long
__host__ __device__ inline uint32_t bucket_count(uint32_t n, uint32_t split_len, uint32_t target) {
  if (n <= split_len) return 0u;
  const uint32_t k = (n + target - 1u) / target;
  return k < (uint32_t)kMaxBuckets ? k : (uint32_t)kMaxBuckets;
}
// upper bound of the number of depth buckets of a frame with R instances: a long list has more than
// split_len >= kMinBucketTarget - 1 instances and at most n / target + 1 <= 2 n / kMinBucketTarget buckets
inline size_t max_depth_buckets(size_t R) { return 2 * R / kMinBucketTarget + 1; }

struct BinningState {
  uint64_t* keys_unsorted;  // [R]
  uint32_t* vals_unsorted;  // [R]
  uint64_t* keys;           // [R] sorted
  uint32_t* vals;           // [R] sorted == reference point_list
  SplatRecord* inst;        // [R] records in sorted order
  uint2* blist;             // [8R] per-block candidate lists {index in the tile list, coverage mask}: block b of a
                            //      tile with range [r, r+n) owns entries 8r + b*n .. 8r + (b+1)*n (block_lists.cu)
  uint32_t* bhits;          // [8R] same indexing: word 32g + p of a block = which candidates of the block's group g
                            //      were blended into pixel p (bit 8*row + col order) by the forward; the backward walks
                            //      exactly those pairs. A last group of fewer than 32 candidates lives in tail_hits.
  uint32_t* vals_tmp;       // [R] ids of the long lists after the depth partition (their depths: second half of
                            //     keys_unsorted, which holds 32-bit depth words on the per-tile path)
  uint4* seg_desc;          // [max_depth_buckets(R)] {first instance, instances, tile, -} per depth bucket
  char* sort_temp;
  size_t sort_temp_bytes;
};

template <typename T>
inline void carve(char*& chunk, T*& ptr, size_t count) {
  uintptr_t p = (reinterpret_cast<uintptr_t>(chunk) + kAlign - 1) & ~uintptr_t(kAlign - 1);
  ptr = reinterpret_cast<T*>(p);
  chunk = reinterpret_cast<char*>(ptr + count);
}

size_t scan_temp_bytes(int P);
size_t sort_temp_bytes(int R);
GeomState carve_geom(char* chunk, int P);
ImageState carve_image(char* chunk, int width, int height);
BinningState carve_binning(char* chunk, int R);

struct Camera {
  float view[16];
  float proj[16];
  float campos[3];
};

// ---- kernel launchers (each returns the cudaError_t of the launch) ----------------------------
cudaError_t launch_preprocess_fwd(const gsr_fwd_args& a, const GeomState& g, int* radii, uint32_t* tile_count,
                                  float focal_x, float focal_y, cudaStream_t s);
cudaError_t launch_mark_visible(int P, const float* means3D, const float* viewmatrix, unsigned char* present,
                                cudaStream_t s);
cudaError_t run_scan(const GeomState& g, int P, cudaStream_t s);
cudaError_t launch_duplicate_with_keys(int P, const GeomState& g, const int* radii, const BinningState& b, int grid_x,
                                       int grid_y, cudaStream_t s);
cudaError_t run_sort(const BinningState& b, int R, int end_bit, cudaStream_t s);
cudaError_t launch_finalize_binning(int R, const BinningState& b, const GeomState& g, int grid_x, uint2* ranges,
                                    cudaStream_t s);
cudaError_t launch_tile_scan(int tiles, const ImageState& img, uint32_t r_cap, cudaStream_t s);
cudaError_t launch_scatter_instances(int P, const GeomState& g, const int* radii, const ImageState& img,
                                     const BinningState& b, int grid_x, int grid_y, cudaStream_t s);
// n_long / n_buckets < 0: not known on the host (sync-free forward); r_cap = instances the binning state is carved for
cudaError_t launch_tile_sort(int tiles, int n_long, int n_buckets, int r_cap, const ImageState& img,
                             const BinningState& b, const GeomState& g, int grid_x, cudaStream_t s);
cudaError_t launch_order_tiles(int tiles, const uint2* ranges, uint32_t* tile_order, uint32_t* tile_last,
                               cudaStream_t s);
cudaError_t launch_block_lists(int tiles, const ImageState& img, const BinningState& b, int grid_x, cudaStream_t s);
cudaError_t launch_blend_fwd_sparse(int width, int height, const ImageState& img, const BinningState& b,
                                    const float* background, float* out_color, float* out_depth, float* out_acc,
                                    cudaStream_t s);
cudaError_t launch_blend_bwd_sparse(int width, int height, const ImageState& img, const BinningState& b,
                                    const float* background, const float* dL_dpix, const float* dL_dacc,
                                    float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                                    cudaStream_t s);
cudaError_t launch_geom_bwd(const gsr_bwd_args& a, const GeomState& g, const int* radii, float focal_x, float focal_y,
                            cudaStream_t s);

uint32_t higher_msb(uint32_t n);  // reference getHigherMsb, rasterizer_impl.cu:35-48

// ---- process-wide tuning / A-B knobs (gsr_configure; initial values from the environment, read ONCE) ----------
struct Tuning {
  int sort_device;    // 1 = always the device-wide binning path (the reference's sequence)     GSR_SORT=device
  int heavy_fwd;      // list length from which a tile's blend warps take the FINE shape         GSR_BLEND_HEAVY
  int heavy_bwd;
  int fwd_ilp;        // pairs evaluated per lane per step, COARSE shape                         GSR_FWD_ILP
  int bwd_ilp;        //                                                                         GSR_BWD_ILP
  int split_len;      // lists longer than this are cut into depth buckets (2^k - 1)             GSR_SORT_SPLIT
  int bucket_target;  // instances per depth bucket                                              GSR_SORT_BUCKET
  int ar_grid;        // CTAs per SM of the window exchange kernels                              GSR_AR_GRID
};
Tuning tuning();  // snapshot (each field is read atomically)

// per-device caches behind one mutex (forward on the mapping thread, backward on an autograd worker: SURVEY 8b)
int device_sm_count();                                            // of the current device
cudaError_t ensure_dynamic_smem(const void* kernel, int bytes);   // cudaFuncSetAttribute once per (kernel, device)

// ---- device helpers ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// torch::nn::functional::normalize(q, dim = 1): q / max(|q|_2, 1e-12) (GaussianModel::Get_rotation, gaussian.cuh:47-49);
// n_out = the denominator (the backward needs it)
__device__ __forceinline__ float4 normalize_quat(const float4 q, float* n_out = nullptr) {
  const float n = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y)),
                                             __fadd_rn(__fmul_rn(q.z, q.z), __fmul_rn(q.w, q.w)))), 1e-12f);
  if (n_out) *n_out = n;
  return make_float4(__fdiv_rn(q.x, n), __fdiv_rn(q.y, n), __fdiv_rn(q.z, n), __fdiv_rn(q.w, n));
}
// GaussianModel::Get_opacity = torch::sigmoid (gaussian.cuh:44-46), as LibTorch's kernel computes it
__device__ __forceinline__ float sigmoid_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

#endif  // __CUDACC__

}  // namespace gsr
