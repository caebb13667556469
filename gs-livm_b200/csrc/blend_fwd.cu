// Front-to-back alpha blend of C (colour), D (depth) and S (silhouette) per 16x16 tile, sm_100a.
//
// Replaces FORWARD::render / renderCUDA<3> (reference src/cuda_rasterizer/forward.cu:291-441).
//
// Design ("blend_fwd"):
//  * one CTA (256 threads, 8 warps) per tile; each warp owns an 8x4-pixel sub-tile;
//  * batches of 256 sorted instances (48-byte records, contiguous thanks to finalize_binning) are
//    staged into shared memory by ONE elected thread with a TMA bulk copy (cp.async.bulk ->
//    SASS UBLKCP) completing on an mbarrier; two stages, so the copy of batch b+1/b+2 overlaps the
//    math of batch b and no thread spends registers or LSU slots on staging;
//  * per batch every warp first tests 32 records per instruction against its sub-tile with the
//    per-Gaussian conservative extent (preprocess.cu) and __ballot_sync, then walks only the
//    survivors; a record that cannot reach alpha >= 1/255 anywhere in the sub-tile costs 1/32 of
//    a lane-iteration instead of a full one. Skipped pairs are exactly pairs the reference `continue`s
//    on, so results are unchanged (including n_contrib, which indexes the full list);
//  * warp-level early termination by __all_sync(done), CTA-level by __syncthreads_or.
//
// Numerics: the per-pair operation DAG (which product is fused, order of accumulation) mirrors the
// reference's sm_100a SASS through explicit _rn intrinsics; exp is the
// same precise expf. With identical inputs C, D, S, final_T and n_contrib come out bit-identical.
#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kStages = 2;
constexpr uint32_t kFull = 0xffffffffu;

__global__ void __launch_bounds__(kTilePixels) blend_fwd_kernel(
    const uint2* __restrict__ ranges, const SplatRecord* __restrict__ inst, int W, int H, int grid_x,
    const float* __restrict__ bg, float* __restrict__ out_color, float* __restrict__ out_depth,
    float* __restrict__ out_acc, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
    uint32_t* __restrict__ tile_last) {
  __shared__ __align__(128) float4 s_rec[kStages][kBatch * 3];
  __shared__ __align__(8) uint64_t s_full[kStages];
  __shared__ uint32_t s_last;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int tile_x = tile % grid_x, tile_y = tile / grid_x;
  const int wx = tile_x * kTile + (warp & 1) * 8, wy = tile_y * kTile + (warp >> 1) * 4;
  const int px = wx + (lane & 7), py = wy + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const float wx0 = (float)wx, wx1 = (float)(wx + 7), wy0 = (float)wy, wy1 = (float)(wy + 3);

  const uint2 range = ranges[tile];
  const int total = (int)(range.y - range.x);
  const int nbatch = (total + kBatch - 1) / kBatch;
  const SplatRecord* src = inst + range.x;

  if (tid == 0) {
    s_last = 0;
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int b = 0; b < min(nbatch, kStages); ++b) {
      const uint32_t bytes = (uint32_t)min(kBatch, total - b * kBatch) * (uint32_t)sizeof(SplatRecord);
      mbar_expect_tx(&s_full[b], bytes);
      bulk_g2s(s_rec[b], src + (size_t)b * kBatch, bytes, &s_full[b]);
    }
  }

  float T = 1.0f, A = 0.f, D = 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  uint32_t last_contributor = 0;
  bool done = !inside;
  bool warp_done = __all_sync(kFull, done);

  for (int b = 0; b < nbatch; ++b) {
    const int stage = b & 1;
    mbar_wait(&s_full[stage], (uint32_t)(b >> 1) & 1u);
    const int cnt = min(kBatch, total - b * kBatch);
    const float4* s = s_rec[stage];

    if (!warp_done) {
      for (int k = 0; k < cnt; k += 32) {
        const int j = k + lane;
        bool hit = false;
        if (j < cnt) {
          const float4 c = s[3 * j];  // {x, y, extent_x, extent_y}
          hit = (c.x + c.z >= wx0) && (c.x - c.z <= wx1) && (c.y + c.w >= wy0) && (c.y - c.w <= wy1);
        }
        uint32_t m = __ballot_sync(kFull, hit);
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          const int jj = k + bit;
          if (!done) {
            const float4 q0 = s[3 * jj];
            const float4 q1 = s[3 * jj + 1];  // {conic.x, conic.y, conic.z, opacity}
            const float dx = __fadd_rn(q0.x, -pxf);
            const float dy = __fadd_rn(q0.y, -pyf);
            const float quad = __fmaf_rn(dx, __fmul_rn(dx, q1.x), __fmul_rn(dy, __fmul_rn(dy, q1.z)));
            const float power = __fmaf_rn(quad, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, q1.y)));
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, __fmul_rn(q1.w, expf(power)));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -alpha));
            if (test_T < 0.0001f) {
              done = true;
              continue;
            }
            const float4 q2 = s[3 * jj + 2];  // {r, g, b, depth}
            A = __fmaf_rn(T, alpha, A);
            D = __fmaf_rn(T, __fmul_rn(alpha, q2.w), D);
            C0 = __fmaf_rn(T, __fmul_rn(alpha, q2.x), C0);
            C1 = __fmaf_rn(T, __fmul_rn(alpha, q2.y), C1);
            C2 = __fmaf_rn(T, __fmul_rn(alpha, q2.z), C2);
            T = test_T;
            last_contributor = (uint32_t)(b * kBatch + jj + 1);
          }
        }
        if (__all_sync(kFull, done)) {
          warp_done = true;
          break;
        }
      }
    }

    // every warp is finished with this stage -> it may be refilled; also the CTA-level early-out
    const int alive = __syncthreads_or(!warp_done);
    if (!alive) {
      if (tid == 0 && b + 1 < nbatch) mbar_wait(&s_full[(b + 1) & 1], (uint32_t)((b + 1) >> 1) & 1u);
      break;
    }
    if (tid == 0 && b + kStages < nbatch) {
      const int nb = b + kStages;
      const uint32_t bytes = (uint32_t)min(kBatch, total - nb * kBatch) * (uint32_t)sizeof(SplatRecord);
      mbar_expect_tx(&s_full[stage], bytes);
      bulk_g2s(s_rec[stage], src + (size_t)nb * kBatch, bytes, &s_full[stage]);
    }
  }

  if (inside) {
    const size_t pix = (size_t)W * py + px;
    const size_t plane = (size_t)W * H;
    final_T[pix] = T;
    n_contrib[pix] = last_contributor;
    out_color[pix] = __fmaf_rn(__ldg(bg + 0), T, C0);
    out_color[plane + pix] = __fmaf_rn(__ldg(bg + 1), T, C1);
    out_color[2 * plane + pix] = __fmaf_rn(__ldg(bg + 2), T, C2);
    out_depth[pix] = D;
    out_acc[pix] = A;
  }
  const uint32_t wmax = __reduce_max_sync(kFull, last_contributor);
  if (lane == 0 && wmax) atomicMax(&s_last, wmax);
  __syncthreads();
  if (tid == 0) tile_last[tile] = s_last;
}

}  // namespace

cudaError_t launch_blend_fwd(int width, int height, const ImageState& img, const BinningState& b,
                             const float* background, float* out_color, float* out_depth, float* out_acc,
                             cudaStream_t s) {
  const int grid_x = (width + kTile - 1) / kTile, grid_y = (height + kTile - 1) / kTile;
  blend_fwd_kernel<<<grid_x * grid_y, kTilePixels, 0, s>>>(img.ranges, b.inst, width, height, grid_x, background,
                                                          out_color, out_depth, out_acc, img.final_T, img.n_contrib,
                                                          img.tile_last);
  return cudaGetLastError();
}

}  // namespace gsr
