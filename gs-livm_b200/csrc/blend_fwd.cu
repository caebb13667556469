// Front-to-back alpha blend of C (colour), D (depth) and S (silhouette) per 16x16 tile, sm_100a.
//
// Replaces FORWARD::render / renderCUDA<3> (reference src/cuda_rasterizer/forward.cu:291-441).
//
// Design ("blend_fwd"):
//  * one CTA per tile: 8 consumer warps, each owning an 8x4-pixel sub-tile, plus 1 producer warp;
//  * warp-specialised TMA pipeline: the producer's elected lane stages batches of 256 sorted instances
//    (48-byte records, contiguous thanks to finalize_binning) into a 4-deep shared-memory ring with
//    cp.async.bulk (SASS UBLKCP) completing on per-stage "full" mbarriers; consumer warps release a
//    stage by arriving on its "empty" mbarrier. There is NO CTA-wide barrier in the loop: a warp whose
//    sub-tile is cheap runs ahead of one whose sub-tile is dense (ncu on the first version of this
//    kernel showed 60 % of all stall samples on the per-batch __syncthreads);
//  * per batch every warp first tests 32 records per instruction against its sub-tile with the
//    per-Gaussian conservative extent (preprocess.cu) and __ballot_sync, then walks only the
//    survivors; skipped pairs are exactly pairs the reference `continue`s on, so results (including
//    n_contrib, which indexes the full list) are unchanged;
//  * survivors are processed four at a time: the four alphas (conic form + expf) are independent and
//    overlap in the pipeline, only the cheap transmittance recurrence is applied in order;
//  * CTAs are launched heaviest-tile-first (tile_order, written by finalize_binning) so the long
//    tiles do not end up alone in the tail of the grid.
//
// Numerics: the per-pair operation DAG (which product is fused, order of accumulation) mirrors the
// reference's sm_100a SASS through explicit _rn intrinsics; exp is the
// same precise expf. With identical inputs C, D, S, final_T and n_contrib come out bit-identical.
#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kStages = 4;
constexpr int kConsumerWarps = 8;
constexpr int kThreads = (kConsumerWarps + 1) * 32;
constexpr int kIlp = 4;
constexpr uint32_t kFull = 0xffffffffu;

struct __align__(128) FwdSmem {
  float4 rec[kStages][kBatch * 3];
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint32_t warps_done;
  uint32_t last;
};

__global__ void __launch_bounds__(kThreads) blend_fwd_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order, const SplatRecord* __restrict__ inst,
    int W, int H, int grid_x, const float* __restrict__ bg, float* __restrict__ out_color,
    float* __restrict__ out_depth, float* __restrict__ out_acc, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ tile_last) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int tile = (int)tile_order[blockIdx.x];
  const uint2 range = ranges[tile];
  const int total = (int)(range.y - range.x);
  const int nbatch = (total + kBatch - 1) / kBatch;
  const SplatRecord* src = inst + range.x;
  volatile uint32_t* warps_done = &sm.warps_done;

  if (tid == 0) {
    sm.warps_done = 0;
    sm.last = 0;
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], kConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ------------------------------ producer warp: one elected lane --------------------------------
    if (lane == 0) {
      int issued = 0;
      for (int b = 0; b < nbatch; ++b) {
        const int stage = b % kStages;
        if (b >= kStages) {
          const uint32_t parity = (uint32_t)(b / kStages - 1) & 1u;
          bool stop = false;
          while (!mbar_try_wait(&sm.empty[stage], parity)) {
            if (*warps_done == kConsumerWarps) {
              stop = true;
              break;
            }
          }
          if (stop) break;
        }
        if (*warps_done == kConsumerWarps) break;
        const uint32_t bytes = (uint32_t)min(kBatch, total - b * kBatch) * (uint32_t)sizeof(SplatRecord);
        mbar_expect_tx(&sm.full[stage], bytes);
        bulk_g2s(sm.rec[stage], src + (size_t)b * kBatch, bytes, &sm.full[stage]);
        issued = b + 1;
      }
      // never leave the CTA with a bulk copy still in flight towards its shared memory
      for (int b = max(0, issued - kStages); b < issued; ++b)
        mbar_wait(&sm.full[b % kStages], (uint32_t)(b / kStages) & 1u);
    }
    return;
  }

  // ---------------------------------- consumer warps ---------------------------------------------
  const int tile_x = tile % grid_x, tile_y = tile / grid_x;
  const int wx = tile_x * kTile + (warp & 1) * 8, wy = tile_y * kTile + (warp >> 1) * 4;
  const int px = wx + (lane & 7), py = wy + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const float wx0 = (float)wx, wx1 = (float)(wx + 7), wy0 = (float)wy, wy1 = (float)(wy + 3);

  float T = 1.0f, A = 0.f, D = 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  uint32_t last_contributor = 0;
  bool done = !inside;
  bool warp_done = __all_sync(kFull, done);
  if (warp_done && lane == 0) atomicAdd(&sm.warps_done, 1u);

  for (int b = 0; b < nbatch; ++b) {
    const int stage = b % kStages;
    {
      const uint32_t parity = (uint32_t)(b / kStages) & 1u;
      bool stop = false;
      while (!mbar_try_wait(&sm.full[stage], parity)) {
        if (*warps_done == kConsumerWarps) {
          stop = true;
          break;
        }
      }
      if (stop || *warps_done == kConsumerWarps) break;
    }
    if (!warp_done) {
      const int cnt = min(kBatch, total - b * kBatch);
      const float4* s = sm.rec[stage];
      const uint32_t base = (uint32_t)(b * kBatch + 1);
      for (int k = 0; k < cnt; k += 32) {
        const int j = k + lane;
        bool hit = false;
        if (j < cnt) {
          const float4 c = s[3 * j];  // {x, y, extent_x, extent_y}
          hit = (c.x + c.z >= wx0) && (c.x - c.z <= wx1) && (c.y + c.w >= wy0) && (c.y - c.w <= wy1);
        }
        uint32_t m = __ballot_sync(kFull, hit);
        while (m) {
          // up to kIlp survivors at a time: alphas first (independent), then the ordered recurrence
          int jj[kIlp];
          float alpha[kIlp];
#pragma unroll
          for (int u = 0; u < kIlp; ++u) {
            const bool valid = m != 0;
            jj[u] = valid ? k + __ffs(m) - 1 : jj[0];
            m &= m - 1;
            const float4 q0 = s[3 * jj[u]];
            const float4 q1 = s[3 * jj[u] + 1];  // {conic.x, conic.y, conic.z, opacity}
            const float dx = __fadd_rn(q0.x, -pxf);
            const float dy = __fadd_rn(q0.y, -pyf);
            const float quad = __fmaf_rn(dx, __fmul_rn(dx, q1.x), __fmul_rn(dy, __fmul_rn(dy, q1.z)));
            const float power = __fmaf_rn(quad, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, q1.y)));
            const float a = fminf(0.99f, __fmul_rn(q1.w, expf(power)));
            // reference: `if (power > 0) continue; ... if (alpha < 1/255) continue;`
            alpha[u] = (valid && !(power > 0.0f) && !(a < 1.0f / 255.0f)) ? a : -1.0f;
          }
#pragma unroll
          for (int u = 0; u < kIlp; ++u) {
            if (alpha[u] >= 0.0f && !done) {
              const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -alpha[u]));
              if (test_T < 0.0001f) {
                done = true;
              } else {
                const float4 q2 = s[3 * jj[u] + 2];  // {r, g, b, depth}
                A = __fmaf_rn(T, alpha[u], A);
                D = __fmaf_rn(T, __fmul_rn(alpha[u], q2.w), D);
                C0 = __fmaf_rn(T, __fmul_rn(alpha[u], q2.x), C0);
                C1 = __fmaf_rn(T, __fmul_rn(alpha[u], q2.y), C1);
                C2 = __fmaf_rn(T, __fmul_rn(alpha[u], q2.z), C2);
                T = test_T;
                last_contributor = base + (uint32_t)jj[u];
              }
            }
          }
        }
        if (__all_sync(kFull, done)) {
          warp_done = true;
          if (lane == 0) atomicAdd(&sm.warps_done, 1u);
          break;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[stage]);
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
  if (lane == 0 && wmax) atomicMax(&sm.last, wmax);
  // consumer-only barrier (the producer warp may already have exited)
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
  if (tid == 0) tile_last[tile] = sm.last;
}

}  // namespace

cudaError_t launch_blend_fwd(int width, int height, const ImageState& img, const BinningState& b,
                             const float* background, float* out_color, float* out_depth, float* out_acc,
                             cudaStream_t s) {
  const int grid_x = (width + kTile - 1) / kTile, grid_y = (height + kTile - 1) / kTile;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(blend_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(FwdSmem));
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  blend_fwd_kernel<<<grid_x * grid_y, kThreads, sizeof(FwdSmem), s>>>(img.ranges, img.tile_order, b.inst, width,
                                                                     height, grid_x, background, out_color, out_depth,
                                                                     out_acc, img.final_T, img.n_contrib,
                                                                     img.tile_last);
  return cudaGetLastError();
}

}  // namespace gsr
