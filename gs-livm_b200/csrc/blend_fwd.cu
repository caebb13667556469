// Front-to-back alpha blend of C (colour), D (depth) and S (silhouette) per 16x16 tile, sm_100a.
//
// Replaces FORWARD::render / renderCUDA<3> (reference src/cuda_rasterizer/forward.cu:291-441).
//
// Design ("blend_fwd"):
//  * one CTA per tile, 8 warps, each warp owns an 8x4-pixel sub-tile and is FULLY independent of the
//    other seven: it stages its own batches of 64 sorted instances (48-byte records, contiguous thanks
//    to finalize_binning) into a private 3-deep shared-memory ring with TMA bulk copies
//    (cp.async.bulk -> SASS UBLKCP) that complete on private mbarriers. There is no CTA-wide barrier
//    and no shared ring: ncu on the first two versions of this kernel (profiles/) showed 60 % of all
//    stall samples on the per-batch __syncthreads, then on the shared ring's full/empty waits, because
//    the sub-tiles of a tile are busy at different depths of the list. The price is that a record is
//    fetched from L2 by each of the 8 warps; the instance array is L2-resident (48 B x R);
//  * per batch the warp tests 32 records per instruction against its sub-tile with the per-Gaussian
//    conservative extent (preprocess.cu), compacts the survivors' indices with __ballot_sync/__popc
//    into a small per-warp list, and only walks those. Skipped pairs are exactly pairs the reference
//    `continue`s on, so results (including n_contrib, which indexes the full list) are unchanged;
//  * survivors are processed four at a time: the four alphas (conic form + expf) are independent and
//    overlap in the pipeline, only the cheap transmittance recurrence is applied in order;
//  * a warp whose 32 pixels are saturated stops fetching and exits;
//  * CTAs are launched heaviest-tile-first (tile_order, written next to finalize_binning).
//
// Numerics: the per-pair operation DAG (which product is fused, order of accumulation) mirrors the
// reference's sm_100a SASS through explicit _rn intrinsics; exp is the
// same precise expf. With identical inputs C, D, S, final_T and n_contrib come out bit-identical.
#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kStages = 3;
constexpr int kWarps = 8;
constexpr int kWB = 64;  // instances per per-warp batch
constexpr int kIlp = 4;
constexpr uint32_t kFull = 0xffffffffu;

struct __align__(128) WarpRing {
  float4 rec[kStages][kWB * 3];
  uint64_t full[kStages];
  uint32_t list[kWB / 4];  // compacted survivor indices, one byte each
};

__global__ void __launch_bounds__(kWarps * 32) blend_fwd_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order, const SplatRecord* __restrict__ inst,
    int W, int H, int grid_x, const float* __restrict__ bg, float* __restrict__ out_color,
    float* __restrict__ out_depth, float* __restrict__ out_acc, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ tile_last) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpRing& ring = reinterpret_cast<WarpRing*>(smem_raw)[warp];

  const int tile = (int)tile_order[blockIdx.x];
  const uint2 range = ranges[tile];
  const int total = (int)(range.y - range.x);
  const int nbatch = (total + kWB - 1) / kWB;
  const SplatRecord* src = inst + range.x;

  const int tile_x = tile % grid_x, tile_y = tile / grid_x;
  const int wx = tile_x * kTile + (warp & 1) * 8, wy = tile_y * kTile + (warp >> 1) * 4;
  const int px = wx + (lane & 7), py = wy + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const float wx0 = (float)wx, wx1 = (float)(wx + 7), wy0 = (float)wy, wy1 = (float)(wy + 3);

  float T = 1.0f, A = 0.f, D = 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  uint32_t last_contributor = 0;
  bool done = !inside;

  if (!__all_sync(kFull, done) && nbatch > 0) {
    auto issue = [&](int b) {  // lane 0 only
      const int stage = b % kStages;
      const uint32_t bytes = (uint32_t)min(kWB, total - b * kWB) * (uint32_t)sizeof(SplatRecord);
      mbar_expect_tx(&ring.full[stage], bytes);
      bulk_g2s(ring.rec[stage], src + (size_t)b * kWB, bytes, &ring.full[stage]);
    };
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) mbar_init(&ring.full[s], 1);
      mbar_fence_init();
      for (int b = 0; b < min(nbatch, kStages); ++b) issue(b);
    }
    __syncwarp();

    int issued = min(nbatch, kStages);
    int b = 0;
    for (; b < nbatch; ++b) {
      const int stage = b % kStages;
      mbar_wait(&ring.full[stage], (uint32_t)(b / kStages) & 1u);
      const int cnt = min(kWB, total - b * kWB);
      const float4* s = ring.rec[stage];
      const uint32_t base = (uint32_t)(b * kWB + 1);

      // ---- cull 32 records per instruction against this warp's sub-tile, compact survivors ----
      int nhit = 0;
      uint8_t* list8 = reinterpret_cast<uint8_t*>(ring.list);
#pragma unroll
      for (int k = 0; k < kWB; k += 32) {
        const int j = k + lane;
        bool hit = false;
        if (j < cnt) {
          const float4 c = s[3 * j];  // {x, y, extent_x, extent_y}
          hit = (c.x + c.z >= wx0) && (c.x - c.z <= wx1) && (c.y + c.w >= wy0) && (c.y - c.w <= wy1);
        }
        const uint32_t m = __ballot_sync(kFull, hit);
        if (hit) list8[nhit + __popc(m & ((1u << lane) - 1u))] = (uint8_t)j;
        nhit += __popc(m);
      }
      __syncwarp();

      // ---- walk the survivors kIlp at a time: alphas first (independent), then the ordered recurrence ----
      for (int g = 0; g < nhit; g += kIlp) {
        const uint32_t packed = ring.list[g >> 2];
        int jj[kIlp];
        float alpha[kIlp];
        float4 feat[kIlp];
        uint32_t any[kIlp];
#pragma unroll
        for (int u = 0; u < kIlp; ++u) {
          const bool valid = g + u < nhit;
          jj[u] = valid ? (int)((packed >> (8 * u)) & 0xffu) : (int)(packed & 0xffu);
          const float4 q0 = s[3 * jj[u]];
          const float4 q1 = s[3 * jj[u] + 1];  // {conic.x, conic.y, conic.z, opacity}
          feat[u] = s[3 * jj[u] + 2];          // {r, g, b, depth}
          const float dx = __fadd_rn(q0.x, -pxf);
          const float dy = __fadd_rn(q0.y, -pyf);
          const float quad = __fmaf_rn(dx, __fmul_rn(dx, q1.x), __fmul_rn(dy, __fmul_rn(dy, q1.z)));
          const float power = __fmaf_rn(quad, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, q1.y)));
          const float a = fminf(0.99f, __fmul_rn(q1.w, expf(power)));
          // reference: `if (power > 0) continue; ... if (alpha < 1/255) continue;`
          const bool on = valid && !done && !(power > 0.0f) && !(a < 1.0f / 255.0f);
          alpha[u] = on ? a : -1.0f;
          any[u] = __ballot_sync(kFull, on);
        }
        // ordered transmittance recurrence, branch-free per lane; a slot no lane needs is skipped warp-wide
#pragma unroll
        for (int u = 0; u < kIlp; ++u) {
          if (any[u] == 0) continue;
          const float a = alpha[u];
          const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -a));
          const bool live = (a >= 0.0f) && !done;
          const bool kill = live && (test_T < 0.0001f);  // reference: done = true; continue
          const bool acc = live && !kill;
          done = done || kill;
          A = acc ? __fmaf_rn(T, a, A) : A;
          D = acc ? __fmaf_rn(T, __fmul_rn(a, feat[u].w), D) : D;
          C0 = acc ? __fmaf_rn(T, __fmul_rn(a, feat[u].x), C0) : C0;
          C1 = acc ? __fmaf_rn(T, __fmul_rn(a, feat[u].y), C1) : C1;
          C2 = acc ? __fmaf_rn(T, __fmul_rn(a, feat[u].z), C2) : C2;
          T = acc ? test_T : T;
          last_contributor = acc ? base + (uint32_t)jj[u] : last_contributor;
        }
      }
      const bool finished = __all_sync(kFull, done);  // also orders this stage's reads before its refill
      if (finished) break;
      if (lane == 0 && b + kStages < nbatch) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(b + kStages);
      }
      if (b + kStages < nbatch) issued = b + kStages + 1;
    }
    // never leave with a bulk copy still in flight towards this CTA's shared memory
    if (lane == 0)
      for (int o = b + 1; o < issued; ++o) mbar_wait(&ring.full[o % kStages], (uint32_t)(o / kStages) & 1u);
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
  if (lane == 0 && wmax) atomicMax(&tile_last[tile], wmax);  // tile_last is zeroed by order_tiles
}

}  // namespace

cudaError_t launch_blend_fwd(int width, int height, const ImageState& img, const BinningState& b,
                             const float* background, float* out_color, float* out_depth, float* out_acc,
                             cudaStream_t s) {
  const int grid_x = (width + kTile - 1) / kTile, grid_y = (height + kTile - 1) / kTile;
  constexpr int kSmem = (int)sizeof(WarpRing) * kWarps;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(blend_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  blend_fwd_kernel<<<grid_x * grid_y, kWarps * 32, kSmem, s>>>(img.ranges, img.tile_order, b.inst, width, height,
                                                              grid_x, background, out_color, out_depth, out_acc,
                                                              img.final_T, img.n_contrib, img.tile_last);
  return cudaGetLastError();
}

}  // namespace gsr
