// Front-to-back alpha blend of C (colour), D (depth) and S (silhouette) per 16x16 tile, sm_100a.
//
// Replaces FORWARD::render / renderCUDA<3> (reference src/cuda_rasterizer/forward.cu:291-441).
//
// Design ("blend_fwd"):
//  * the unit of work is one WARP that owns a small block of pixels of one tile and is FULLY independent
//    of every other warp: it stages its own batches of 64 sorted instances (48-byte records, contiguous
//    thanks to finalize_binning) into a private 3-deep shared-memory ring with TMA bulk copies
//    (cp.async.bulk -> SASS UBLKCP) that complete on private mbarriers. There is no CTA-wide barrier
//    and no shared ring: ncu on the first two versions of this kernel (profiles/) showed 60 % of all
//    stall samples on the per-batch __syncthreads, then on a shared ring's full/empty waits, because
//    the sub-blocks of a tile are busy at different depths of the list. The price is that a record is
//    fetched from L2 by every warp of the tile; the instance array is L2-resident (48 B x R);
//  * per batch the warp tests 32 records per instruction against its pixel block with the per-Gaussian
//    conservative extent (preprocess.cu), compacts the survivors' indices with __ballot_sync/__popc
//    into a small per-warp list, and only walks those. Skipped pairs are exactly pairs the reference
//    `continue`s on, so results (including n_contrib, which indexes the full list) are unchanged;
//  * survivors are processed several at a time: their alphas (conic form + expf) are independent and
//    overlap in the pipeline, only the cheap transmittance recurrence is applied in order, branch-free;
//  * two shapes of warp. COARSE (tiles with short lists): 32 pixels (8x4) per warp, 8 warps per tile.
//    FINE (tiles whose list is >= kHeavy instances, a few percent of the tiles but ~40 % of the work on
//    NTU-VIRAL-shaped scenes): 8 pixels (4x2) per warp and FOUR lanes per pixel; the four lanes evaluate
//    the alphas of four different survivors in parallel and exchange them with shuffles, then all apply
//    the same ordered recurrence. That cuts the longest serial chain of the frame (one warp walking
//    ~2 900 survivors) by ~3.5x at ~20 % more total instructions for those tiles;
//  * a warp whose pixels are all saturated stops fetching and exits;
//  * CTAs are launched heaviest-tile-first (tile_order, written next to finalize_binning).
//
// Numerics: the per-pair operation DAG (which product is fused, order of accumulation) mirrors the
// reference's sm_100a SASS through explicit _rn intrinsics; exp is the
// same precise expf. With identical inputs C, D, S, final_T and n_contrib come out bit-identical.
#include <cstdlib>

#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kStages = 3;
constexpr int kWarps = 8;
constexpr int kWB = 64;        // instances per per-warp batch
constexpr int kHeavy = 4096;   // tiles with at least this many instances use the FINE shape
constexpr int kFineCap = 384;  // FINE candidates: the first kFineCap tiles of tile_order (longest lists)
constexpr uint32_t kFull = 0xffffffffu;

struct __align__(128) WarpRing {
  float4 rec[kStages][kWB * 3];
  uint64_t full[kStages];
  uint32_t list[kWB / 4];  // compacted survivor indices, one byte each
};

struct FwdArgs {
  const uint2* ranges;
  const uint32_t* tile_order;
  const SplatRecord* inst;
  int W, H, grid_x, tiles, heavy, fine_tiles;
  const float* bg;
  float *out_color, *out_depth, *out_acc, *final_T;
  uint32_t *n_contrib, *tile_last;
};

// SLOTS lanes cooperate on one pixel; a warp covers 32/SLOTS pixels: 8x4 (SLOTS = 1) or 4x2 (SLOTS = 4).
template <int SLOTS>
__device__ __forceinline__ void blend_warp(const FwdArgs& a, WarpRing& ring, int tile, int block_in_tile, int lane) {
  constexpr int PIX = 32 / SLOTS;
  constexpr int BW = SLOTS == 1 ? 8 : 4, BH = PIX / BW;
  constexpr int ILP = SLOTS == 1 ? 4 : 2;  // alphas per lane per step
  constexpr int GROUP = ILP * SLOTS;       // survivors per step
  static_assert(GROUP == 4 || GROUP == 8, "list words are consumed whole");

  const uint2 range = a.ranges[tile];
  const int total = (int)(range.y - range.x);
  const int nbatch = (total + kWB - 1) / kWB;
  const SplatRecord* src = a.inst + range.x;

  const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
  const int wx = tile_x * kTile + (block_in_tile % (kTile / BW)) * BW;
  const int wy = tile_y * kTile + (block_in_tile / (kTile / BW)) * BH;
  const int pix = lane % PIX, slot = lane / PIX;
  const int px = wx + pix % BW, py = wy + pix / BW;
  const bool inside = px < a.W && py < a.H;
  const float pxf = (float)px, pyf = (float)py;
  const float wx0 = (float)wx, wx1 = (float)(wx + BW - 1), wy0 = (float)wy, wy1 = (float)(wy + BH - 1);

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

      // ---- cull 32 records per instruction against this warp's pixel block, compact survivors ----
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

      // ---- walk the survivors GROUP at a time: alphas first (independent), then the ordered recurrence ----
      for (int g = 0; g < nhit; g += GROUP) {
        float alpha_mine[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
          const int h = g + u * SLOTS + slot;  // survivor this lane evaluates
          const bool valid = h < nhit;
          const int jj = (int)list8[valid ? h : g];
          const float4 q0 = s[3 * jj];
          const float4 q1 = s[3 * jj + 1];  // {conic.x, conic.y, conic.z, opacity}
          const float dx = __fadd_rn(q0.x, -pxf);
          const float dy = __fadd_rn(q0.y, -pyf);
          const float quad = __fmaf_rn(dx, __fmul_rn(dx, q1.x), __fmul_rn(dy, __fmul_rn(dy, q1.z)));
          const float power = __fmaf_rn(quad, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, q1.y)));
          const float al = fminf(0.99f, __fmul_rn(q1.w, expf(power)));
          // reference: `if (power > 0) continue; ... if (alpha < 1/255) continue;`
          alpha_mine[u] = (valid && !done && !(power > 0.0f) && !(al < 1.0f / 255.0f)) ? al : -1.0f;
        }
        // ordered transmittance recurrence, branch-free per lane; a survivor no lane needs is skipped warp-wide
#pragma unroll
        for (int t = 0; t < GROUP; ++t) {
          float al = alpha_mine[t / SLOTS];
          if (SLOTS > 1) al = __shfl_sync(kFull, al, (t % SLOTS) * PIX + pix);
          if (__ballot_sync(kFull, al >= 0.0f) == 0) continue;
          const int jj = (int)list8[min(g + t, nhit - 1)];
          const float4 q2 = s[3 * jj + 2];  // {r, g, b, depth}
          const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -al));
          const bool live = (al >= 0.0f) && !done;
          const bool kill = live && (test_T < 0.0001f);  // reference: done = true; continue
          const bool acc = live && !kill;
          done = done || kill;
          A = acc ? __fmaf_rn(T, al, A) : A;
          D = acc ? __fmaf_rn(T, __fmul_rn(al, q2.w), D) : D;
          C0 = acc ? __fmaf_rn(T, __fmul_rn(al, q2.x), C0) : C0;
          C1 = acc ? __fmaf_rn(T, __fmul_rn(al, q2.y), C1) : C1;
          C2 = acc ? __fmaf_rn(T, __fmul_rn(al, q2.z), C2) : C2;
          T = acc ? test_T : T;
          last_contributor = acc ? base + (uint32_t)jj : last_contributor;
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

  if (inside && slot == 0) {
    const size_t p = (size_t)a.W * py + px;
    const size_t plane = (size_t)a.W * a.H;
    a.final_T[p] = T;
    a.n_contrib[p] = last_contributor;
    a.out_color[p] = __fmaf_rn(__ldg(a.bg + 0), T, C0);
    a.out_color[plane + p] = __fmaf_rn(__ldg(a.bg + 1), T, C1);
    a.out_color[2 * plane + p] = __fmaf_rn(__ldg(a.bg + 2), T, C2);
    a.out_depth[p] = D;
    a.out_acc[p] = A;
  }
  const uint32_t wmax = __reduce_max_sync(kFull, last_contributor);
  if (lane == 0 && wmax) atomicMax(&a.tile_last[tile], wmax);  // tile_last is zeroed by order_tiles
}

// grid = 4*fine_tiles FINE candidates (tile_order[b/4], quarter b%4) followed by tiles COARSE candidates; a CTA
// whose tile belongs to the other shape exits at once. Heavy tiles therefore start first.
__global__ void __launch_bounds__(kWarps * 32) blend_fwd_kernel(const FwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpRing& ring = reinterpret_cast<WarpRing*>(smem_raw)[warp];
  const int b = (int)blockIdx.x;
  if (b < 4 * a.fine_tiles) {
    const int tile = (int)a.tile_order[b >> 2];
    const uint2 r = a.ranges[tile];
    if ((int)(r.y - r.x) < a.heavy) return;
    blend_warp<4>(a, ring, tile, (b & 3) * kWarps + warp, lane);
  } else {
    const int tile = (int)a.tile_order[b - 4 * a.fine_tiles];
    const uint2 r = a.ranges[tile];
    // FINE CTAs exist for the first fine_tiles entries of the order only: a heavy tile beyond them is done here
    if ((int)(r.y - r.x) >= a.heavy && b - 4 * a.fine_tiles < a.fine_tiles) return;
    blend_warp<1>(a, ring, tile, warp, lane);
  }
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
  FwdArgs a;
  a.ranges = img.ranges, a.tile_order = img.tile_order, a.inst = b.inst;
  a.W = width, a.H = height, a.grid_x = grid_x, a.tiles = grid_x * grid_y;
  static const int heavy = [] {
    const char* e = getenv("GSR_FWD_HEAVY");  // tuning knob: list length from which a tile uses the FINE warp shape
    return e ? atoi(e) : kHeavy;
  }();
  a.heavy = heavy;
  a.fine_tiles = a.tiles < kFineCap ? a.tiles : kFineCap;  // only the longest kFineCap lists can be FINE
  a.bg = background, a.out_color = out_color, a.out_depth = out_depth, a.out_acc = out_acc;
  a.final_T = img.final_T, a.n_contrib = img.n_contrib, a.tile_last = img.tile_last;
  blend_fwd_kernel<<<4 * a.fine_tiles + a.tiles, kWarps * 32, kSmem, s>>>(a);
  return cudaGetLastError();
}

}  // namespace gsr
