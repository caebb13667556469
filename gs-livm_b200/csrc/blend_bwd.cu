// Back-to-front backward of the tile blend: dL/dC, dL/dS -> per-Gaussian dL/d{mean2D, conic, opacity,
// colour}. sm_100a.
//
// Replaces BACKWARD::render / renderCUDA<3> (reference src/cuda_rasterizer/backward.cu:438-603,675-713).
// Like the reference, NO depth gradient is produced (backward.cu:567-571 hard-codes c_d = 1 for the
// silhouette channel and rasterizer.cu:77-80 drops grad_out_depth).
//
// Design ("blend_bwd"):
//  * same structure as blend_fwd, traversed in reverse: every warp is independent, owns a small block of
//    pixels of one tile and a private TMA-fed ring (records + Gaussian ids) with private mbarriers;
//    sub-block ballot culling with survivor compaction; everything behind the warp's own last
//    contributor (max n_contrib of its pixels) is never fetched;
//  * the reference issues 9 float atomics per blended (pixel, Gaussian) pair, all lanes of a warp hitting
//    the same 9 addresses. Here the lanes walk the same Gaussian in lock-step, the 9 partial sums are
//    reduced across lanes with a transposing butterfly and leave through predicated RED.ADD.F32
//    instructions in which every participating lane owns a distinct address;
//  * two shapes of warp, as in blend_fwd. COARSE (short lists): 32 pixels (8x4) per warp, survivors two
//    at a time, one 14-shuffle butterfly per survivor. FINE (lists >= kHeavy): 8 pixels (4x2) per warp,
//    four lanes per pixel: the four lanes evaluate exp/alpha of four different survivors in parallel,
//    all replay the cheap ordered recurrence (T, accumulated colour) and snapshot the state in front of
//    their own survivor, then each 8-lane group computes the gradient terms of ONE of four consecutive
//    survivors, so the expensive part runs 4-wide and a single 10-shuffle butterfly (xor 4,2,1 stays
//    inside a group) reduces four survivors at once; one RED instruction commits 32 sums. This removes
//    the frame's longest serial chains (one warp walking ~2 400 survivors) and costs fewer
//    instructions per survivor than the coarse shape.
//  Summation order differs from the reference (as it does between two runs of the reference itself),
//  so gradients are compared at 1e-4 relative, not bitwise.
#include <cstdlib>

#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kStages = 3;
constexpr int kWarps = 8;
constexpr int kWB = 64;       // instances per per-warp batch
constexpr int kHeavy = 3072;   // tiles with at least this many instances use the FINE shape
constexpr int kFineCap = 384;  // FINE candidates: the first kFineCap tiles of tile_order (longest lists)
constexpr uint32_t kFull = 0xffffffffu;
constexpr int kIdPad = 4;     // ids are bulk-copied from a 16-byte aligned-down start

struct __align__(128) WarpRing {
  float4 rec[kStages][kWB * 3];
  uint32_t id[kStages][kWB + kIdPad];
  uint64_t full[kStages];
  uint32_t list[kWB / 4];  // compacted survivor indices (descending), one byte each
};

struct BwdArgs {
  const uint2* ranges;
  const uint32_t* tile_order;
  const SplatRecord* inst;
  const uint32_t* point_list;
  int W, H, grid_x, tiles, heavy, fine_tiles;
  const float* bg;
  const float* final_T;
  const uint32_t* n_contrib;
  const float *dL_dpix, *dL_dacc;
  float *dL_dmean2D, *dL_dconic, *dL_dopacity, *dL_dcolor;
};

// Sum 8 per-lane values across 32 lanes; afterwards every lane holds the total of value (lane >> 2) & 7.
__device__ __forceinline__ float transpose_reduce8_warp(const float v[8], int lane) {
  float w[4], u[2];
  bool up = lane & 16;
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = (up ? v[i + 4] : v[i]) + __shfl_xor_sync(kFull, up ? v[i] : v[i + 4], 16);
  up = lane & 8;
#pragma unroll
  for (int i = 0; i < 2; ++i) u[i] = (up ? w[i + 2] : w[i]) + __shfl_xor_sync(kFull, up ? w[i] : w[i + 2], 8);
  up = lane & 4;
  float t = (up ? u[1] : u[0]) + __shfl_xor_sync(kFull, up ? u[0] : u[1], 4);
  t += __shfl_xor_sync(kFull, t, 2);
  t += __shfl_xor_sync(kFull, t, 1);
  return t;
}

// Sum 8 per-lane values across each aligned group of 8 lanes; lane l ends with the total of value l & 7.
__device__ __forceinline__ float transpose_reduce8_oct(const float v[8], int lane) {
  float w[4], u[2];
  bool up = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = (up ? v[i + 4] : v[i]) + __shfl_xor_sync(kFull, up ? v[i] : v[i + 4], 4);
  up = lane & 2;
#pragma unroll
  for (int i = 0; i < 2; ++i) u[i] = (up ? w[i + 2] : w[i]) + __shfl_xor_sync(kFull, up ? w[i] : w[i + 2], 2);
  up = lane & 1;
  return (up ? u[1] : u[0]) + __shfl_xor_sync(kFull, up ? u[0] : u[1], 1);
}

// pointer / stride of gradient word v (0..2 colour, 3..4 mean2D.xy, 5..7 conic.x,.y,.w)
__device__ __forceinline__ void grad_slot(const BwdArgs& a, int v, float*& ptr, int& stride) {
  if (v < 3) ptr = a.dL_dcolor + v, stride = 3;
  else if (v < 5) ptr = a.dL_dmean2D + (v - 3), stride = 3;
  else ptr = a.dL_dconic + (v == 7 ? 3 : v - 5), stride = 4;
}

// Per-pixel state of the reverse traversal (reference backward.cu:480-501).
struct PixelState {
  float T, T_final, acc0, acc1, acc2, acc_s, lc0, lc1, lc2, last_s, last_alpha;
  float dpx0, dpx1, dpx2, dacc, bg_dot;
};

// One blended (pixel, Gaussian) pair: advance the recurrence and emit the 9 gradient terms
// (reference backward.cu:546-600). q1 = {conic, opacity}, q2 = {rgb, depth}.
__device__ __forceinline__ void pair_backward(PixelState& s, float alpha, float G, float dx, float dy,
                                              const float4& q1, const float4& q2, float ddelx_dx, float ddely_dy,
                                              float g[8], float& g_op) {
  // one correctly rounded reciprocal serves both divisions by (1 - alpha) of the reference (:546, :581)
  const float inv = __frcp_rn(1.f - alpha);
  s.T = s.T * inv;
  const float dchannel_dcolor = alpha * s.T;
  const float keep = 1.f - s.last_alpha;
  s.acc0 = s.last_alpha * s.lc0 + keep * s.acc0;
  s.acc1 = s.last_alpha * s.lc1 + keep * s.acc1;
  s.acc2 = s.last_alpha * s.lc2 + keep * s.acc2;
  s.acc_s = s.last_alpha * s.last_s + keep * s.acc_s;
  s.lc0 = q2.x, s.lc1 = q2.y, s.lc2 = q2.z, s.last_s = 1.0f;
  float dL_dalpha = (q2.x - s.acc0) * s.dpx0 + (q2.y - s.acc1) * s.dpx1 + (q2.z - s.acc2) * s.dpx2;
  dL_dalpha += (1.0f - s.acc_s) * s.dacc;
  dL_dalpha *= s.T;
  s.last_alpha = alpha;
  dL_dalpha += (-s.T_final * inv) * s.bg_dot;
  const float dL_dG = q1.w * dL_dalpha;
  const float gdx = G * dx, gdy = G * dy;
  const float dG_ddelx = -gdx * q1.x - gdy * q1.y;
  const float dG_ddely = -gdy * q1.z - gdx * q1.y;
  g[0] = dchannel_dcolor * s.dpx0;
  g[1] = dchannel_dcolor * s.dpx1;
  g[2] = dchannel_dcolor * s.dpx2;
  g[3] = dL_dG * dG_ddelx * ddelx_dx;
  g[4] = dL_dG * dG_ddely * ddely_dy;
  g[5] = -0.5f * gdx * dx * dL_dG;
  g[6] = -0.5f * gdx * dy * dL_dG;
  g[7] = -0.5f * gdy * dy * dL_dG;
  g_op = G * dL_dalpha;
}

template <int SLOTS>
__device__ __forceinline__ void blend_bwd_warp(const BwdArgs& a, WarpRing& ring, int tile, int block_in_tile,
                                               int lane) {
  constexpr int PIX = 32 / SLOTS;
  constexpr int BW = SLOTS == 1 ? 8 : 4, BH = PIX / BW;
  constexpr int ILP = 2;              // exp/alpha evaluations per lane per step
  constexpr int GROUP = ILP * SLOTS;  // survivors per step: 2 (coarse) or 8 (fine)

  const uint2 range = a.ranges[tile];
  const int total = (int)(range.y - range.x);
  const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
  const int wx = tile_x * kTile + (block_in_tile % (kTile / BW)) * BW;
  const int wy = tile_y * kTile + (block_in_tile / (kTile / BW)) * BH;
  const int pix = lane % PIX, slot = lane / PIX;
  const int px = wx + pix % BW, py = wy + pix / BW;
  const bool inside = px < a.W && py < a.H;
  const float pxf = (float)px, pyf = (float)py;
  const float wx0 = (float)wx, wx1 = (float)(wx + BW - 1), wy0 = (float)wy, wy1 = (float)(wy + BH - 1);

  const size_t p = (size_t)a.W * py + px;
  const size_t plane = (size_t)a.W * a.H;
  const uint32_t last_contributor = inside ? a.n_contrib[p] : 0u;
  const uint32_t warp_last = __reduce_max_sync(kFull, last_contributor);
  if (warp_last == 0) return;  // nothing was blended into this warp's pixels

  PixelState st;
  st.T_final = inside ? a.final_T[p] : 0.f;
  st.T = st.T_final;
  st.acc0 = st.acc1 = st.acc2 = st.acc_s = st.lc0 = st.lc1 = st.lc2 = st.last_s = st.last_alpha = 0.f;
  st.dpx0 = st.dpx1 = st.dpx2 = st.dacc = 0.f;
  if (inside) {
    st.dpx0 = a.dL_dpix[p], st.dpx1 = a.dL_dpix[plane + p], st.dpx2 = a.dL_dpix[2 * plane + p];
    st.dacc = a.dL_dacc[p];
  }
  st.bg_dot = __ldg(a.bg) * st.dpx0 + __ldg(a.bg + 1) * st.dpx1 + __ldg(a.bg + 2) * st.dpx2;
  const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;

  // which gradient word this lane commits after a reduction
  float* gptr = nullptr;
  int gstride = 0;
  if (SLOTS == 1) {
    if ((lane & 3) == 0) grad_slot(a, lane >> 2, gptr, gstride);
    else if (lane == 1) gptr = a.dL_dopacity, gstride = 1;
  } else {
    grad_slot(a, pix, gptr, gstride);  // lane (slot, pix) commits word `pix` of the survivor its group kept
  }

  // batches first_batch, first_batch-1, ..., 0; iteration `it` uses ring stage it % kStages
  const int first_batch = (int)((warp_last - 1) / kWB);
  auto issue = [&](int b, int it) {  // lane 0 only
    const int stage = it % kStages;
    const int cnt = min(kWB, total - b * kWB);
    const uint32_t rec_bytes = (uint32_t)cnt * (uint32_t)sizeof(SplatRecord);
    const size_t start = (size_t)range.x + (size_t)b * kWB;
    const size_t a0 = start & ~(size_t)3;  // bulk copies need 16-byte aligned global addresses
    const uint32_t id_bytes = (uint32_t)((((start - a0) + cnt) * 4 + 15) & ~(size_t)15);
    mbar_expect_tx(&ring.full[stage], rec_bytes + id_bytes);
    bulk_g2s(ring.rec[stage], a.inst + start, rec_bytes, &ring.full[stage]);
    bulk_g2s(ring.id[stage], a.point_list + a0, id_bytes, &ring.full[stage]);
  };
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&ring.full[s], 1);
    mbar_fence_init();
    for (int it = 0; it < kStages && first_batch - it >= 0; ++it) issue(first_batch - it, it);
  }
  __syncwarp();

  for (int b = first_batch, it = 0; b >= 0; --b, ++it) {
    const int stage = it % kStages;
    mbar_wait(&ring.full[stage], (uint32_t)(it / kStages) & 1u);
    const int cnt = min(kWB, total - b * kWB);
    const float4* s = ring.rec[stage];
    const uint32_t* ids = ring.id[stage] + (((size_t)range.x + (size_t)b * kWB) & 3);
    const uint32_t batch_base = (uint32_t)b * kWB;

    // ---- cull + compact, highest index first ----
    int nhit = 0;
    uint8_t* list8 = reinterpret_cast<uint8_t*>(ring.list);
#pragma unroll
    for (int k = kWB - 32; k >= 0; k -= 32) {
      const int j = k + lane;
      bool hit = false;
      if (j < cnt && batch_base + (uint32_t)j < warp_last) {
        const float4 c = s[3 * j];
        hit = (c.x + c.z >= wx0) && (c.x - c.z <= wx1) && (c.y + c.w >= wy0) && (c.y - c.w <= wy1);
      }
      const uint32_t m = __ballot_sync(kFull, hit);
      if (hit) list8[nhit + __popc(m & ~((2u << lane) - 1u))] = (uint8_t)j;  // rank among higher lanes
      nhit += __popc(m);
    }
    __syncwarp();

    for (int g = 0; g < nhit; g += GROUP) {
      // ---- exp / alpha of this lane's ILP survivors (independent of the recurrence) ----
      float G_mine[ILP], a_mine[ILP];
#pragma unroll
      for (int u = 0; u < ILP; ++u) {
        const int h = g + u * SLOTS + slot;
        const bool valid = h < nhit;
        const int jj = (int)list8[valid ? h : g];
        const float4 q0 = s[3 * jj];
        const float4 q1 = s[3 * jj + 1];
        const float dx = q0.x - pxf, dy = q0.y - pyf;
        const float power = __fmaf_rn(__fmaf_rn(dx, __fmul_rn(dx, q1.x), __fmul_rn(dy, __fmul_rn(dy, q1.z))), -0.5f,
                                      -__fmul_rn(dy, __fmul_rn(dx, q1.y)));
        const float gg = expf(power);
        const float al = fminf(0.99f, __fmul_rn(q1.w, gg));
        const bool on = valid && (batch_base + (uint32_t)jj < last_contributor) && !(power > 0.0f) &&
                        !(al < 1.0f / 255.0f);
        G_mine[u] = gg, a_mine[u] = on ? al : -1.0f;
      }

      if (SLOTS == 1) {
        // ---- COARSE: one 32-lane butterfly per survivor ----
#pragma unroll
        for (int t = 0; t < GROUP; ++t) {
          const float al = a_mine[t];
          const bool active = al >= 0.0f;
          if (!__any_sync(kFull, active)) continue;
          const int jj = (int)list8[min(g + t, nhit - 1)];
          float gr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          float g_op = 0.f;
          if (active) {
            const float4 q0 = s[3 * jj];
            pair_backward(st, al, G_mine[t], q0.x - pxf, q0.y - pyf, s[3 * jj + 1], s[3 * jj + 2], ddelx_dx, ddely_dy,
                          gr, g_op);
          }
          const float tsum = transpose_reduce8_warp(gr, lane);
          float o = g_op;
#pragma unroll
          for (int sh = 16; sh > 0; sh >>= 1) o += __shfl_xor_sync(kFull, o, sh);
          if (gptr) atomicAdd(gptr + (size_t)ids[jj] * gstride, lane == 1 ? o : tsum);
        }
      } else {
        // ---- FINE: every lane replays the cheap ordered recurrence of 4 consecutive survivors and snapshots the
        // state in front of ITS survivor (the one its 8-lane group evaluated above); the expensive gradient
        // terms are then computed once, for four survivors in parallel, and reduced by one butterfly ----
#pragma unroll
        for (int half = 0; half < GROUP / SLOTS; ++half) {
          const int h_mine = g + half * SLOTS + slot;
          const int jj_mine = (int)list8[min(h_mine, nhit - 1)];
          float snap_T = 0.f, snap_a0 = 0.f, snap_a1 = 0.f, snap_a2 = 0.f, snap_as = 0.f;
          uint32_t any_bits = 0;
#pragma unroll
          for (int q = 0; q < SLOTS; ++q) {
            const float al = __shfl_sync(kFull, a_mine[half], q * PIX + pix);
            const bool active = al >= 0.0f;
            if (!__any_sync(kFull, active)) continue;
            any_bits |= 1u << q;
            if (active) {
              const int jj = (int)list8[min(g + half * SLOTS + q, nhit - 1)];
              const float4 q2 = s[3 * jj + 2];
              st.T = st.T * __frcp_rn(1.f - al);
              const float keep = 1.f - st.last_alpha;
              st.acc0 = st.last_alpha * st.lc0 + keep * st.acc0;
              st.acc1 = st.last_alpha * st.lc1 + keep * st.acc1;
              st.acc2 = st.last_alpha * st.lc2 + keep * st.acc2;
              st.acc_s = st.last_alpha * st.last_s + keep * st.acc_s;
              st.lc0 = q2.x, st.lc1 = q2.y, st.lc2 = q2.z, st.last_s = 1.0f;
              st.last_alpha = al;
              if (slot == q) snap_T = st.T, snap_a0 = st.acc0, snap_a1 = st.acc1, snap_a2 = st.acc2, snap_as = st.acc_s;
            }
          }
          if (any_bits == 0) continue;
          float gr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          float g_op = 0.f;
          const float al = a_mine[half];
          if (al >= 0.0f) {  // gradient terms of this lane's own survivor (reference backward.cu:547-600)
            const float4 q0 = s[3 * jj_mine], q1 = s[3 * jj_mine + 1], q2 = s[3 * jj_mine + 2];
            const float dx = q0.x - pxf, dy = q0.y - pyf, G = G_mine[half];
            const float dchannel_dcolor = al * snap_T;
            float dL_dalpha = (q2.x - snap_a0) * st.dpx0 + (q2.y - snap_a1) * st.dpx1 + (q2.z - snap_a2) * st.dpx2;
            dL_dalpha += (1.0f - snap_as) * st.dacc;
            dL_dalpha *= snap_T;
            dL_dalpha += (-st.T_final * __frcp_rn(1.f - al)) * st.bg_dot;
            const float dL_dG = q1.w * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddelx = -gdx * q1.x - gdy * q1.y;
            const float dG_ddely = -gdy * q1.z - gdx * q1.y;
            gr[0] = dchannel_dcolor * st.dpx0;
            gr[1] = dchannel_dcolor * st.dpx1;
            gr[2] = dchannel_dcolor * st.dpx2;
            gr[3] = dL_dG * dG_ddelx * ddelx_dx;
            gr[4] = dL_dG * dG_ddely * ddely_dy;
            gr[5] = -0.5f * gdx * dx * dL_dG;
            gr[6] = -0.5f * gdx * dy * dL_dG;
            gr[7] = -0.5f * gdy * dy * dL_dG;
            g_op = G * dL_dalpha;
          }
          const float tsum = transpose_reduce8_oct(gr, lane);
          float o = g_op;
          o += __shfl_xor_sync(kFull, o, 4);
          o += __shfl_xor_sync(kFull, o, 2);
          o += __shfl_xor_sync(kFull, o, 1);
          if ((any_bits >> slot) & 1u) {
            const size_t id = ids[jj_mine];
            atomicAdd(gptr + id * gstride, tsum);
            if (pix == 0) atomicAdd(a.dL_dopacity + id, o);
          }
        }
      }
    }
    __syncwarp();  // all lanes are done with this stage before it is refilled
    if (lane == 0 && b - kStages >= 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(b - kStages, it + kStages);
    }
  }
}

// grid = 4*fine_tiles FINE candidates (tile_order[b/4], quarter b%4) followed by tiles COARSE candidates.
__global__ void __launch_bounds__(kWarps * 32) blend_bwd_kernel(const BwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpRing& ring = reinterpret_cast<WarpRing*>(smem_raw)[warp];
  const int b = (int)blockIdx.x;
  if (b < 4 * a.fine_tiles) {
    const int tile = (int)a.tile_order[b >> 2];
    const uint2 r = a.ranges[tile];
    if ((int)(r.y - r.x) < a.heavy) return;
    blend_bwd_warp<4>(a, ring, tile, (b & 3) * kWarps + warp, lane);
  } else {
    const int tile = (int)a.tile_order[b - 4 * a.fine_tiles];
    const uint2 r = a.ranges[tile];
    // FINE CTAs exist for the first fine_tiles entries of the order only: a heavy tile beyond them is done here
    if ((int)(r.y - r.x) >= a.heavy && b - 4 * a.fine_tiles < a.fine_tiles) return;
    blend_bwd_warp<1>(a, ring, tile, warp, lane);
  }
}

}  // namespace

cudaError_t launch_blend_bwd(int width, int height, const ImageState& img, const BinningState& b,
                             const float* background, const float* dL_dpix, const float* dL_dacc, float* dL_dmean2D,
                             float* dL_dconic, float* dL_dopacity, float* dL_dcolor, cudaStream_t s) {
  const int grid_x = (width + kTile - 1) / kTile, grid_y = (height + kTile - 1) / kTile;
  constexpr int kSmem = (int)sizeof(WarpRing) * kWarps;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(blend_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  BwdArgs a;
  a.ranges = img.ranges, a.tile_order = img.tile_order, a.inst = b.inst, a.point_list = b.vals;
  a.W = width, a.H = height, a.grid_x = grid_x, a.tiles = grid_x * grid_y;
  static const int heavy = [] {
    const char* e = getenv("GSR_BWD_HEAVY");  // tuning knob: list length from which a tile uses the FINE warp shape
    return e ? atoi(e) : kHeavy;
  }();
  a.heavy = heavy;
  a.fine_tiles = a.tiles < kFineCap ? a.tiles : kFineCap;  // only the longest kFineCap lists can be FINE
  a.bg = background, a.final_T = img.final_T, a.n_contrib = img.n_contrib;
  a.dL_dpix = dL_dpix, a.dL_dacc = dL_dacc;
  a.dL_dmean2D = dL_dmean2D, a.dL_dconic = dL_dconic, a.dL_dopacity = dL_dopacity, a.dL_dcolor = dL_dcolor;
  blend_bwd_kernel<<<4 * a.fine_tiles + a.tiles, kWarps * 32, kSmem, s>>>(a);
  return cudaGetLastError();
}

}  // namespace gsr
