// Back-to-front backward of the tile blend: dL/dC, dL/dS -> per-Gaussian dL/d{mean2D, conic, opacity,
// colour}. sm_100a.
//
// Replaces BACKWARD::render / renderCUDA<3> (reference src/cuda_rasterizer/backward.cu:438-603,675-713).
// Like the reference, NO depth gradient is produced (backward.cu:567-571 hard-codes c_d = 1 for the
// silhouette channel and rasterizer.cu:77-80 drops grad_out_depth).
//
// Design ("blend_bwd"):
//  * same structure as blend_fwd, traversed in reverse: 8 fully independent warps per tile, each with a
//    private TMA-fed ring (records + Gaussian ids) and private mbarriers, sub-tile ballot culling with
//    survivor compaction; everything behind the warp's own last contributor (max n_contrib of its 32
//    pixels) is never fetched; survivors are processed two at a time so that the exp / alpha of the
//    second overlaps the recurrence and warp reduction of the first;
//  * the reference issues 9 float atomics per blended (pixel, Gaussian) pair, all lanes of a warp
//    hitting the same 9 addresses. Here the warp walks the same Gaussian in lock-step, the 9 partial
//    sums are reduced across the 32 lanes with a transposing butterfly (14 shuffles for 9 values
//    instead of 45) and leave through ONE predicated RED.ADD.F32 instruction in which 9 lanes each own
//    a distinct address: 32x fewer L2 atomics and no same-address serialisation.
//  Summation order differs from the reference (as it does between two runs of the reference itself),
//  so gradients are compared at 1e-4 relative, not bitwise.
#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kStages = 3;
constexpr int kWarps = 8;
constexpr int kWB = 64;  // instances per per-warp batch
constexpr int kIlp = 2;
constexpr uint32_t kFull = 0xffffffffu;
constexpr int kIdPad = 4;  // ids are bulk-copied from a 16-byte aligned-down start

// Sum 8 per-lane values across the warp; afterwards every lane holds the total of value (lane >> 2) & 7.
__device__ __forceinline__ float transpose_reduce8(const float v[8], int lane) {
  float w[4], u[2];
  bool up = lane & 16;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = up ? v[i] : v[i + 4];
    const float keep = up ? v[i + 4] : v[i];
    w[i] = keep + __shfl_xor_sync(kFull, send, 16);
  }
  up = lane & 8;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = up ? w[i] : w[i + 2];
    const float keep = up ? w[i + 2] : w[i];
    u[i] = keep + __shfl_xor_sync(kFull, send, 8);
  }
  up = lane & 4;
  float t = (up ? u[1] : u[0]) + __shfl_xor_sync(kFull, up ? u[0] : u[1], 4);
  t += __shfl_xor_sync(kFull, t, 2);
  t += __shfl_xor_sync(kFull, t, 1);
  return t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

struct __align__(128) WarpRing {
  float4 rec[kStages][kWB * 3];
  uint32_t id[kStages][kWB + kIdPad];
  uint64_t full[kStages];
  uint32_t list[kWB / 4];  // compacted survivor indices (descending), one byte each
};

__global__ void __launch_bounds__(kWarps * 32) blend_bwd_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order, const SplatRecord* __restrict__ inst,
    const uint32_t* __restrict__ point_list, int W, int H, int grid_x, const float* __restrict__ bg,
    const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
    const float* __restrict__ dL_dpix, const float* __restrict__ dL_dacc, float* __restrict__ dL_dmean2D,
    float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpRing& ring = reinterpret_cast<WarpRing*>(smem_raw)[warp];

  const int tile = (int)tile_order[blockIdx.x];
  const uint2 range = ranges[tile];
  const int total = (int)(range.y - range.x);

  const int tile_x = tile % grid_x, tile_y = tile / grid_x;
  const int wx = tile_x * kTile + (warp & 1) * 8, wy = tile_y * kTile + (warp >> 1) * 4;
  const int px = wx + (lane & 7), py = wy + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const float wx0 = (float)wx, wx1 = (float)(wx + 7), wy0 = (float)wy, wy1 = (float)(wy + 3);

  const size_t pix = (size_t)W * py + px;
  const size_t plane = (size_t)W * H;
  const uint32_t last_contributor = inside ? n_contrib[pix] : 0u;
  const uint32_t warp_last = __reduce_max_sync(kFull, last_contributor);
  if (warp_last == 0) return;  // nothing was blended into this warp's 32 pixels

  const float T_final = inside ? final_T[pix] : 0.f;
  float dpx0 = 0.f, dpx1 = 0.f, dpx2 = 0.f, dacc = 0.f;
  if (inside) {
    dpx0 = dL_dpix[pix], dpx1 = dL_dpix[plane + pix], dpx2 = dL_dpix[2 * plane + pix];
    dacc = dL_dacc[pix];
  }
  const float bg_dot = __ldg(bg) * dpx0 + __ldg(bg + 1) * dpx1 + __ldg(bg + 2) * dpx2;

  float T = T_final;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc_s = 0.f;  // accum_rec, accum_acc_rec
  float lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_s = 0.f, last_alpha = 0.f;
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

  // which of the 9 gradient words this lane commits after the warp reduction
  float* gptr = nullptr;
  int gstride = 0;
  {
    const int v = lane >> 2;
    if ((lane & 3) == 0) {
      if (v < 3) gptr = dL_dcolor + v, gstride = 3;
      else if (v < 5) gptr = dL_dmean2D + (v - 3), gstride = 3;
      else gptr = dL_dconic + (v == 7 ? 3 : v - 5), gstride = 4;
    } else if (lane == 1) {
      gptr = dL_dopacity, gstride = 1;
    }
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
    bulk_g2s(ring.rec[stage], inst + start, rec_bytes, &ring.full[stage]);
    bulk_g2s(ring.id[stage], point_list + a0, id_bytes, &ring.full[stage]);
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

    for (int g = 0; g < nhit; g += kIlp) {
      // kIlp survivors per step (back to front): G / alpha of all are independent of the recurrence
      const uint32_t packed = ring.list[g >> 2] >> (8 * (g & 3));
      int jj[kIlp];
      float G[kIlp], alpha[kIlp], dxs[kIlp], dys[kIlp];
#pragma unroll
      for (int u = 0; u < kIlp; ++u) {
        const bool valid = g + u < nhit;
        jj[u] = valid ? (int)((packed >> (8 * u)) & 0xffu) : (int)(packed & 0xffu);
        const float4 q0 = s[3 * jj[u]];
        const float4 q1 = s[3 * jj[u] + 1];
        const float dx = q0.x - pxf, dy = q0.y - pyf;
        const float power = __fmaf_rn(__fmaf_rn(dx, __fmul_rn(dx, q1.x), __fmul_rn(dy, __fmul_rn(dy, q1.z))), -0.5f,
                                      -__fmul_rn(dy, __fmul_rn(dx, q1.y)));
        const float gg = expf(power);
        const float a = fminf(0.99f, __fmul_rn(q1.w, gg));
        const bool on = valid && (batch_base + (uint32_t)jj[u] < last_contributor) && !(power > 0.0f) &&
                        !(a < 1.0f / 255.0f);
        G[u] = gg, alpha[u] = on ? a : -1.0f, dxs[u] = dx, dys[u] = dy;
      }
#pragma unroll
      for (int u = 0; u < kIlp; ++u) {
        const bool active = alpha[u] >= 0.0f;
        if (!__any_sync(kFull, active)) continue;
        float gr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float g_op = 0.f;
        if (active) {
          const float4 q1 = s[3 * jj[u] + 1];
          const float4 q2 = s[3 * jj[u] + 2];
          const float a = alpha[u], dx = dxs[u], dy = dys[u];
          const float one_m_alpha = 1.f - a;
          T = T / one_m_alpha;
          const float dchannel_dcolor = a * T;
          const float keep = 1.f - last_alpha;
          acc0 = last_alpha * lc0 + keep * acc0;
          acc1 = last_alpha * lc1 + keep * acc1;
          acc2 = last_alpha * lc2 + keep * acc2;
          acc_s = last_alpha * last_s + keep * acc_s;
          lc0 = q2.x, lc1 = q2.y, lc2 = q2.z, last_s = 1.0f;
          float dL_dalpha = (q2.x - acc0) * dpx0 + (q2.y - acc1) * dpx1 + (q2.z - acc2) * dpx2;
          dL_dalpha += (1.0f - acc_s) * dacc;
          dL_dalpha *= T;
          last_alpha = a;
          dL_dalpha += (-T_final / one_m_alpha) * bg_dot;
          const float dL_dG = q1.w * dL_dalpha;
          const float gdx = G[u] * dx, gdy = G[u] * dy;
          const float dG_ddelx = -gdx * q1.x - gdy * q1.y;
          const float dG_ddely = -gdy * q1.z - gdx * q1.y;
          gr[0] = dchannel_dcolor * dpx0;
          gr[1] = dchannel_dcolor * dpx1;
          gr[2] = dchannel_dcolor * dpx2;
          gr[3] = dL_dG * dG_ddelx * ddelx_dx;
          gr[4] = dL_dG * dG_ddely * ddely_dy;
          gr[5] = -0.5f * gdx * dx * dL_dG;
          gr[6] = -0.5f * gdx * dy * dL_dG;
          gr[7] = -0.5f * gdy * dy * dL_dG;
          g_op = G[u] * dL_dalpha;
        }
        const float t = transpose_reduce8(gr, lane);
        const float o = warp_sum(g_op);
        if (gptr) atomicAdd(gptr + (size_t)ids[jj[u]] * gstride, lane == 1 ? o : t);
      }
    }
    __syncwarp();  // all lanes are done with this stage before it is refilled
    if (lane == 0 && b - kStages >= 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(b - kStages, it + kStages);
    }
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
  blend_bwd_kernel<<<grid_x * grid_y, kWarps * 32, kSmem, s>>>(
      img.ranges, img.tile_order, b.inst, b.vals, width, height, grid_x, background, img.final_T, img.n_contrib,
      dL_dpix, dL_dacc, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
  return cudaGetLastError();
}

}  // namespace gsr
