// Back-to-front backward of the tile blend: dL/dC, dL/dS -> per-Gaussian dL/d{mean2D, conic, opacity,
// colour}. sm_100a.
//
// Replaces BACKWARD::render / renderCUDA<3> (reference src/cuda_rasterizer/backward.cu:438-603,675-713).
// Like the reference, NO depth gradient is produced (backward.cu:567-571 hard-codes c_d = 1 for the
// silhouette channel and rasterizer.cu:77-80 drops grad_out_depth).
//
// Design ("blend_bwd"):
//  * same tiling / TMA bulk staging / sub-tile ballot culling as blend_fwd, traversed in reverse;
//    whole batches behind the tile's last contributor (tile_last, written by the forward) are never
//    loaded;
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

constexpr int kStages = 2;
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

__global__ void __launch_bounds__(kTilePixels) blend_bwd_kernel(
    const uint2* __restrict__ ranges, const SplatRecord* __restrict__ inst, const uint32_t* __restrict__ point_list,
    int W, int H, int grid_x, const float* __restrict__ bg, const float* __restrict__ final_T,
    const uint32_t* __restrict__ n_contrib, const uint32_t* __restrict__ tile_last,
    const float* __restrict__ dL_dpix, const float* __restrict__ dL_dacc, float* __restrict__ dL_dmean2D,
    float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor) {
  __shared__ __align__(128) float4 s_rec[kStages][kBatch * 3];
  __shared__ __align__(16) uint32_t s_id[kStages][kBatch + kIdPad];
  __shared__ __align__(8) uint64_t s_full[kStages];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const uint32_t t_last = tile_last[tile];
  if (t_last == 0) return;  // nothing was blended in this tile (uniform for the CTA)

  const int tile_x = tile % grid_x, tile_y = tile / grid_x;
  const int wx = tile_x * kTile + (warp & 1) * 8, wy = tile_y * kTile + (warp >> 1) * 4;
  const int px = wx + (lane & 7), py = wy + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const float wx0 = (float)wx, wx1 = (float)(wx + 7), wy0 = (float)wy, wy1 = (float)(wy + 3);

  const uint2 range = ranges[tile];
  const int total = (int)(range.y - range.x);
  const int first_batch = (int)((t_last - 1) / kBatch);  // batch holding the tile's last contributor

  auto issue = [&](int b) {  // elected thread: stage batch b (records + Gaussian ids)
    const int stage = (first_batch - b) & 1;
    const int cnt = min(kBatch, total - b * kBatch);
    const uint32_t rec_bytes = (uint32_t)cnt * (uint32_t)sizeof(SplatRecord);
    const size_t start = (size_t)range.x + (size_t)b * kBatch;
    const size_t a0 = start & ~(size_t)3;
    const uint32_t id_bytes = (uint32_t)((((start - a0) + cnt) * 4 + 15) & ~(size_t)15);
    mbar_expect_tx(&s_full[stage], rec_bytes + id_bytes);
    bulk_g2s(s_rec[stage], inst + start, rec_bytes, &s_full[stage]);
    bulk_g2s(s_id[stage], point_list + a0, id_bytes, &s_full[stage]);
  };

  if (tid == 0) {
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    issue(first_batch);
    if (first_batch >= 1) issue(first_batch - 1);
  }

  const size_t pix = (size_t)W * py + px;
  const size_t plane = (size_t)W * H;
  const float T_final = inside ? final_T[pix] : 0.f;
  const uint32_t last_contributor = inside ? n_contrib[pix] : 0u;
  float dpx0 = 0.f, dpx1 = 0.f, dpx2 = 0.f, dacc = 0.f;
  if (inside) {
    dpx0 = dL_dpix[pix], dpx1 = dL_dpix[plane + pix], dpx2 = dL_dpix[2 * plane + pix];
    dacc = dL_dacc[pix];
  }
  const float bg_dot = __ldg(bg) * dpx0 + __ldg(bg + 1) * dpx1 + __ldg(bg + 2) * dpx2;
  const uint32_t warp_last = __reduce_max_sync(kFull, last_contributor);

  float T = T_final;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc_s = 0.f;      // accum_rec, accum_acc_rec
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

  for (int b = first_batch, it = 0; b >= 0; --b, ++it) {
    const int stage = it & 1;
    mbar_wait(&s_full[stage], (uint32_t)(it >> 1) & 1u);
    const int cnt = min(kBatch, total - b * kBatch);
    const float4* s = s_rec[stage];
    const uint32_t* ids = s_id[stage] + (((size_t)range.x + (size_t)b * kBatch) & 3);
    const uint32_t batch_base = (uint32_t)b * kBatch;

    if (batch_base < warp_last) {
      for (int k = ((cnt - 1) >> 5) << 5; k >= 0; k -= 32) {
        const int j = k + lane;
        bool hit = false;
        if (j < cnt && batch_base + (uint32_t)j < warp_last) {
          const float4 c = s[3 * j];
          hit = (c.x + c.z >= wx0) && (c.x - c.z <= wx1) && (c.y + c.w >= wy0) && (c.y - c.w <= wy1);
        }
        uint32_t m = __ballot_sync(kFull, hit);
        while (m) {
          const int bit = 31 - __clz(m);
          m &= ~(1u << bit);
          const int jj = k + bit;
          const float4 q0 = s[3 * jj];
          const float4 q1 = s[3 * jj + 1];
          float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          float g_op = 0.f;
          bool active = false;
          if (batch_base + (uint32_t)jj < last_contributor) {
            const float dx = q0.x - pxf, dy = q0.y - pyf;
            const float power = __fmaf_rn(__fmaf_rn(dx, __fmul_rn(dx, q1.x), __fmul_rn(dy, __fmul_rn(dy, q1.z))),
                                          -0.5f, -__fmul_rn(dy, __fmul_rn(dx, q1.y)));
            if (power <= 0.0f) {
              const float G = expf(power);
              const float alpha = fminf(0.99f, __fmul_rn(q1.w, G));
              if (alpha >= 1.0f / 255.0f) {
                active = true;
                const float4 q2 = s[3 * jj + 2];
                const float one_m_alpha = 1.f - alpha;
                T = T / one_m_alpha;
                const float dchannel_dcolor = alpha * T;
                const float keep = 1.f - last_alpha;
                acc0 = last_alpha * lc0 + keep * acc0;
                acc1 = last_alpha * lc1 + keep * acc1;
                acc2 = last_alpha * lc2 + keep * acc2;
                acc_s = last_alpha * last_s + keep * acc_s;
                lc0 = q2.x, lc1 = q2.y, lc2 = q2.z, last_s = 1.0f;
                float dL_dalpha = (q2.x - acc0) * dpx0 + (q2.y - acc1) * dpx1 + (q2.z - acc2) * dpx2;
                dL_dalpha += (1.0f - acc_s) * dacc;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / one_m_alpha) * bg_dot;
                const float dL_dG = q1.w * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * q1.x - gdy * q1.y;
                const float dG_ddely = -gdy * q1.z - gdx * q1.y;
                g[0] = dchannel_dcolor * dpx0;
                g[1] = dchannel_dcolor * dpx1;
                g[2] = dchannel_dcolor * dpx2;
                g[3] = dL_dG * dG_ddelx * ddelx_dx;
                g[4] = dL_dG * dG_ddely * ddely_dy;
                g[5] = -0.5f * gdx * dx * dL_dG;
                g[6] = -0.5f * gdx * dy * dL_dG;
                g[7] = -0.5f * gdy * dy * dL_dG;
                g_op = G * dL_dalpha;
              }
            }
          }
          if (__any_sync(kFull, active)) {
            const float t = transpose_reduce8(g, lane);
            const float o = warp_sum(g_op);
            if (gptr) atomicAdd(gptr + (size_t)ids[jj] * gstride, lane == 1 ? o : t);
          }
        }
      }
    }

    __syncthreads();  // stage free again
    if (tid == 0 && b - kStages >= 0) issue(b - kStages);
  }
}

}  // namespace

cudaError_t launch_blend_bwd(int width, int height, const ImageState& img, const BinningState& b,
                             const float* background, const float* dL_dpix, const float* dL_dacc, float* dL_dmean2D,
                             float* dL_dconic, float* dL_dopacity, float* dL_dcolor, cudaStream_t s) {
  const int grid_x = (width + kTile - 1) / kTile, grid_y = (height + kTile - 1) / kTile;
  blend_bwd_kernel<<<grid_x * grid_y, kTilePixels, 0, s>>>(img.ranges, b.inst, b.vals, width, height, grid_x,
                                                          background, img.final_T, img.n_contrib, img.tile_last,
                                                          dL_dpix, dL_dacc, dL_dmean2D, dL_dconic, dL_dopacity,
                                                          dL_dcolor);
  return cudaGetLastError();
}

}  // namespace gsr
