// Back-to-front backward of the tile blend, coverage-mask formulation: dL/dC, dL/dS -> per-Gaussian
// dL/d{mean2D, conic, opacity, colour}. sm_100a.
//
// Replaces BACKWARD::render / renderCUDA<3> (reference src/cuda_rasterizer/backward.cu:438-603,675-713).
// Like the reference, NO depth gradient is produced (backward.cu:567-571, rasterizer.cu:77-80).
//
// Why. On voxelised maps an instance is blended into ~7 of its tile's 256 pixels. The reference (and any
// pixel-parallel kernel that walks instances in lock-step) pays for every (pixel, instance) pair of a block;
// ncu on the lock-step kernel (blend_bwd.cu) shows ~1000 issued thread-instructions per blended pair. Here the
// work is split so that each part runs over what it actually needs:
//
// One CTA (256 threads) per tile, batches of 128 sorted instances, last batch first (starting at the tile's
// last contributor), staged by TMA bulk copies (cp.async.bulk -> UBLKCP) into a 3-deep ring. Per batch:
//   A  two threads per instance: walk the pixels of the instance's box of {alpha >= 1/255} clipped to the tile
//      (16-byte aux record from the binning stage), zero the pair's hand-off slot, evaluate `power` with the
//      reference's operation DAG and set the instance's bit in the pixel's coverage mask where the pair can
//      pass the reference's tests (shared-memory atomicOr);
//   B  thread = pixel: walk the set bits back to front, redo the reference's exact tests, run the reference's
//      recurrence (T, accumulated colour) and leave two numbers per blended pair in the pair's slot:
//      s = G * dL/dalpha and w = alpha * T;
//   C  two threads per instance: sum the instance's slots into the 9 gradient words
//      (all linear in s and w: dL/dopacity = sum s, dL/dcolour = sum w dL/dC, the mean2D / conic terms are
//      o * s times polynomials of the pixel offset) and commit them with one RED.ADD.F32 per word.
// Instances whose clipped box is larger than kSparseMax pixels ("dense": close-up splats) keep the lock-step
// treatment inside phase B — all lanes evaluate them together, the 9 sums are reduced across the warp with a
// transposing butterfly — because for them most lanes of a pixel block do have work.
//
// Summation order differs from the reference (as it does between two runs of the reference itself), so
// gradients are compared at 1e-4 relative, not bitwise.
#include <cstdlib>

#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kThreads = 256;
constexpr int kNB = 128;           // instances per batch
constexpr int kWords = kNB / 32;
constexpr int kRing = 3;
constexpr int kSparseMax = 32;     // boxes up to this many pixels hand their pairs over through slots
constexpr int kWarpSlots = 16 * kSparseMax;  // 16 instances per warp in phases A and C
constexpr uint32_t kFull = 0xffffffffu;

struct __align__(128) BwdSmem {
  SplatRecord rec[kRing][kNB];      // 18 KB
  InstAux aux[kRing][kNB];          //  6 KB
  float2 slot[8 * kWarpSlots];      // 32 KB  {s, w} per (sparse instance, box pixel)
  uint32_t mask[kWords][kThreads];  //  4 KB  [word][y * 16 + x]
  float dpx[3][kThreads];           //  3 KB  dL/dC per pixel, [channel][y * 16 + x]
  uint32_t meta[kNB];               // x0 | y0 << 4 | box width << 8 | slot base << 13 | sparse << 30 | dense << 31
  uint32_t dense[kWords];           // bit = dense instance
  uint64_t full[kRing];
};

struct BwdArgs {
  const uint2* ranges;
  const uint32_t* tile_order;
  const uint32_t* tile_last;
  const SplatRecord* inst;
  const InstAux* aux;
  int W, H, grid_x, tiles;
  const float* bg;
  const float* final_T;
  const uint32_t* n_contrib;
  const float *dL_dpix, *dL_dacc;
  float *dL_dmean2D, *dL_dconic, *dL_dopacity, *dL_dcolor;
};

__device__ __forceinline__ float power_of(float dx, float dy, float ca, float cb, float cc) {
  const float quad = __fmaf_rn(dx, __fmul_rn(dx, ca), __fmul_rn(dy, __fmul_rn(dy, cc)));
  return __fmaf_rn(quad, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, cb)));
}

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

// One blended (pixel, Gaussian) pair: advance the recurrence (reference backward.cu:546-581) and return
// dL/dalpha; `weight` = alpha * T (dchannel_dcolor).
__device__ __forceinline__ float pair_step(PixelState& s, float alpha, const float4& q2, float& weight) {
  // one correctly rounded reciprocal serves both divisions by (1 - alpha) of the reference (:546, :581)
  const float inv = __frcp_rn(1.f - alpha);
  s.T = s.T * inv;
  weight = alpha * s.T;
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
  return dL_dalpha;
}

__global__ void __launch_bounds__(kThreads) blend_bwd_mask_kernel(const BwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = (int)a.tile_order[blockIdx.x];
  const uint32_t tile_last = a.tile_last[tile];
  if (tile_last == 0) return;  // nothing was blended in this tile
  const uint2 range = a.ranges[tile];
  const int total = (int)(range.y - range.x);
  const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
  const int tpx0 = tile_x * kTile, tpy0 = tile_y * kTile;
  const float tx0 = (float)tpx0, ty0 = (float)tpy0;

  // this thread's pixel (phase B): a warp owns an 8x4 block
  const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
  const int mp = ly * 16 + lx;
  const int pxi = tpx0 + lx, pyi = tpy0 + ly;
  const bool inside = pxi < a.W && pyi < a.H;
  const float pxf = (float)pxi, pyf = (float)pyi;
  const size_t gp = (size_t)a.W * pyi + pxi;
  const size_t plane = (size_t)a.W * a.H;
  const uint32_t last_contributor = inside ? a.n_contrib[gp] : 0u;

  PixelState st;
  st.T_final = inside ? a.final_T[gp] : 0.f;
  st.T = st.T_final;
  st.acc0 = st.acc1 = st.acc2 = st.acc_s = st.lc0 = st.lc1 = st.lc2 = st.last_s = st.last_alpha = 0.f;
  st.dpx0 = st.dpx1 = st.dpx2 = st.dacc = 0.f;
  if (inside) {
    st.dpx0 = a.dL_dpix[gp], st.dpx1 = a.dL_dpix[plane + gp], st.dpx2 = a.dL_dpix[2 * plane + gp];
    st.dacc = a.dL_dacc[gp];
  }
  st.bg_dot = __ldg(a.bg) * st.dpx0 + __ldg(a.bg + 1) * st.dpx1 + __ldg(a.bg + 2) * st.dpx2;
  sm.dpx[0][mp] = st.dpx0, sm.dpx[1][mp] = st.dpx1, sm.dpx[2][mp] = st.dpx2;
  const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;
  for (int w = 0; w < kWords; ++w) sm.mask[w][tid] = 0;
  if (tid < kWords) sm.dense[tid] = 0;

  // which gradient word this lane commits after a lock-step (dense) reduction
  float* gptr = nullptr;
  int gstride = 0;
  if ((lane & 3) == 0) grad_slot(a, lane >> 2, gptr, gstride);
  else if (lane == 1) gptr = a.dL_dopacity, gstride = 1;

  // batches first_batch, first_batch-1, ..., 0; iteration `it` uses ring stage it % kRing
  const int first_batch = (int)((tile_last - 1) / kNB);
  auto issue = [&](int b, int it) {  // thread 0 only
    const int stage = it % kRing;
    const uint32_t cnt = (uint32_t)min(kNB, total - b * kNB);
    const size_t first = (size_t)range.x + (size_t)b * kNB;
    mbar_expect_tx(&sm.full[stage], cnt * (uint32_t)(sizeof(SplatRecord) + sizeof(InstAux)));
    bulk_g2s(sm.rec[stage], a.inst + first, cnt * (uint32_t)sizeof(SplatRecord), &sm.full[stage]);
    bulk_g2s(sm.aux[stage], a.aux + first, cnt * (uint32_t)sizeof(InstAux), &sm.full[stage]);
  };
  if (tid == 0) {
    for (int s = 0; s < kRing; ++s) mbar_init(&sm.full[s], 1);
    mbar_fence_init();
    for (int it = 0; it < 2 && first_batch - it >= 0; ++it) issue(first_batch - it, it);
  }
  __syncthreads();

  // phases A and C: two threads per instance, interleaved rows
  const int ja = tid >> 1, half = tid & 1;
  float2* wslot = sm.slot + warp * kWarpSlots;

  for (int b = first_batch, it = 0; b >= 0; --b, ++it) {
    const int stage = it % kRing;
    const int cnt = min(kNB, total - b * kNB);
    const int batch_base = b * kNB;
    const SplatRecord* rec = sm.rec[stage];
    const InstAux* aux = sm.aux[stage];
    mbar_wait(&sm.full[stage], (uint32_t)(it / kRing) & 1u);

    // ---------------- phase A: instance -> coverage bits, zeroed slots ----------------
    uint32_t box = 0;
    float pmin = 0.f;
    if (ja < cnt && (uint32_t)(batch_base + ja) < tile_last) {
      const uint2 bx = *reinterpret_cast<const uint2*>(&aux[ja]);
      box = bx.x, pmin = __uint_as_float(bx.y);
    }
    const int x0 = box & 15, x1 = (box >> 4) & 15, y0 = (box >> 8) & 15, y1 = (box >> 12) & 15;
    const int bw = x1 - x0 + 1;
    const int area = (box & kBoxValid) ? bw * (y1 - y0 + 1) : 0;
    const bool sparse = area > 0 && area <= kSparseMax;
    const bool dense = area > kSparseMax;
    // slot base: exclusive prefix of the sparse areas over the warp's 16 instances (even lanes carry them)
    int incl = (sparse && half == 0) ? area : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += v;
    }
    int sbase = incl - ((sparse && half == 0) ? area : 0);
    sbase = __shfl_sync(kFull, sbase, lane & ~1);
    uint32_t* mword = sm.mask[ja >> 5];
    const uint32_t bit = 1u << (ja & 31);
    float2 xy = make_float2(0.f, 0.f);
    float4 q1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sparse) {
      xy = *reinterpret_cast<const float2*>(&rec[ja].q0);
      q1 = rec[ja].q1;
      const float fx0 = tx0 + (float)x0;
      for (int y = y0 + half; y <= y1; y += 2) {
        const float dy = __fadd_rn(xy.y, -(ty0 + (float)y));
        float fx = fx0;
        uint32_t* mrow = mword + y * 16;
        float2* srow = wslot + sbase + (y - y0) * bw - x0;
        for (int x = x0; x <= x1; ++x, fx += 1.0f) {
          srow[x] = make_float2(0.f, 0.f);
          const float power = power_of(__fadd_rn(xy.x, -fx), dy, q1.x, q1.y, q1.z);
          if (!(power > 0.0f) && !(power < pmin)) atomicOr(&mrow[x], bit);
        }
      }
    }
    if (half == 0 && ja < kNB)
      sm.meta[ja] = (uint32_t)x0 | ((uint32_t)y0 << 4) | ((uint32_t)bw << 8) |
                    ((uint32_t)(warp * kWarpSlots + sbase) << 13) | (sparse ? 1u << 30 : 0u) | (dense ? 1u << 31 : 0u);
    // dense boxes: the whole warp walks them, 32 pixels per step
    uint32_t big = __ballot_sync(kFull, dense && half == 0);
    if (lane == 0) {  // this warp's 16 instances are bits 16 * (warp & 1) .. +15 of word warp >> 1
      uint32_t dbits = 0;
      for (uint32_t t = big; t; t &= t - 1) dbits |= 1u << ((__ffs(t) - 1) >> 1);
      if (dbits) atomicOr(&sm.dense[warp >> 1], dbits << (16 * (warp & 1)));
    }
    while (big) {
      const int src = __ffs(big) - 1;
      big &= big - 1;
      const int jj = (tid & ~31) / 2 + (src >> 1);
      const uint2 bx = *reinterpret_cast<const uint2*>(&aux[jj]);
      const float smin = __uint_as_float(bx.y);
      const int sx0 = bx.x & 15, sy0 = (bx.x >> 8) & 15;
      const int sw = (int)((bx.x >> 4) & 15) - sx0 + 1, sarea = sw * ((int)((bx.x >> 12) & 15) - sy0 + 1);
      const float2 sxy = *reinterpret_cast<const float2*>(&rec[jj].q0);
      const float4 s1 = rec[jj].q1;
      const uint32_t sbit = 1u << (jj & 31);
      for (int i = lane; i < sarea; i += 32) {
        const int yy = i / sw, y = sy0 + yy, x = sx0 + (i - yy * sw);
        const float power = power_of(__fadd_rn(sxy.x, -(tx0 + (float)x)), __fadd_rn(sxy.y, -(ty0 + (float)y)), s1.x,
                                     s1.y, s1.z);
        if (!(power > 0.0f) && !(power < smin)) atomicOr(&mword[y * 16 + x], sbit);
      }
    }

    __syncthreads();  // (1) masks, slots, meta of this batch are complete; phase C of the previous batch is over
    if (tid == 0 && b - 2 >= 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(b - 2, it + 2);
    }

    // ---------------- phase B: pixel walks its coverage bits back to front ----------------
    {
      const int nw = (cnt + 31) >> 5;
      for (int w = nw - 1; w >= 0; --w) {
        uint32_t m = sm.mask[w][mp];
        if (m) sm.mask[w][mp] = 0;
        const int lim = (int)last_contributor - (batch_base + w * 32);  // bits below `lim` are contributors
        m = lim <= 0 ? 0u : (lim < 32 ? (m & ((1u << lim) - 1u)) : m);
        const uint32_t dall = sm.dense[w];
        uint32_t dm = __reduce_or_sync(kFull, m & dall);  // dense instances that reach this warp's pixels
        while (true) {
          const int d = dm ? 31 - __clz(dm) : -1;  // uniform
          const uint32_t above = d < 0 ? kFull : (d == 31 ? 0u : ~((2u << d) - 1u));
          uint32_t ms = m & ~dall & above;
          m &= ~above | dall;  // those sparse bits are consumed now
          while (ms) {  // sparse pairs: every lane on its own instance
            const int bpos = 31 - __clz(ms);
            ms ^= 1u << bpos;
            const int j = w * 32 + bpos;
            const float2 jxy = *reinterpret_cast<const float2*>(&rec[j].q0);
            const float4 j1 = rec[j].q1;
            const float dx = jxy.x - pxf, dy = jxy.y - pyf;
            const float power = power_of(dx, dy, j1.x, j1.y, j1.z);
            if (power > 0.0f) continue;
            const float G = expf(power);
            const float alpha = fminf(0.99f, __fmul_rn(j1.w, G));
            if (alpha < 1.0f / 255.0f) continue;
            float weight;
            const float dL_dalpha = pair_step(st, alpha, rec[j].q2, weight);
            const uint32_t mt = sm.meta[j];
            const int k = (int)((mt >> 13) & 0x1ffffu) + (ly - (int)((mt >> 4) & 15)) * (int)((mt >> 8) & 31) +
                          (lx - (int)(mt & 15));
            sm.slot[k] = make_float2(G * dL_dalpha, weight);
          }
          if (d < 0) break;
          {  // dense instance d: all lanes together, butterfly reduction, one RED per lane
            const int j = w * 32 + d;
            const float2 jxy = *reinterpret_cast<const float2*>(&rec[j].q0);
            const float4 j1 = rec[j].q1;
            const float dx = jxy.x - pxf, dy = jxy.y - pyf;
            const float power = power_of(dx, dy, j1.x, j1.y, j1.z);
            const float G = expf(power);
            const float alpha = fminf(0.99f, __fmul_rn(j1.w, G));
            const bool active = ((m >> d) & 1u) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
            if (__any_sync(kFull, active)) {
              float gr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
              float g_op = 0.f;
              if (active) {
                const float4 j2 = rec[j].q2;
                float weight;
                const float dL_dalpha = pair_step(st, alpha, j2, weight);
                const float dL_dG = j1.w * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                gr[0] = weight * st.dpx0, gr[1] = weight * st.dpx1, gr[2] = weight * st.dpx2;
                gr[3] = dL_dG * (-gdx * j1.x - gdy * j1.y) * ddelx_dx;
                gr[4] = dL_dG * (-gdy * j1.z - gdx * j1.y) * ddely_dy;
                gr[5] = -0.5f * gdx * dx * dL_dG;
                gr[6] = -0.5f * gdx * dy * dL_dG;
                gr[7] = -0.5f * gdy * dy * dL_dG;
                g_op = G * dL_dalpha;
              }
              const float tsum = transpose_reduce8_warp(gr, lane);
              float o = g_op;
#pragma unroll
              for (int sh = 16; sh > 0; sh >>= 1) o += __shfl_xor_sync(kFull, o, sh);
              if (gptr) atomicAdd(gptr + (size_t)aux[j].id * gstride, lane == 1 ? o : tsum);
            }
          }
          dm &= ~(1u << d);
        }
      }
    }

    __syncthreads();  // (2) every blended pair of a sparse instance has left {s, w} in its slot

    // ---------------- phase C: instance sums its slots into 9 gradient words ----------------
    {
      float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f, g4 = 0.f, g5 = 0.f, g6 = 0.f, g7 = 0.f, gop = 0.f;
      if (sparse) {
        const float fx0 = tx0 + (float)x0;
        for (int y = y0 + half; y <= y1; y += 2) {
          const float dy = xy.y - (ty0 + (float)y);
          float fx = fx0;
          const float2* srow = wslot + sbase + (y - y0) * bw - x0;
          const int prow = y * 16;
          for (int x = x0; x <= x1; ++x, fx += 1.0f) {
            const float2 v = srow[x];
            if (v.x == 0.0f && v.y == 0.0f) continue;
            const float dx = xy.x - fx;
            const float t = q1.w * v.x, tx = t * dx, ty = t * dy;
            g0 = fmaf(v.y, sm.dpx[0][prow + x], g0);
            g1 = fmaf(v.y, sm.dpx[1][prow + x], g1);
            g2 = fmaf(v.y, sm.dpx[2][prow + x], g2);
            g3 += -tx * q1.x - ty * q1.y;
            g4 += -ty * q1.z - tx * q1.y;
            g5 = fmaf(tx, dx, g5);
            g6 = fmaf(tx, dy, g6);
            g7 = fmaf(ty, dy, g7);
            gop += v.x;
          }
        }
      }
      // combine the two halves of the instance
      g0 += __shfl_xor_sync(kFull, g0, 1), g1 += __shfl_xor_sync(kFull, g1, 1), g2 += __shfl_xor_sync(kFull, g2, 1);
      g3 += __shfl_xor_sync(kFull, g3, 1), g4 += __shfl_xor_sync(kFull, g4, 1), g5 += __shfl_xor_sync(kFull, g5, 1);
      g6 += __shfl_xor_sync(kFull, g6, 1), g7 += __shfl_xor_sync(kFull, g7, 1), gop += __shfl_xor_sync(kFull, gop, 1);
      if (sparse && half == 0 &&
          (g0 != 0.f || g1 != 0.f || g2 != 0.f || g3 != 0.f || g4 != 0.f || g5 != 0.f || g6 != 0.f || g7 != 0.f ||
           gop != 0.f)) {
        const size_t id = aux[ja].id;
        atomicAdd(a.dL_dcolor + 3 * id, g0);
        atomicAdd(a.dL_dcolor + 3 * id + 1, g1);
        atomicAdd(a.dL_dcolor + 3 * id + 2, g2);
        atomicAdd(a.dL_dmean2D + 3 * id, g3 * ddelx_dx);
        atomicAdd(a.dL_dmean2D + 3 * id + 1, g4 * ddely_dy);
        atomicAdd(a.dL_dconic + 4 * id, -0.5f * g5);
        atomicAdd(a.dL_dconic + 4 * id + 1, -0.5f * g6);
        atomicAdd(a.dL_dconic + 4 * id + 3, -0.5f * g7);
        atomicAdd(a.dL_dopacity + id, gop);
      }
    }
    if (warp < kWords && lane == 0) sm.dense[warp] = 0;  // read in phase B only; next written after barrier (3)
    __syncthreads();  // (3) slots / meta / dense bits may be rewritten
  }
}

}  // namespace

cudaError_t launch_blend_bwd_mask(int width, int height, const ImageState& img, const BinningState& b,
                                  const float* background, const float* dL_dpix, const float* dL_dacc,
                                  float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                                  cudaStream_t s) {
  const int grid_x = (width + kTile - 1) / kTile, grid_y = (height + kTile - 1) / kTile;
  constexpr int kSmem = (int)sizeof(BwdSmem);
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(blend_bwd_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  BwdArgs a;
  a.ranges = img.ranges, a.tile_order = img.tile_order, a.tile_last = img.tile_last, a.inst = b.inst, a.aux = b.aux;
  a.W = width, a.H = height, a.grid_x = grid_x, a.tiles = grid_x * grid_y;
  a.bg = background, a.final_T = img.final_T, a.n_contrib = img.n_contrib;
  a.dL_dpix = dL_dpix, a.dL_dacc = dL_dacc;
  a.dL_dmean2D = dL_dmean2D, a.dL_dconic = dL_dconic, a.dL_dopacity = dL_dopacity, a.dL_dcolor = dL_dcolor;
  blend_bwd_mask_kernel<<<a.tiles, kThreads, kSmem, s>>>(a);
  return cudaGetLastError();
}

}  // namespace gsr
