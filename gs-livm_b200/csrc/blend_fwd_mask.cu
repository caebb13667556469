// Front-to-back alpha blend of C (colour), D (depth) and S (silhouette) per 16x16 tile, sm_100a:
// the "coverage mask" formulation.
//
// Replaces FORWARD::render / renderCUDA<3> (reference src/cuda_rasterizer/forward.cu:291-441).
//
// Why. On NTU-VIRAL-shaped maps (voxelised surfaces, 16 small Gaussians per 0.2 m voxel) a sorted instance
// reaches alpha >= 1/255 on ~7 of the 256 pixels of its tile: the reference evaluates 235 M (pixel, Gaussian)
// pairs per 752x480 frame to blend 8.4 M of them (tools/workload_stats.py). Pixel-parallel kernels pay for
// every pair a pixel block *might* see; this kernel pays for the pixels an instance *does* cover.
//
// One CTA (256 threads) per tile, batches of 256 sorted instances staged by TMA bulk copies
// (cp.async.bulk -> SASS UBLKCP) into a 3-deep shared-memory ring completing on mbarriers. Per batch:
//   phase A, thread = instance: walk the few pixels of the instance's conservative box of {alpha >= 1/255}
//     clipped to the tile (precomputed per instance by finalize_binning, 16-byte aux record), evaluate
//     `power` with the reference's operation DAG and, where the pair can pass the reference's tests, set bit
//     (instance % 32) of word (instance / 32) of that pixel's coverage mask (shared-memory atomicOr).
//     Instances covering more than kLoopMax pixels are walked by the whole warp, 32 pixels per step;
//   one __syncthreads (masks are double-buffered, so A of batch n+1 overlaps B of batch n);
//   phase B, thread = pixel: walk the set bits of the pixel's 8 mask words in list order and blend exactly
//     like the reference (same tests, same operation DAG, same order), so every skipped pair is a pair the
//     reference `continue`s on and C, D, S, final_T and n_contrib are unchanged.
// Phase A's test is `power` against -(ln(255 * opacity) + margin): conservative, never exact; phase B repeats
// the reference's exact tests.
#include <cstdlib>

#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kThreads = 256;
constexpr int kMB = 256;       // instances per batch
constexpr int kWords = kMB / 32;
constexpr int kRing = 3;       // record ring depth
constexpr int kLoopMax = 32;   // clipped boxes up to this many pixels are walked by their own thread
constexpr uint32_t kFull = 0xffffffffu;

struct __align__(128) FwdSmem {
  SplatRecord rec[kRing][kMB];           // 36 KB
  InstAux aux[kRing][kMB];               // 12 KB
  uint32_t mask[2][kWords][kThreads];    // 16 KB, [word][y * 16 + x]
  uint64_t full[kRing];
};

struct FwdArgs {
  const uint2* ranges;
  const uint32_t* tile_order;
  const SplatRecord* inst;
  const InstAux* aux;
  int W, H, grid_x, tiles;
  const float* bg;
  float *out_color, *out_depth, *out_acc, *final_T;
  uint32_t *n_contrib, *tile_last;
};

// The reference's `power` (forward.cu:357-360) with the operation DAG of its sm_100a SASS.
__device__ __forceinline__ float power_of(float dx, float dy, float ca, float cb, float cc) {
  const float quad = __fmaf_rn(dx, __fmul_rn(dx, ca), __fmul_rn(dy, __fmul_rn(dy, cc)));
  return __fmaf_rn(quad, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, cb)));
}

__global__ void __launch_bounds__(kThreads) blend_fwd_mask_kernel(const FwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int tile = (int)a.tile_order[blockIdx.x];
  const uint2 range = a.ranges[tile];
  const int total = (int)(range.y - range.x);
  const int nbatch = (total + kMB - 1) / kMB;
  const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
  const int tpx0 = tile_x * kTile, tpy0 = tile_y * kTile;
  const float tx0 = (float)tpx0, ty0 = (float)tpy0;

  // this thread's pixel (phase B): a warp owns an 8x4 block
  const int lx = ((tid >> 5) & 1) * 8 + (lane & 7), ly = (tid >> 6) * 4 + (lane >> 3);
  const int mp = ly * 16 + lx;  // index of the pixel's mask words
  const int pxi = tpx0 + lx, pyi = tpy0 + ly;
  const bool inside = pxi < a.W && pyi < a.H;
  const float pxf = (float)pxi, pyf = (float)pyi;

  float T = 1.0f, A = 0.f, D = 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  uint32_t last = 0;
  bool done = !inside;

  for (int w = 0; w < kWords; ++w) sm.mask[0][w][tid] = 0, sm.mask[1][w][tid] = 0;

  auto issue = [&](int b) {  // thread 0: TMA bulk copies of batch b (records + aux)
    const int stage = b % kRing;
    const uint32_t cnt = (uint32_t)min(kMB, total - b * kMB);
    const size_t first = (size_t)range.x + (size_t)b * kMB;
    mbar_expect_tx(&sm.full[stage], cnt * (uint32_t)(sizeof(SplatRecord) + sizeof(InstAux)));
    bulk_g2s(sm.rec[stage], a.inst + first, cnt * (uint32_t)sizeof(SplatRecord), &sm.full[stage]);
    bulk_g2s(sm.aux[stage], a.aux + first, cnt * (uint32_t)sizeof(InstAux), &sm.full[stage]);
  };
  int issued = 0, waited = 0;  // meaningful in thread 0 only
  if (tid == 0) {
    for (int s = 0; s < kRing; ++s) mbar_init(&sm.full[s], 1);
    mbar_fence_init();
    for (; issued < min(nbatch, 2); ++issued) issue(issued);
  }
  __syncthreads();

  for (int b = 0; b < nbatch; ++b) {
    const int stage = b % kRing, buf = b & 1;
    const int cnt = min(kMB, total - b * kMB);
    const SplatRecord* rec = sm.rec[stage];
    const InstAux* aux = sm.aux[stage];
    mbar_wait(&sm.full[stage], (uint32_t)(b / kRing) & 1u);
    if (tid == 0) waited = b + 1;

    // ---------------- phase A: instance -> coverage bits ----------------
    {
      const int j = tid;
      uint32_t box = 0;
      float pmin = 0.f;
      if (j < cnt) {
        const uint2 bx = *reinterpret_cast<const uint2*>(&aux[j]);
        box = bx.x, pmin = __uint_as_float(bx.y);
      }
      const int x0 = box & 15, x1 = (box >> 4) & 15, y0 = (box >> 8) & 15, y1 = (box >> 12) & 15;
      const int area = (box & kBoxValid) ? (x1 - x0 + 1) * (y1 - y0 + 1) : 0;
      uint32_t* mword = sm.mask[buf][j >> 5];
      if (area > 0 && area <= kLoopMax) {
        const uint32_t bit = 1u << (j & 31);
        const float2 xy = *reinterpret_cast<const float2*>(&rec[j].q0);
        const float4 q1 = rec[j].q1;
        float fy = ty0 + (float)y0;
        const float fx0 = tx0 + (float)x0;
        for (int y = y0; y <= y1; ++y, fy += 1.0f) {
          const float dy = __fadd_rn(xy.y, -fy);
          float fx = fx0;
          uint32_t* mrow = mword + y * 16;
          for (int x = x0; x <= x1; ++x, fx += 1.0f) {
            const float power = power_of(__fadd_rn(xy.x, -fx), dy, q1.x, q1.y, q1.z);
            if (!(power > 0.0f) && !(power < pmin)) atomicOr(&mrow[x], bit);
          }
        }
      }
      // large boxes: the whole warp walks them, 32 pixels per step
      uint32_t big = __ballot_sync(kFull, area > kLoopMax);
      while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int jj = (tid & ~31) + src;
        const uint2 bx = *reinterpret_cast<const uint2*>(&aux[jj]);
        const float smin = __uint_as_float(bx.y);
        const int sx0 = bx.x & 15, sy0 = (bx.x >> 8) & 15;
        const int sw = (int)((bx.x >> 4) & 15) - sx0 + 1, sarea = sw * ((int)((bx.x >> 12) & 15) - sy0 + 1);
        const float2 xy = *reinterpret_cast<const float2*>(&rec[jj].q0);
        const float4 s1 = rec[jj].q1;
        const uint32_t sbit = 1u << src;
        for (int i = lane; i < sarea; i += 32) {
          const int yy = i / sw, y = sy0 + yy, x = sx0 + (i - yy * sw);
          const float dy = __fadd_rn(xy.y, -(ty0 + (float)y));
          const float dx = __fadd_rn(xy.x, -(tx0 + (float)x));
          const float power = power_of(dx, dy, s1.x, s1.y, s1.z);
          if (!(power > 0.0f) && !(power < smin)) atomicOr(&mword[y * 16 + x], sbit);
        }
      }
    }

    const int ndone = __syncthreads_count(done);
    // every thread is past phase B of batch b-1: its ring stage may be refilled with batch b+2
    if (tid == 0 && b + 2 < nbatch) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(b + 2);
      issued = b + 3;
    }
    if (ndone == kThreads) break;

    // ---------------- phase B: pixel walks its coverage bits in list order, two at a time ----------------
    const int nw = (cnt + 31) >> 5;
    uint32_t* mcol = &sm.mask[buf][0][mp];
    int w = -1;
    uint32_t m = 0;
    auto next_bit = [&]() -> int {  // batch-local index of the pixel's next covering instance, -1 when exhausted
      while (m == 0) {
        if (++w >= nw) return -1;
        m = mcol[w * kThreads];
        if (m) mcol[w * kThreads] = 0;
      }
      const int bpos = __ffs(m) - 1;
      m &= m - 1;
      return w * 32 + bpos;
    };
    auto alpha_of = [&](int j) -> float {  // reference tests: power > 0 and alpha < 1/255 -> skipped (-1)
      const float2 xy = *reinterpret_cast<const float2*>(&rec[j].q0);
      const float4 q1 = rec[j].q1;
      const float power = power_of(__fadd_rn(xy.x, -pxf), __fadd_rn(xy.y, -pyf), q1.x, q1.y, q1.z);
      const float al = fminf(0.99f, __fmul_rn(q1.w, expf(power)));
      return (!(power > 0.0f) && !(al < 1.0f / 255.0f)) ? al : -1.0f;
    };
    auto apply = [&](int j, float al) {
      if (al < 0.0f || done) return;
      const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -al));
      if (test_T < 0.0001f) {  // forward.cu:372-376
        done = true;
        return;
      }
      const float4 q2 = rec[j].q2;
      A = __fmaf_rn(T, al, A);
      D = __fmaf_rn(T, __fmul_rn(al, q2.w), D);
      C0 = __fmaf_rn(T, __fmul_rn(al, q2.x), C0);
      C1 = __fmaf_rn(T, __fmul_rn(al, q2.y), C1);
      C2 = __fmaf_rn(T, __fmul_rn(al, q2.z), C2);
      T = test_T;
      last = (uint32_t)(b * kMB + j + 1);
    };
    while (!done) {
      const int j0 = next_bit();
      if (j0 < 0) break;
      const int j1 = next_bit();
      const float al0 = alpha_of(j0);
      const float al1 = j1 >= 0 ? alpha_of(j1) : -1.0f;
      apply(j0, al0);
      apply(j1, al1);
    }
    for (++w; w < nw; ++w) mcol[w * kThreads] = 0;  // words this pixel did not get to (it is done)
  }

  // never leave with a bulk copy still in flight towards this CTA's shared memory
  if (tid == 0)
    for (int o = waited; o < issued; ++o) mbar_wait(&sm.full[o % kRing], (uint32_t)(o / kRing) & 1u);

  if (inside) {
    const size_t p = (size_t)a.W * pyi + pxi;
    const size_t plane = (size_t)a.W * a.H;
    a.final_T[p] = T;
    a.n_contrib[p] = last;
    a.out_color[p] = __fmaf_rn(__ldg(a.bg + 0), T, C0);
    a.out_color[plane + p] = __fmaf_rn(__ldg(a.bg + 1), T, C1);
    a.out_color[2 * plane + p] = __fmaf_rn(__ldg(a.bg + 2), T, C2);
    a.out_depth[p] = D;
    a.out_acc[p] = A;
  }
  const uint32_t wmax = __reduce_max_sync(kFull, last);
  if (lane == 0 && wmax) atomicMax(&a.tile_last[tile], wmax);  // tile_last is zeroed by order_tiles
}

}  // namespace

cudaError_t launch_blend_fwd_mask(int width, int height, const ImageState& img, const BinningState& b,
                                  const float* background, float* out_color, float* out_depth, float* out_acc,
                                  cudaStream_t s) {
  const int grid_x = (width + kTile - 1) / kTile, grid_y = (height + kTile - 1) / kTile;
  constexpr int kSmem = (int)sizeof(FwdSmem);
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(blend_fwd_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  FwdArgs a;
  a.ranges = img.ranges, a.tile_order = img.tile_order, a.inst = b.inst, a.aux = b.aux;
  a.W = width, a.H = height, a.grid_x = grid_x, a.tiles = grid_x * grid_y;
  a.bg = background, a.out_color = out_color, a.out_depth = out_depth, a.out_acc = out_acc;
  a.final_T = img.final_T, a.n_contrib = img.n_contrib, a.tile_last = img.tile_last;
  blend_fwd_mask_kernel<<<a.tiles, kThreads, kSmem, s>>>(a);
  return cudaGetLastError();
}

}  // namespace gsr
