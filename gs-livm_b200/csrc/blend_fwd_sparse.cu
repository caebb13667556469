// Front-to-back alpha blend of C (colour), D (depth) and S (silhouette), pair-sparse formulation. sm_100a.
//
// Replaces FORWARD::render / renderCUDA<3> (reference src/cuda_rasterizer/forward.cu:291-441).
//
// The reference evaluates every (pixel, instance) pair of a tile: 235 M pairs per 752x480 frame of a
// 1 M-Gaussian NTU-VIRAL-shaped map, of which 8.4 M reach alpha >= 1/255 and 13.9 M lie inside the
// instances' conservative boxes. This kernel evaluates the pairs inside the boxes only:
//  * a warp owns one 8x4 pixel block of a tile, lane = pixel. Its work list is the block's candidate list
//    written by block_lists.cu: {index in the tile's sorted list, 32-bit coverage mask}, in sorted order;
//  * 32 candidates at a time: lane s fetches candidate s (list entry, then its 48-byte record from the
//    sorted instance array; both are register-prefetched one and two groups ahead), parks the record in
//    shared memory, and the 32 coverage masks (lane = candidate, bit = pixel) are transposed across the
//    warp in 5 shuffle steps into 32 per-pixel masks (lane = pixel, bit = candidate);
//  * every lane then walks the set bits of ITS mask in order: exactly the pairs of its pixel that can reach
//    1/255, ILP at a time (the alphas are independent; only the transmittance recurrence is ordered).
//    Lanes no longer wait on each other's pairs: a lane is busy for (pairs of its pixel) steps instead of
//    (instances overlapping the block) steps.
// Pairs outside a box are pairs the reference `continue`s on, and the per-pair operation DAG (which
// product is fused, order of accumulation) mirrors the reference's sm_100a SASS through explicit _rn
// intrinsics, so C, D, S, final_T and n_contrib are bit-identical to blend_fwd.cu and to the reference.
#include <cstdlib>

#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kWarps = 8;
constexpr int kWindow = 4;  // groups of 32 candidates resident per warp (power of two)
constexpr uint32_t kFull = 0xffffffffu;
constexpr int kNever = 0x7fffffff;

struct __align__(16) WarpBuf {
  float4 a[kWindow * 32];  // {mean2D.x, mean2D.y, conic.x, conic.y}
  float4 b[kWindow * 32];  // {conic.z, opacity, r, g}
  float4 c[kWindow * 32];  // {b, depth, bits(index in tile list + 1), -}
  uint32_t mask[kWindow * 32];  // [segment][pixel] candidates of the segment's group that cover the pixel
};

struct FwdArgs {
  const uint2* ranges;
  const uint32_t* tile_order;
  const SplatRecord* inst;
  const uint2* blist;
  const uint32_t* block_count;
  uint32_t* block_used;
  int W, H, grid_x;
  const float* bg;
  float *out_color, *out_depth, *out_acc, *final_T;
  uint32_t *n_contrib, *tile_last;
};

// 32x32 bit transpose across the warp: lane = candidate, bit = pixel  ->  lane = pixel, bit = candidate
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(kFull, x, j);
    x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
  }
  return x;
}

template <int ILP>
__global__ void __launch_bounds__(kWarps * 32) blend_fwd_sparse_kernel(const FwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpBuf& buf = reinterpret_cast<WarpBuf*>(smem_raw)[warp];
  const int tile = (int)a.tile_order[blockIdx.x];
  const uint2 range = a.ranges[tile];
  const int n = (int)(range.y - range.x);
  const int cnt = n ? (int)a.block_count[tile * 8 + warp] : 0;
  const uint2* list = a.blist + (size_t)8 * range.x + (size_t)warp * n;
  const SplatRecord* src = a.inst + range.x;

  const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
  const int px = tile_x * kTile + (warp & 1) * 8 + (lane & 7);
  const int py = tile_y * kTile + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < a.W && py < a.H;
  const float pxf = (float)px, pyf = (float)py;

  float T = 1.0f, A = 0.f, D = 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  uint32_t last_contributor = 0;
  bool done = !inside;
  int last_group = -1;  // group of this pixel's last contributor

  const int ngroups = (cnt + 31) >> 5;
  // register pipeline: list entry two groups ahead of the staging point, record one group ahead
  uint2 ent = make_uint2(0u, 0u), ent_next = make_uint2(0u, 0u);
  float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
  if (lane < cnt) ent = list[lane];
  if (32 + lane < cnt) ent_next = list[32 + lane];
  if (lane < cnt) {
    const float4* q = reinterpret_cast<const float4*>(src + ent.x);
    r0 = q[0], r1 = q[1], r2 = q[2];
  }

  // Lanes walk the list at their own pace inside a window of kWindow groups: `staged` groups have been parked in
  // shared memory (group g in segment g % kWindow, records and per-pixel masks). A lane holds the bits it still
  // owes of group `base` in `lo` and, once that group is staged, of group base + 1 in `hi`.
  uint32_t lo = 0u, hi = 0u;
  bool hi_loaded = false;
  int base = -1, staged = 0;

  while (true) {
    const int low = done ? kNever : (lo ? base : base + 1);  // oldest group this lane may still read
    const int min_low = __reduce_min_sync(kFull, low);       // also orders the window's reads before its re-use
    if (min_low >= ngroups) break;                           // every lane is saturated or has consumed its list
    while (staged < ngroups && staged - min_low < kWindow) {
      const int seg = (staged % kWindow) * 32 + lane;
      buf.a[seg] = make_float4(r0.x, r0.y, r1.x, r1.y);
      buf.b[seg] = make_float4(r1.z, r1.w, r2.x, r2.y);
      buf.c[seg] = make_float4(r2.z, r2.w, __uint_as_float(ent.x + 1u), 0.f);
      buf.mask[seg] = transpose32((staged * 32 + lane < cnt) ? ent.y : 0u, lane);
      ++staged;
      // prefetch: records of the next group to stage, list entries of the one after
      ent = ent_next;
      if (staged * 32 + lane < cnt) {
        const float4* q = reinterpret_cast<const float4*>(src + ent.x);
        r0 = q[0], r1 = q[1], r2 = q[2];
      }
      ent_next = ((staged + 1) * 32 + lane < cnt) ? list[(staged + 1) * 32 + lane] : make_uint2(0u, 0u);
      __syncwarp();
    }

    if (!done) {
      // advance by at most one group per step, branch-free
      if (!hi_loaded && base + 1 < staged) hi = buf.mask[((base + 1) % kWindow) * 32 + lane], hi_loaded = true;
      if (lo == 0u && hi_loaded) {
        lo = hi, hi = 0u, ++base;
        hi_loaded = base + 1 < staged;
        if (hi_loaded) hi = buf.mask[((base + 1) % kWindow) * 32 + lane];
      }
      unsigned long long c = ((unsigned long long)hi << 32) | lo;
      const int seg0 = (base & (kWindow - 1)) * 32;  // base >= 0 whenever c != 0
      int e[ILP], grp[ILP];
      bool valid[ILP];
#pragma unroll
      for (int u = 0; u < ILP; ++u) {
        valid[u] = c != 0ull;
        const int bit = valid[u] ? __ffsll((long long)c) - 1 : 0;
        c &= c - 1ull;
        e[u] = (seg0 + bit) & (kWindow * 32 - 1);
        grp[u] = base + (bit >> 5);
      }
      lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
      float al[ILP];
#pragma unroll
      for (int u = 0; u < ILP; ++u) {
        const float4 qa = buf.a[e[u]];
        const float4 qb = buf.b[e[u]];
        const float dx = __fadd_rn(qa.x, -pxf);
        const float dy = __fadd_rn(qa.y, -pyf);
        const float quad = __fmaf_rn(dx, __fmul_rn(dx, qa.z), __fmul_rn(dy, __fmul_rn(dy, qb.x)));
        const float power = __fmaf_rn(quad, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, qa.w)));
        const float v = fminf(0.99f, __fmul_rn(qb.y, expf(power)));
        // reference: `if (power > 0) continue; ... if (alpha < 1/255) continue;`
        al[u] = (valid[u] && !(power > 0.0f) && !(v < 1.0f / 255.0f)) ? v : -1.0f;
      }
#pragma unroll
      for (int u = 0; u < ILP; ++u) {
        if (al[u] >= 0.0f && !done) {
          const float4 qb = buf.b[e[u]];
          const float4 qc = buf.c[e[u]];
          const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -al[u]));
          if (test_T < 0.0001f) {  // reference: done = true; continue
            done = true;
          } else {
            A = __fmaf_rn(T, al[u], A);
            D = __fmaf_rn(T, __fmul_rn(al[u], qc.y), D);
            C0 = __fmaf_rn(T, __fmul_rn(al[u], qb.z), C0);
            C1 = __fmaf_rn(T, __fmul_rn(al[u], qb.w), C1);
            C2 = __fmaf_rn(T, __fmul_rn(al[u], qc.x), C2);
            T = test_T;
            last_contributor = __float_as_uint(qc.z);
            last_group = grp[u];
          }
        }
      }
    }
  }

  if (inside) {
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
  const int used = __reduce_max_sync(kFull, last_group + 1);
  if (lane == 0) {
    a.block_used[tile * 8 + warp] = (uint32_t)used;  // groups of this block's list the backward has to walk
    if (wmax) atomicMax(&a.tile_last[tile], wmax);   // tile_last is zeroed by order_tiles
  }
}

}  // namespace

cudaError_t launch_blend_fwd_sparse(int width, int height, const ImageState& img, const BinningState& b,
                                    const float* background, float* out_color, float* out_depth, float* out_acc,
                                    cudaStream_t s) {
  const int grid_x = (width + kTile - 1) / kTile, grid_y = (height + kTile - 1) / kTile;
  FwdArgs a;
  a.ranges = img.ranges, a.tile_order = img.tile_order, a.inst = b.inst, a.blist = b.blist;
  a.block_count = img.block_count, a.block_used = img.block_used;
  a.W = width, a.H = height, a.grid_x = grid_x;
  a.bg = background, a.out_color = out_color, a.out_depth = out_depth, a.out_acc = out_acc;
  a.final_T = img.final_T, a.n_contrib = img.n_contrib, a.tile_last = img.tile_last;
  static const int ilp = [] {
    const char* e = getenv("GSR_FWD_ILP");  // tuning knob: alphas evaluated per lane per step (1, 2 or 4)
    return e ? atoi(e) : 4;
  }();
  const int tiles = grid_x * grid_y;
  constexpr int kSmem = (int)sizeof(WarpBuf) * kWarps;
  auto kernel = ilp == 1 ? blend_fwd_sparse_kernel<1> : ilp == 2 ? blend_fwd_sparse_kernel<2> : blend_fwd_sparse_kernel<4>;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  kernel<<<tiles, kWarps * 32, kSmem, s>>>(a);
  return cudaGetLastError();
}

}  // namespace gsr
