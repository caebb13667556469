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
#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kWarps = 8;
constexpr int kFineCap = 384;  // FINE candidates: the first kFineCap tiles of tile_order (longest lists)
constexpr uint32_t kFull = 0xffffffffu;

struct __align__(16) WarpBuf {
  float4 a[32];  // {mean2D.x, mean2D.y, conic.x, conic.y}
  float4 b[32];  // {conic.z, opacity, r, g}
  float4 c[32];  // {b, depth, bits(index in tile list + 1), -}
};

struct FwdArgs {
  const uint2* ranges;
  const uint32_t* tile_order;
  const SplatRecord* inst;
  const uint2* blist;
  const uint4* bdesc;
  uint32_t* last_group;
  uint32_t* bhits;
  uint32_t* tail_hits;
  int W, H, grid_x, heavy, fine_tiles;
  const float* bg;
  float *out_color, *out_depth, *out_acc, *final_T;
  uint32_t *n_contrib, *tile_last;
  unsigned long long* work;  // pairs blended, accumulated per warp (bench.py's executed-work figure)
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

// alpha of one (pixel, candidate) pair, in the reference's operation order (forward.cu:358-375); -1 = skipped
__device__ __forceinline__ float pair_alpha(const float4 qa, const float4 qb, float pxf, float pyf, bool valid) {
  const float dx = __fadd_rn(qa.x, -pxf);
  const float dy = __fadd_rn(qa.y, -pyf);
  const float quad = __fmaf_rn(dx, __fmul_rn(dx, qa.z), __fmul_rn(dy, __fmul_rn(dy, qb.x)));
  const float power = __fmaf_rn(quad, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, qa.w)));
  const float v = fminf(0.99f, __fmul_rn(qb.y, expf(power)));
  // reference: `if (power > 0) continue; ... if (alpha < 1/255) continue;`
  return (valid && !(power > 0.0f) && !(v < 1.0f / 255.0f)) ? v : -1.0f;
}

// One warp, one 8x4 pixel block of a tile (COARSE, SLOTS = 1: lane = pixel, ILP pairs per step) or one 4x2 quarter
// `sub` of it (FINE, SLOTS = 4: four lanes per pixel). FINE is for the tiles with the longest lists, whose
// per-pixel chains of ~1 000 dependent pairs would otherwise outlast the rest of the frame: the four lanes of a
// pixel evaluate the alphas of four consecutive candidates of the pixel at once, exchange them with shuffles,
// and all replay the ordered transmittance recurrence; each lane accumulates one output channel (slot 0: S and
// D, slots 1-3: R, G, B) so every channel still sees the reference's exact sequence of operations.
template <int SLOTS, int ILP>
__device__ __forceinline__ void blend_block(const FwdArgs& a, WarpBuf& buf, const uint4 desc, int block, int sub, int lane) {
  // desc = {range start, list length, candidates of this block, tile} (block_lists.cu)
  const int tile = (int)desc.w, n = (int)desc.y, cnt = (int)desc.z;
  const uint2* list = a.blist + (size_t)8 * desc.x + (size_t)block * n;
  const SplatRecord* src = a.inst + desc.x;
  uint32_t* hits = a.bhits + (size_t)8 * desc.x + (size_t)block * n;  // full groups; the partial last one: tail_hits

  const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
  const int pix = SLOTS == 1 ? lane : (lane & 7), slot = SLOTS == 1 ? 0 : (lane >> 3);
  // pixel inside the 8x4 block, and the bit that stands for it in the coverage masks
  const int bx = SLOTS == 1 ? (pix & 7) : 4 * (sub & 1) + (pix & 3);
  const int by = SLOTS == 1 ? (pix >> 3) : 2 * (sub >> 1) + (pix >> 2);
  const int px = tile_x * kTile + (block & 1) * 8 + bx;
  const int py = tile_y * kTile + (block >> 1) * 4 + by;
  const bool inside = px < a.W && py < a.H;
  const float pxf = (float)px, pyf = (float)py;

  float T = 1.0f, A = 0.f, acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;  // COARSE: D, R, G, B; FINE: acc0 = own channel
  uint32_t last_contributor = 0;
  uint32_t n_hit = 0;  // this lane's blended pairs (FINE: slot 0 counts)
  bool done = !inside;
  int last_group = 0;  // 1 + group of this pixel's last contributor
  // FINE: which word of a parked record this lane's channel is (slot 0: depth, 1-3: r, g, b)
  const float* chan = SLOTS == 1 ? nullptr
                                 : (slot == 0 ? &buf.c[0].y : slot == 1 ? &buf.b[0].z : slot == 2 ? &buf.b[0].w : &buf.c[0].x);

  const int ngroups = (cnt + 31) >> 5;
  // register pipeline: list entry two groups ahead, record one group ahead
  uint2 ent = make_uint2(0u, 0u), ent_next = make_uint2(0u, 0u);
  float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
  // FINE: the four quarters of a block walk the same list; a quarter fetches only the records whose mask touches it
  // (the records of a 10 000-instance tile are ~1 MB, and an SM's share of the memory system is ~50 GB/s)
  const uint32_t mine_bits = SLOTS == 1 ? 0xffffffffu : (0x0f0fu << (16 * (sub >> 1) + 4 * (sub & 1)));
  if (lane < cnt) ent = list[lane];
  if (32 + lane < cnt) ent_next = list[32 + lane];
  if (lane < cnt && (ent.y & mine_bits)) {
    const float4* q = reinterpret_cast<const float4*>(src + ent.x);
    r0 = q[0], r1 = q[1], r2 = q[2];
  }

  for (int g = 0; g < ngroups; ++g) {
    // park this group's records; the vote at the end of the previous iteration ordered its reads
    buf.a[lane] = make_float4(r0.x, r0.y, r1.x, r1.y);
    buf.b[lane] = make_float4(r1.z, r1.w, r2.x, r2.y);
    buf.c[lane] = make_float4(r2.z, r2.w, __uint_as_float(ent.x + 1u), 0.f);
    uint32_t x = (g * 32 + lane < cnt) ? ent.y : 0u;
    if (SLOTS == 4) {  // this quarter's 4x2 bits, replicated so that all four lanes of a pixel receive its mask
      const int sh = 16 * (sub >> 1) + 4 * (sub & 1);
      x = (((x >> sh) & 0xfu) | (((x >> (sh + 8)) & 0xfu) << 4)) * 0x01010101u;
    }
    __syncwarp();

    // prefetch: records of group g+1, list entries of group g+2
    ent = ent_next;
    if ((g + 1) * 32 + lane < cnt && (ent.y & mine_bits)) {
      const float4* q = reinterpret_cast<const float4*>(src + ent.x);
      r0 = q[0], r1 = q[1], r2 = q[2];
    }
    ent_next = ((g + 2) * 32 + lane < cnt) ? list[(g + 2) * 32 + lane] : make_uint2(0u, 0u);

    x = transpose32(x, lane);  // every lane takes part, saturated or not
    uint32_t cand = done ? 0u : x;
    const uint32_t before = last_contributor;
    uint32_t hm = 0u;  // candidates of this group blended into this pixel

    if (SLOTS == 1) {
      while (cand) {  // lanes leave as their own pixel runs out of pairs
        int e[ILP];
        float al[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
          const bool valid = cand != 0u;
          e[u] = valid ? __ffs((int)cand) - 1 : 0;
          cand &= cand - 1u;
          al[u] = pair_alpha(buf.a[e[u]], buf.b[e[u]], pxf, pyf, valid);
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) {  // ordered recurrence, branch-free: only T * (1 - alpha) is on the chain
          const float4 qb = buf.b[e[u]];
          const float4 qc = buf.c[e[u]];
          const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -al[u]));
          const bool on = al[u] >= 0.0f && !done;
          const bool kill = on && test_T < 0.0001f;  // reference: done = true; continue
          const bool ok = on && !kill;
          A = ok ? __fmaf_rn(T, al[u], A) : A;
          acc0 = ok ? __fmaf_rn(T, __fmul_rn(al[u], qc.y), acc0) : acc0;
          acc1 = ok ? __fmaf_rn(T, __fmul_rn(al[u], qb.z), acc1) : acc1;
          acc2 = ok ? __fmaf_rn(T, __fmul_rn(al[u], qb.w), acc2) : acc2;
          acc3 = ok ? __fmaf_rn(T, __fmul_rn(al[u], qc.x), acc3) : acc3;
          T = ok ? test_T : T;
          last_contributor = ok ? __float_as_uint(qc.z) : last_contributor;
          hm |= ok ? (1u << e[u]) : 0u;
          done = done || kill;
        }
        if (done) cand = 0u;
      }
    } else {
      while (__any_sync(kFull, cand != 0u)) {  // the four lanes of a pixel hold the same mask
        int e[4];
        bool valid[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          valid[j] = cand != 0u;
          e[j] = valid[j] ? __ffs((int)cand) - 1 : 0;
          cand &= cand - 1u;
        }
        const int mine = slot == 0 ? e[0] : slot == 1 ? e[1] : slot == 2 ? e[2] : e[3];
        const bool vmine = slot == 0 ? valid[0] : slot == 1 ? valid[1] : slot == 2 ? valid[2] : valid[3];
        const float al = pair_alpha(buf.a[mine], buf.b[mine], pxf, pyf, vmine);
        float alj[4], val[4];
        uint32_t idx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          alj[j] = __shfl_sync(kFull, al, j * 8 + pix);
          val[j] = chan[4 * e[j]];
          idx[j] = __float_as_uint(buf.c[e[j]].z);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // ordered recurrence, branch-free: only T * (1 - alpha) is on the chain
          const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -alj[j]));
          const bool on = alj[j] >= 0.0f && !done;
          const bool kill = on && test_T < 0.0001f;
          const bool ok = on && !kill;
          A = ok ? __fmaf_rn(T, alj[j], A) : A;
          acc0 = ok ? __fmaf_rn(T, __fmul_rn(alj[j], val[j]), acc0) : acc0;
          T = ok ? test_T : T;
          last_contributor = ok ? idx[j] : last_contributor;
          hm |= ok ? (1u << e[j]) : 0u;
          done = done || kill;
        }
        if (done) cand = 0u;
      }
    }
    if (slot == 0) {  // the backward walks exactly the blended pairs: word = pixel (bit order of the coverage masks)
      n_hit += __popc(hm);
      uint32_t* dst = ((g + 1) * 32 <= cnt) ? hits + g * 32 : a.tail_hits + (size_t)(tile * 8 + block) * 32;
      dst[8 * by + bx] = hm;
    }
    if (last_contributor != before) last_group = g + 1;
    if (__all_sync(kFull, done)) break;  // also orders this group's shared-memory reads before the next stores
  }

  if (inside) {
    const size_t p = (size_t)a.W * py + px;
    const size_t plane = (size_t)a.W * a.H;
    if (SLOTS == 1) {
      a.final_T[p] = T;
      a.n_contrib[p] = last_contributor;
      a.last_group[p] = (uint32_t)last_group;
      a.out_color[p] = __fmaf_rn(__ldg(a.bg + 0), T, acc1);
      a.out_color[plane + p] = __fmaf_rn(__ldg(a.bg + 1), T, acc2);
      a.out_color[2 * plane + p] = __fmaf_rn(__ldg(a.bg + 2), T, acc3);
      a.out_depth[p] = acc0;
      a.out_acc[p] = A;
    } else if (slot == 0) {
      a.final_T[p] = T;
      a.n_contrib[p] = last_contributor;
      a.last_group[p] = (uint32_t)last_group;
      a.out_depth[p] = acc0;
      a.out_acc[p] = A;
    } else {
      a.out_color[(size_t)(slot - 1) * plane + p] = __fmaf_rn(__ldg(a.bg + (slot - 1)), T, acc0);
    }
  }
  const uint32_t wmax = __reduce_max_sync(kFull, last_contributor);
  if (lane == 0 && wmax) atomicMax(&a.tile_last[tile], wmax);  // tile_last is zeroed by order_tiles
  const uint32_t w_hit = __reduce_add_sync(kFull, n_hit);
  if (lane == 0 && w_hit) atomicAdd(a.work, (unsigned long long)w_hit);
}

// grid = 4*fine_tiles FINE candidates (tile_order[b/4], two blocks b%4 per CTA) followed by `tiles` COARSE candidates;
// a CTA whose tile belongs to the other shape exits at once. The longest lists therefore start first.
template <int ILP>
__global__ void __launch_bounds__(kWarps * 32) blend_fwd_sparse_kernel(const FwdArgs a) {
  __shared__ WarpBuf bufs[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = (int)blockIdx.x;
  if (b < 4 * a.fine_tiles) {
    const int block = 2 * (b & 3) + (warp >> 2);
    const uint4 d = a.bdesc[(b >> 2) * 8 + block];
    if ((int)d.y < a.heavy) return;
    blend_block<4, 1>(a, bufs[warp], d, block, warp & 3, lane);
  } else {
    const int pos = b - 4 * a.fine_tiles;
    const uint4 d = a.bdesc[pos * 8 + warp];
    // FINE CTAs exist for the first fine_tiles entries of the order only: a heavy tile beyond them is done here
    if ((int)d.y >= a.heavy && pos < a.fine_tiles) return;
    blend_block<1, ILP>(a, bufs[warp], d, warp, 0, lane);
  }
}

}  // namespace

// List length from which a tile's warps take the FINE shape (Tuning::heavy_fwd / heavy_bwd). The hit masks and
// last_group the forward leaves are laid out per pixel, independent of the shape, so the two kernels need not agree:
// the forward gains from FINE a little earlier (3072) than the backward (4096), whose FINE warps also pay four
// phase-C passes per group.

cudaError_t launch_blend_fwd_sparse(int width, int height, const ImageState& img, const BinningState& b,
                                    const float* background, float* out_color, float* out_depth, float* out_acc,
                                    cudaStream_t s) {
  const int grid_x = (width + kTile - 1) / kTile, grid_y = (height + kTile - 1) / kTile;
  const int tiles = grid_x * grid_y;
  FwdArgs a;
  a.ranges = img.ranges, a.tile_order = img.tile_order, a.inst = b.inst, a.blist = b.blist;
  a.bdesc = img.bdesc, a.last_group = img.last_group;
  a.bhits = b.bhits, a.tail_hits = img.tail_hits;
  a.W = width, a.H = height, a.grid_x = grid_x;
  const Tuning t = tuning();
  a.heavy = t.heavy_fwd;
  a.fine_tiles = tiles < kFineCap ? tiles : kFineCap;  // only the longest kFineCap lists can be FINE
  a.bg = background, a.out_color = out_color, a.out_depth = out_depth, a.out_acc = out_acc;
  a.final_T = img.final_T, a.n_contrib = img.n_contrib, a.tile_last = img.tile_last;
  a.work = reinterpret_cast<unsigned long long*>(img.summary + 8);
  const int ilp = t.fwd_ilp;  // alphas evaluated per lane per step, COARSE shape (1, 2 or 4)
  const int grid = 4 * a.fine_tiles + tiles;
  if (ilp == 1) blend_fwd_sparse_kernel<1><<<grid, kWarps * 32, 0, s>>>(a);
  else if (ilp == 2) blend_fwd_sparse_kernel<2><<<grid, kWarps * 32, 0, s>>>(a);
  else blend_fwd_sparse_kernel<4><<<grid, kWarps * 32, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace gsr
