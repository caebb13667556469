// Per-block candidate lists: for every 8x4 pixel block of every tile, the sorted instances whose
// conservative {alpha >= 1/255} box contains at least one pixel centre of the block, each with a 32-bit
// coverage mask (bit 8*row + col = that pixel is inside the box). sm_100a.
//
// The reference has no such stage: FORWARD::render / BACKWARD::render walk the whole tile list for every
// pixel and `continue` on pairs below 1/255 (src/cuda_rasterizer/forward.cu:360-375, backward.cu:528-544).
// On NTU-VIRAL-shaped maps a Gaussian reaches alpha >= 1/255 on ~7 pixels of the 256 of its tile, so the
// blend kernels here (blend_fwd_sparse.cu, blend_bwd_sparse.cu) evaluate only (pixel, instance) pairs
// inside the box: a pair outside it is a pair the reference skips, so results are unchanged.
//
// One CTA per tile, heaviest tile first, 256 sorted instances per iteration (thread = instance, the next
// iteration's q0 word already in flight): ballots give every warp its entries per block, one barrier and a
// 64-word exchange give the write positions (so every block list stays in sorted order), and the thread
// writes {index in the tile list, mask} to the lists of the blocks it covers. Block b of a tile whose range
// starts at r and holds n instances owns the n entries from 8*r + b*n: no allocation pass, 64 B of address
// space per instance, touched only as far as the lists are long.
#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kWarps = 8;
constexpr uint32_t kFull = 0xffffffffu;

struct BlockListArgs {
  const uint2* ranges;
  const uint32_t* tile_order;
  const SplatRecord* inst;
  uint2* blist;
  uint32_t* block_count;
  uint4* bdesc;
  unsigned long long* work;  // (pixel, candidate) pairs inside the candidates' boxes: the alphas the blend evaluates
  int grid_x;
};

// Coverage masks of one instance for the 8 blocks of its tile (block = 2*by + bx, 8 px wide, 4 px high).
__device__ __forceinline__ void block_masks(const float4 c, float tx0, float ty0, uint32_t cov[8]) {
  // pixel px is inside iff x - ex <= px <= x + ex (the test the warp-shaped kernels apply per block)
  const float fx0 = fmaxf(ceilf(c.x - c.z), tx0), fx1 = fminf(floorf(c.x + c.z), tx0 + 15.0f);
  const float fy0 = fmaxf(ceilf(c.y - c.w), ty0), fy1 = fminf(floorf(c.y + c.w), ty0 + 15.0f);
  const bool any = (fx0 <= fx1) && (fy0 <= fy1);  // false for NaN and for ex = ey = -inf
  uint32_t xm = 0, ym = 0;
  if (any) {
    const int X0 = (int)(fx0 - tx0), X1 = (int)(fx1 - tx0), Y0 = (int)(fy0 - ty0), Y1 = (int)(fy1 - ty0);
    xm = ((2u << (X1 - X0)) - 1u) << X0;
    ym = ((2u << (Y1 - Y0)) - 1u) << Y0;
  }
  const uint32_t xm8[2] = {xm & 0xffu, (xm >> 8) & 0xffu};
#pragma unroll
  for (int by = 0; by < 4; ++by) {
    // spread the 4 row bits of this block row to bit 0 of 4 bytes: bit k -> bit 8k
    const uint32_t rep = (((ym >> (4 * by)) & 0xfu) * 0x00204081u) & 0x01010101u;
    cov[2 * by] = xm8[0] * rep;
    cov[2 * by + 1] = xm8[1] * rep;
  }
}

__global__ void __launch_bounds__(kWarps * 32) block_lists_kernel(const BlockListArgs a) {
  __shared__ uint32_t counts[2][kWarps * 8];  // [iteration parity][warp * 8 + block]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = (int)a.tile_order[blockIdx.x];
  const uint2 range = a.ranges[tile];
  const int n = (int)(range.y - range.x);
  if (n == 0) {
    if (threadIdx.x < 8) {
      a.block_count[tile * 8 + threadIdx.x] = 0;
      a.bdesc[blockIdx.x * 8 + threadIdx.x] = make_uint4(0u, 0u, 0u, (uint32_t)tile);
    }
    return;
  }
  const float tx0 = (float)((tile % a.grid_x) * kTile), ty0 = (float)((tile / a.grid_x) * kTile);
  const SplatRecord* src = a.inst + range.x;
  uint2* out = a.blist + (size_t)8 * range.x;
  const uint32_t lt = (1u << lane) - 1u;
  const int b_mine = lane & 7, w_lo = lane >> 3;  // this lane sums block b_mine over warps w_lo and w_lo + 4

  uint32_t run = 0;  // entries written so far to block (lane & 7), identical in every warp
  uint32_t pairs = 0;  // this thread's share of the tile's (pixel, candidate) pairs
  const float4 none = make_float4(0.f, 0.f, -INFINITY, -INFINITY);
  float4 q_next = (int)threadIdx.x < n ? src[threadIdx.x].q0 : none;
  for (int base = 0, it = 0; base < n; base += kWarps * 32, ++it) {
    const int i = base + (int)threadIdx.x;
    const float4 q = q_next;
    const int i_next = i + kWarps * 32;
    q_next = i_next < n ? src[i_next].q0 : none;  // in flight while this iteration is ranked and written

    uint32_t cov[8];
    block_masks(q, tx0, ty0, cov);
    uint32_t m[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) m[b] = __ballot_sync(kFull, cov[b] != 0), pairs += __popc(cov[b]);
    uint32_t* cnt = counts[it & 1];
    if (lane < 8) {
      uint32_t mine = m[0];
#pragma unroll
      for (int b = 1; b < 8; ++b) mine = lane == b ? m[b] : mine;
      cnt[warp * 8 + lane] = (uint32_t)__popc(mine);
    }
    __syncthreads();
    // entries of block b_mine written by earlier warps in this iteration, and by all warps
    const uint32_t c0 = cnt[w_lo * 8 + b_mine], c1 = cnt[(w_lo + 4) * 8 + b_mine];
    uint32_t before = (w_lo < warp ? c0 : 0u) + (w_lo + 4 < warp ? c1 : 0u);
    uint32_t total = c0 + c1;
    before += __shfl_xor_sync(kFull, before, 8), total += __shfl_xor_sync(kFull, total, 8);
    before += __shfl_xor_sync(kFull, before, 16), total += __shfl_xor_sync(kFull, total, 16);
    const uint32_t start = run + before;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const uint32_t pos = __shfl_sync(kFull, start, b);
      if (cov[b]) out[(size_t)b * n + pos + __popc(m[b] & lt)] = make_uint2((uint32_t)i, cov[b]);
    }
    run += total;
    // counts[it & 1] is rewritten two iterations from now, behind the next iteration's barrier
  }
  pairs = __reduce_add_sync(kFull, pairs);
  if (lane == 0 && pairs) atomicAdd(a.work, (unsigned long long)pairs);
  if (warp == 0 && lane < 8) {
    a.block_count[tile * 8 + lane] = run;
    a.bdesc[blockIdx.x * 8 + lane] = make_uint4(range.x, (uint32_t)n, run, (uint32_t)tile);
  }
}

}  // namespace

cudaError_t launch_block_lists(int tiles, const ImageState& img, const BinningState& b, int grid_x, cudaStream_t s) {
  BlockListArgs a;
  a.ranges = img.ranges, a.tile_order = img.tile_order, a.inst = b.inst, a.blist = b.blist;
  a.block_count = img.block_count, a.bdesc = img.bdesc, a.grid_x = grid_x;
  a.work = reinterpret_cast<unsigned long long*>(img.summary + 6);
  block_lists_kernel<<<tiles, kWarps * 32, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace gsr
