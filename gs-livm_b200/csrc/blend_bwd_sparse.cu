// Back-to-front backward of the tile blend, pair-sparse formulation: dL/dC, dL/dS -> per-Gaussian
// dL/d{mean2D, conic, opacity, colour}. sm_100a.
//
// Replaces BACKWARD::render / renderCUDA<3> (reference src/cuda_rasterizer/backward.cu:438-603,675-713).
// Like the reference, NO depth gradient is produced (backward.cu:567-571, rasterizer.cu:77-80).
//
// The forward (blend_fwd_sparse.cu) leaves, per 8x4 pixel block and per group of 32 candidates of the block's
// list (block_lists.cu), one word per pixel: which candidates of the group it blended into that pixel. Those
// are exactly the pairs the reference's backward does not `continue` on (backward.cu:528-544), so this kernel
// walks them and nothing else, groups in reverse from each pixel's last one (last_group):
//  B. lane = PIXEL walks the set bits of its hit word from the top: re-evaluates alpha, advances the
//     reference's reverse recurrence (T, accumulated colour behind the pair) and leaves the only two
//     numbers of the pair that depend on that recurrence in a shared-memory slot [candidate][pixel]:
//     u = G * dL/dalpha and w = alpha * T;
//  C. lane = CANDIDATE walks the set bits of the transposed hit words (its pixels), reads the slots and
//     accumulates the 9 sums of its Gaussian in registers; every gradient is linear in u or w:
//        dL/dcolour_c += w dL/dC_c            dL/dopacity += u
//        dL/dconic    += -1/2 (op u) {dx^2, dx dy, dy^2}
//        dL/dmean2D   += -(op u) {cx dx + cy dy, cz dy + cy dx} {W/2, H/2}
//     and leaves with 9 RED.ADD.F32 per (Gaussian, block) -- no shuffle butterfly per pair, no atomics per
//     pair (the reference issues 9 same-address atomics per blended pair). Records of candidates that were
//     blended nowhere in the block are never fetched.
// The sums are the reference's sums in another order and association (as between two runs of the
// reference itself): compared at 1e-4 relative, not bitwise.
#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kWarps = 4;  // 10.5 KB of shared memory per warp: 4-warp CTAs pack an SM more tightly than 8-warp ones
constexpr int kFineCap = 384;  // FINE candidates: the first kFineCap tiles of tile_order (longest lists)
constexpr uint32_t kFull = 0xffffffffu;
constexpr int kSlotStride = 33;  // [candidate][pixel] slots, padded against bank conflicts (FINE uses 9 of them)

struct __align__(16) WarpBuf {
  float4 a[32];  // {mean2D.x, mean2D.y, conic.x, conic.y}
  float4 b[32];  // {conic.z, opacity, r, g}
  float cb[32];  // b
  float4 pix[32];  // per pixel {dL/dC_r, dL/dC_g, dL/dC_b, -}
  float2 pxy[32];  // per pixel its coordinates as floats (phase C: no int -> float conversions per pair)
  float2 slot[32 * kSlotStride];
};

struct BwdArgs {
  const uint2* ranges;
  const uint32_t* tile_order;
  const SplatRecord* inst;
  const uint32_t* point_list;
  const uint2* blist;
  const uint4* bdesc;
  const uint32_t* bhits;
  const uint32_t* tail_hits;
  const uint32_t* last_group;
  int W, H, grid_x, heavy, fine_tiles;
  const float* bg;
  const float* final_T;
  const float *dL_dpix, *dL_dacc;
  float *dL_dmean2D, *dL_dconic, *dL_dopacity, *dL_dcolor;
};

__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(kFull, x, j);
    x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
  }
  return x;
}

// Per-pixel state of the reverse traversal (reference backward.cu:480-501).
struct PixelState {
  float T, T_final, acc0, acc1, acc2, acc_s, lc0, lc1, lc2, last_s, last_alpha;
  float dpx0, dpx1, dpx2, dacc, bg_dot;
};

// Advance the recurrence over one blended pair (reference backward.cu:546-585) and return dL/dalpha of the pair.
// rgb = the candidate's colour.
__device__ __forceinline__ float pair_recurrence(PixelState& s, float alpha, float r, float g, float b) {
  // one reciprocal serves both divisions by (1 - alpha) of the reference (:546, :581). MUFU.RCP (1 ulp; 1 - alpha is in
  // [0.01, 1]) instead of the IEEE sequence: gradients are compared at 1e-4, and the kernel is issue-bound
  float inv;  // (div.approx would add its |y| > 2^126 scaling sequence: ~12 instructions per pair)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.f - alpha));
  s.T = s.T * inv;
  const float keep = 1.f - s.last_alpha;
  s.acc0 = s.last_alpha * s.lc0 + keep * s.acc0;
  s.acc1 = s.last_alpha * s.lc1 + keep * s.acc1;
  s.acc2 = s.last_alpha * s.lc2 + keep * s.acc2;
  s.acc_s = s.last_alpha * s.last_s + keep * s.acc_s;
  s.lc0 = r, s.lc1 = g, s.lc2 = b, s.last_s = 1.0f;
  float dL_dalpha = (r - s.acc0) * s.dpx0 + (g - s.acc1) * s.dpx1 + (b - s.acc2) * s.dpx2;
  dL_dalpha += (1.0f - s.acc_s) * s.dacc;
  dL_dalpha *= s.T;
  s.last_alpha = alpha;
  dL_dalpha += (-s.T_final * inv) * s.bg_dot;
  return dL_dalpha;
}

// exp / alpha of a blended pair (the pair is known to pass its tests: no threshold is decided here, so the backward
// can use the two-instruction ex2 path -- relative error ~2e-7 for power in [-5.6, 0] -- where the forward needs expf's
// exact bits)
__device__ __forceinline__ void pair_alpha(const float4 qa, const float4 qb, float pxf, float pyf, float& alpha, float& G) {
  const float dx = __fadd_rn(qa.x, -pxf), dy = __fadd_rn(qa.y, -pyf);
  const float power = __fmaf_rn(__fmaf_rn(dx, __fmul_rn(dx, qa.z), __fmul_rn(dy, __fmul_rn(dy, qb.x))), -0.5f,
                                -__fmul_rn(dy, __fmul_rn(dx, qa.w)));
  G = __expf(power);
  alpha = fminf(0.99f, __fmul_rn(qb.y, G));
}

// One warp, one 8x4 pixel block of a tile (COARSE, SLOTS = 1) or one 4x2 quarter `sub` of it with four lanes per
// pixel (FINE, SLOTS = 4; see blend_fwd_sparse.cu): the four lanes evaluate exp/alpha of four consecutive
// pairs of their pixel at once, all replay the ordered recurrence, and the lane that evaluated a pair leaves
// its (u, w) in the pair's slot.
template <int SLOTS, int ILP>
__device__ __forceinline__ void blend_bwd_block(const BwdArgs& a, WarpBuf& buf, const uint4 desc, int block, int sub,
                                                int lane) {
  constexpr int STRIDE = SLOTS == 1 ? kSlotStride : 9;  // slot row stride
  // desc = {range start, list length, candidates of this block, tile} (block_lists.cu)
  const int tile = (int)desc.w, n = (int)desc.y, cnt = (int)desc.z;
  if (cnt == 0) return;
  const uint2* list = a.blist + (size_t)8 * desc.x + (size_t)block * n;
  const uint32_t* hits = a.bhits + (size_t)8 * desc.x + (size_t)block * n;
  const uint32_t* tail = a.tail_hits + (size_t)(tile * 8 + block) * 32;
  const SplatRecord* src = a.inst + desc.x;
  const uint32_t* ids = a.point_list + desc.x;

  const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
  const int pix = SLOTS == 1 ? lane : (lane & 7), slot = SLOTS == 1 ? 0 : (lane >> 3);
  // pixel inside the 8x4 block (= its bit in the coverage / its word in the hit masks), origin of this warp's pixels
  const int bx = SLOTS == 1 ? (pix & 7) : 4 * (sub & 1) + (pix & 3);
  const int by = SLOTS == 1 ? (pix >> 3) : 2 * (sub >> 1) + (pix >> 2);
  const int wx = tile_x * kTile + (block & 1) * 8 + (SLOTS == 1 ? 0 : 4 * (sub & 1));
  const int wy = tile_y * kTile + (block >> 1) * 4 + (SLOTS == 1 ? 0 : 2 * (sub >> 1));
  const int px = tile_x * kTile + (block & 1) * 8 + bx, py = tile_y * kTile + (block >> 1) * 4 + by;
  const int word = 8 * by + bx;
  const bool inside = px < a.W && py < a.H;
  const float pxf = (float)px, pyf = (float)py;
  const size_t p = (size_t)a.W * py + px;
  const size_t plane = (size_t)a.W * a.H;
  const int lg = inside ? (int)a.last_group[p] : 0;  // groups [0, lg) hold this pixel's pairs
  const int used = __reduce_max_sync(kFull, lg);
  if (used == 0) return;  // nothing was blended into this warp's pixels

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
  if (slot == 0) buf.pix[pix] = make_float4(st.dpx0, st.dpx1, st.dpx2, 0.f), buf.pxy[pix] = make_float2(pxf, pyf);
  const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;
  const float wxf = (float)wx, wyf = (float)wy;

  // hit word of this lane's pixel for group g (zero beyond the pixel's own last group)
  auto hit_word = [&](int g) -> uint32_t {
    if (g < 0 || g >= lg) return 0u;
    return ((g + 1) * 32 <= cnt) ? hits[g * 32 + word] : tail[word];
  };
  // candidate-major view of a group's hit words: lane = candidate, bit = pixel of this warp
  auto by_candidate = [&](uint32_t h) -> uint32_t {
    const uint32_t t = transpose32(h, lane);
    return SLOTS == 1 ? t : (t & 0xffu);
  };

  // Register pipeline, groups used-1 .. 0: hit words two groups ahead; one group ahead their candidate-major view
  // and, for the candidates that were blended somewhere in this warp's pixels, the list entry -> record and id.
  uint32_t h_cur = hit_word(used - 1), h_next = hit_word(used - 2);
  uint32_t cm_next = by_candidate(h_cur);
  float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
  uint32_t gid = 0;
  if (cm_next) {
    const uint32_t at = list[(used - 1) * 32 + lane].x;
    const float4* q = reinterpret_cast<const float4*>(src + at);
    r0 = q[0], r1 = q[1], r2 = q[2];
    gid = ids[at];
  }

  for (int g = used - 1; g >= 0; --g) {
    buf.a[lane] = make_float4(r0.x, r0.y, r1.x, r1.y);
    buf.b[lane] = make_float4(r1.z, r1.w, r2.x, r2.y);
    buf.cb[lane] = r2.z;
    const uint32_t cm = cm_next, my_id = gid;
    uint32_t hh = h_cur;
    // this candidate's own parameters for phase C (before the pipeline registers move on)
    const float mx = r0.x, my = r0.y, cx = r1.x, cy = r1.y, cz = r1.z, op = r1.w;
    __syncwarp();

    h_cur = h_next;
    h_next = hit_word(g - 2);
    cm_next = by_candidate(h_cur);
    if (cm_next) {  // g >= 1 here: h_cur is zero below group 0
      const uint32_t at = list[(g - 1) * 32 + lane].x;
      const float4* q = reinterpret_cast<const float4*>(src + at);
      r0 = q[0], r1 = q[1], r2 = q[2];
      gid = ids[at];
    }

    // ---- phase B: lane = pixel, pairs from the back ----
    if (SLOTS == 1) {
      while (hh) {
        int e[ILP];
        bool valid[ILP];
        float al[ILP], G[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
          valid[u] = hh != 0u;
          e[u] = valid[u] ? 31 - __clz((int)hh) : 0;
          hh &= ~(1u << e[u]);
          pair_alpha(buf.a[e[u]], buf.b[e[u]], pxf, pyf, al[u], G[u]);
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
          if (valid[u]) {  // reference backward.cu:546-600
            const float4 qb = buf.b[e[u]];
            const float dL_dalpha = pair_recurrence(st, al[u], qb.z, qb.w, buf.cb[e[u]]);
            buf.slot[e[u] * STRIDE + pix] = make_float2(G[u] * dL_dalpha, al[u] * st.T);
          }
        }
      }
    } else {
      while (__any_sync(kFull, hh != 0u)) {  // the four lanes of a pixel hold the same word
        int e[4];
        bool valid[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          valid[j] = hh != 0u;
          e[j] = valid[j] ? 31 - __clz((int)hh) : 0;
          hh &= ~(1u << e[j]);
        }
        const int mine = slot == 0 ? e[0] : slot == 1 ? e[1] : slot == 2 ? e[2] : e[3];
        const bool vmine = slot == 0 ? valid[0] : slot == 1 ? valid[1] : slot == 2 ? valid[2] : valid[3];
        float al, G;
        pair_alpha(buf.a[mine], buf.b[mine], pxf, pyf, al, G);
        float2 out = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float alj = __shfl_sync(kFull, al, j * 8 + pix);
          if (valid[j]) {
            const float4 qb = buf.b[e[j]];
            const float dL_dalpha = pair_recurrence(st, alj, qb.z, qb.w, buf.cb[e[j]]);
            if (slot == j) out = make_float2(G * dL_dalpha, alj * st.T);
          }
        }
        if (vmine) buf.slot[mine * STRIDE + pix] = out;
      }
    }
    __syncwarp();

    // ---- phase C: lane = candidate ----
    if (cm) {
      float s_op = 0.f, s_x = 0.f, s_y = 0.f, s_xx = 0.f, s_xy = 0.f, s_yy = 0.f, s_c0 = 0.f, s_c1 = 0.f, s_c2 = 0.f;
      uint32_t todo = cm;
      const float2* row = buf.slot + lane * STRIDE;
      while (todo) {
        const int k = __ffs((int)todo) - 1;
        todo &= todo - 1u;
        const float2 uw = row[k];
        const float4 dp = buf.pix[k];
        const float2 pq = buf.pxy[k];
        const float dx = mx - pq.x, dy = my - pq.y;
        const float v = op * uw.x;
        const float vdx = v * dx, vdy = v * dy;
        s_op += uw.x;
        s_x += vdx, s_y += vdy;
        s_xx = fmaf(vdx, dx, s_xx), s_xy = fmaf(vdx, dy, s_xy), s_yy = fmaf(vdy, dy, s_yy);
        s_c0 = fmaf(uw.y, dp.x, s_c0), s_c1 = fmaf(uw.y, dp.y, s_c1), s_c2 = fmaf(uw.y, dp.z, s_c2);
      }
      const size_t id = my_id;
      atomicAdd(a.dL_dcolor + 3 * id + 0, s_c0);
      atomicAdd(a.dL_dcolor + 3 * id + 1, s_c1);
      atomicAdd(a.dL_dcolor + 3 * id + 2, s_c2);
      atomicAdd(a.dL_dmean2D + 3 * id + 0, -(cx * s_x + cy * s_y) * ddelx_dx);
      atomicAdd(a.dL_dmean2D + 3 * id + 1, -(cz * s_y + cy * s_x) * ddely_dy);
      atomicAdd(a.dL_dconic + 4 * id + 0, -0.5f * s_xx);
      atomicAdd(a.dL_dconic + 4 * id + 1, -0.5f * s_xy);
      atomicAdd(a.dL_dconic + 4 * id + 3, -0.5f * s_yy);
      atomicAdd(a.dL_dopacity + id, s_op);
    }
    __syncwarp();  // slots and records are rewritten by the next group
  }
}

// grid = 8*fine_tiles FINE candidates (tile_order[b/8], block b%8, one 4x2 quarter per warp) followed by 2*tiles
// COARSE candidates (tile_order[b/2], four blocks per CTA); a CTA whose tile belongs to the other shape exits.
template <int ILP>
__global__ void __launch_bounds__(kWarps * 32) blend_bwd_sparse_kernel(const BwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpBuf& buf = reinterpret_cast<WarpBuf*>(smem_raw)[warp];
  const int b = (int)blockIdx.x;
  if (b < 8 * a.fine_tiles) {
    const uint4 d = a.bdesc[b];  // (position b / 8, block b % 8)
    if ((int)d.y < a.heavy) return;
    blend_bwd_block<4, 1>(a, buf, d, b & 7, warp, lane);
  } else {
    const int c = b - 8 * a.fine_tiles;
    const int block = (c & 1) * 4 + warp;
    const uint4 d = a.bdesc[(c >> 1) * 8 + block];
    // FINE CTAs exist for the first fine_tiles entries of the order only: a heavy tile beyond them is done here
    if ((int)d.y >= a.heavy && (c >> 1) < a.fine_tiles) return;
    blend_bwd_block<1, ILP>(a, buf, d, block, 0, lane);
  }
}

}  // namespace

cudaError_t launch_blend_bwd_sparse(int width, int height, const ImageState& img, const BinningState& b,
                                    const float* background, const float* dL_dpix, const float* dL_dacc,
                                    float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                                    cudaStream_t s) {
  const int grid_x = (width + kTile - 1) / kTile, grid_y = (height + kTile - 1) / kTile;
  const int tiles = grid_x * grid_y;
  constexpr int kSmem = (int)sizeof(WarpBuf) * kWarps;
  const Tuning t = tuning();
  auto kernel = t.bwd_ilp == 1 ? blend_bwd_sparse_kernel<1> : blend_bwd_sparse_kernel<2>;  // pairs per lane per step
  if (cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), kSmem)) return e;
  BwdArgs a;
  a.ranges = img.ranges, a.tile_order = img.tile_order, a.inst = b.inst, a.point_list = b.vals;
  a.blist = b.blist, a.bdesc = img.bdesc;
  a.bhits = b.bhits, a.tail_hits = img.tail_hits, a.last_group = img.last_group;
  a.W = width, a.H = height, a.grid_x = grid_x;
  a.heavy = t.heavy_bwd;
  a.fine_tiles = tiles < kFineCap ? tiles : kFineCap;
  a.bg = background, a.final_T = img.final_T;
  a.dL_dpix = dL_dpix, a.dL_dacc = dL_dacc;
  a.dL_dmean2D = dL_dmean2D, a.dL_dconic = dL_dconic, a.dL_dopacity = dL_dopacity, a.dL_dcolor = dL_dcolor;
  kernel<<<8 * a.fine_tiles + 2 * tiles, kWarps * 32, kSmem, s>>>(a);
  return cudaGetLastError();
}

}  // namespace gsr
