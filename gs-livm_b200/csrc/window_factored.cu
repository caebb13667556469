// Factored gradient exchange of the view-sharded keyframe window. sm_100a.
//
// 81 % of the 236 B per Gaussian a dense all-reduce moves are the dL_dsh rows, and per view those rows have rank one:
//     dL_dsh_v[k][c] = basis_k(dir_v) * dRGB_v[c]        (reference computeColorFromSH backward,
//                                                          src/cuda_rasterizer/backward.cu:56-135)
// with dir_v = normalize(mean - campos_v), known to every rank (parameters are replicated, the window's cameras are
// host data), and dRGB_v = dL_dcolor_v with the channels the forward clamped zeroed (backward.cu:41-43). So instead of
// summing 48 floats per Gaussian across the ranks, every rank
//   1. PUSHES its view's 3 masked dRGB floats per rendered Gaussian (+ one visibility bit) into slot v of every rank
//      (push_rgb_kernel: multimem.st through the NVSwitch multicast mapping, or plain peer stores) -- as soon as the
//      blend backward is done, i.e. UNDER geom_bwd, which no longer writes dL_dsh at all (GSR_BWD_SH_FACTORED);
//   2. all-reduces only the other 11 floats (mean3D, scale, rotation, opacity) and only for the Gaussians some view of
//      the step rendered (rest_*_kernel: groups of four Gaussians whose gathered visibility bits are all clear are
//      skipped; on the bench's window that is 37 % of them);
//   3. rebuilds  dL_dsh = sum_v basis(dir_v) (x) dRGB_v  locally (expand_sh_kernel), views in a fixed order, so every
//      rank ends with identical bits. The expansion is HBM-bound local work and runs beside the all-reduce.
// Wire traffic per Gaussian and rank: 44 B all-reduced (rendered by someone) + 12 B x (views rendered elsewhere that
// saw the Gaussian) gathered, instead of 236 B all-reduced.
//
// Cross-rank ordering is ours as well and lives INSIDE these kernels: every rank has a pad of flag words in the
// symmetric allocation, one row per channel, one word per source rank. A signal is a release store of the step's
// epoch into the peers' pads (one-way, no round trip), a wait spins on the own pad until every peer's word has
// reached the epoch. The epoch is a word in local device memory that the last kernel of the step advances, so a
// captured CUDA graph replays correctly. Channels: READY (my gradients are complete), PART (p2p: my slice is reduced),
// DONE (my part of the all-reduce is finished), PUSHED (my dRGB slot is written everywhere).
#include "gsr_common.cuh"
#include "sh_basis.cuh"

namespace gsr {
namespace {

constexpr int kMaxRanks = 16;
constexpr int kMaxViews = 16;
enum { CH_READY = 0, CH_PART = 1, CH_DONE = 2, CH_PUSHED = 3 };
// local words behind the flag rows of the own pad: the epoch and the last-block tickets of the signalling kernels
enum { L_EPOCH = 0, L_TICKET_PUSH = 1, L_TICKET_A = 2, L_TICKET_B = 3, L_TIMEOUT = 4 };
constexpr int kLocalAt = GSR_WINDOW_FLAG_CHANNELS * kMaxRanks;
// A wait gives up after this many SM cycles (~10 s at 1.9 GHz) instead of hanging the GPU when a peer has died or the
// ranks issue different call sequences: it records 1 + channel in the pad's L_TIMEOUT word (gsr_window_status) and lets
// the kernel run on -- the step's result is then garbage, but the process can report it and exit.
constexpr long long kWaitCycles = 20000000000ll;

struct Comm {
  int rank, world;            // world == 1: no cross-rank work at all
  uint32_t* pad[kMaxRanks];   // rank r's pad as mapped here
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// what a flag must have reached for chunk `sub` (< 32) of the current step
__device__ __forceinline__ uint32_t step_target(const Comm& c, int sub) {
  return (*reinterpret_cast<volatile uint32_t*>(c.pad[c.rank] + kLocalAt + L_EPOCH) + 1u) * 32u + (uint32_t)sub;
}
// called by threads 0 .. world-1 of ONE block, after that block has observed every write the signal covers
__device__ __forceinline__ void signal_peers(const Comm& c, int channel, uint32_t e) {
  const int p = threadIdx.x;
  if (p < c.world && p != c.rank) {
    __threadfence_system();
    st_release_sys(c.pad[p] + channel * kMaxRanks + c.rank, e);
  }
}
// called by every thread of a block; returns once all peers have signalled `e` on the channel
__device__ __forceinline__ void wait_peers(const Comm& c, int channel, uint32_t e) {
  const int p = threadIdx.x;
  if (p < c.world && p != c.rank) {
    const uint32_t* f = c.pad[c.rank] + channel * kMaxRanks + p;
    const long long t0 = clock64();
    for (uint32_t spins = 0; (int32_t)(ld_acquire_sys(f) - e) < 0; ++spins) {
      if ((spins & 1023u) == 1023u && clock64() - t0 > kWaitCycles) {
        atomicExch(c.pad[c.rank] + kLocalAt + L_TIMEOUT, 1u + (uint32_t)channel);
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
}
// true in exactly one block of the grid: the last one to get here (its threads then see every other block's writes)
__device__ __forceinline__ bool last_block(const Comm& c, int ticket_word) {
  __shared__ bool s_last;
  __threadfence();  // this block's stores are ordered before its ticket; the last block's system-scope fence + release
                    // store (signal_peers) then publishes all of them (fence cumulativity)
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* t = c.pad[c.rank] + kLocalAt + ticket_word;
    const uint32_t n = gridDim.x * gridDim.y;
    s_last = atomicAdd(t, 1u) == n - 1u;
    if (s_last) *t = 0u;  // replay-safe
  }
  __syncthreads();
  return s_last;
}

__global__ void __launch_bounds__(32) wait_kernel(Comm c, int channel, int sub, int advance_epoch) {
  wait_peers(c, channel, step_target(c, sub));
  if (advance_epoch && threadIdx.x == 0) *(c.pad[c.rank] + kLocalAt + L_EPOCH) += 1u;
}

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.weak.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st4(float4* mc, const float4& v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void multimem_st1(uint32_t* mc, uint32_t v) {
  asm volatile("multimem.st.weak.global.b32 [%0], %1;" ::"l"(mc), "r"(v) : "memory");
}

// ---- push ---------------------------------------------------------------------------------------------------------
struct PushArgs {
  int first, end, sub, n_dst, multicast;  // Gaussians [first, end), first a multiple of 128
  const float* dL_dcolor;      // [P,3] as accumulated by the blend backward
  const uint8_t* clamped;      // [P] bit c: channel c clamped by the forward
  const int* radii;            // [P]
  float4* rgb[kMaxRanks];      // slot of THIS rank's view in every rank's gather area ([0] = multicast mapping if multicast)
  uint32_t* vis[kMaxRanks];    // its visibility words
  Comm comm;
};

// One thread per four Gaussians = 12 floats = three 16-byte words of the slot; a warp covers 128 Gaussians = four
// visibility words. Words whose four Gaussians were all culled are not sent: nobody reads them (the readers look at the
// visibility bit first). The last block to finish signals PUSHED.
template <bool kAligned>
__global__ void __launch_bounds__(256) push_rgb_kernel(const PushArgs a) {
  const int t = a.first / 4 + blockIdx.x * blockDim.x + threadIdx.x;
  const int g0 = 4 * t;
  float v[12];
  uint32_t nib = 0;
  if (g0 + 3 < a.end) {
    int4 r;
    uint32_t cl4;
    if (kAligned) {
      r = __ldg(reinterpret_cast<const int4*>(a.radii) + t);
      cl4 = __ldg(reinterpret_cast<const uint32_t*>(a.clamped) + t);
      const float4* c4 = reinterpret_cast<const float4*>(a.dL_dcolor) + 3 * (size_t)t;
      const float4 x = __ldg(c4), y = __ldg(c4 + 1), z = __ldg(c4 + 2);
      v[0] = x.x, v[1] = x.y, v[2] = x.z, v[3] = x.w, v[4] = y.x, v[5] = y.y, v[6] = y.z, v[7] = y.w;
      v[8] = z.x, v[9] = z.y, v[10] = z.z, v[11] = z.w;
    } else {
      r = make_int4(a.radii[g0], a.radii[g0 + 1], a.radii[g0 + 2], a.radii[g0 + 3]);
      cl4 = a.clamped[g0] | (a.clamped[g0 + 1] << 8) | (a.clamped[g0 + 2] << 16) | ((uint32_t)a.clamped[g0 + 3] << 24);
#pragma unroll
      for (int i = 0; i < 12; ++i) v[i] = __ldg(a.dL_dcolor + 3 * (size_t)g0 + i);
    }
    nib = (r.x > 0 ? 1u : 0u) | (r.y > 0 ? 2u : 0u) | (r.z > 0 ? 4u : 0u) | (r.w > 0 ? 8u : 0u);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (!((nib >> j) & 1u) || ((cl4 >> (8 * j + c)) & 1u)) v[3 * j + c] = 0.f;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = g0 + j;
      const bool on = g < a.end && a.radii[g] > 0;
      const uint32_t cl = on ? a.clamped[g] : 0u;
      if (on) nib |= 1u << j;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[3 * j + c] = (on && !((cl >> c) & 1u)) ? __ldg(a.dL_dcolor + 3 * (size_t)g + c) : 0.f;
    }
  }
  // four visibility words per warp: lanes 8w .. 8w+7 hold the nibbles of word w
  const int lane = threadIdx.x & 31;
  uint32_t word = nib << (4 * (lane & 7));
  word |= __shfl_xor_sync(0xffffffffu, word, 1);
  word |= __shfl_xor_sync(0xffffffffu, word, 2);
  word |= __shfl_xor_sync(0xffffffffu, word, 4);
  const int n_words = (a.end + 31) / 32;
  const int w = t >> 3;
  const bool send_word = (lane & 7) == 0 && w < n_words;
  // the twelve floats go through shared memory so that the stores to the peers are coalesced 16-byte words (a warp
  // writes 512 contiguous bytes per instruction); word q of the block belongs to thread q / 3 alone
  __shared__ float4 s_q[3 * 256];
  __shared__ uint8_t s_on[256];
  s_q[3 * threadIdx.x] = make_float4(v[0], v[1], v[2], v[3]);
  s_q[3 * threadIdx.x + 1] = make_float4(v[4], v[5], v[6], v[7]);
  s_q[3 * threadIdx.x + 2] = make_float4(v[8], v[9], v[10], v[11]);
  s_on[threadIdx.x] = nib != 0;
  __syncthreads();
  const size_t block_q = 3 * ((size_t)a.first / 4 + (size_t)blockIdx.x * blockDim.x);
  if (a.multicast) {
    if (send_word) multimem_st1(a.vis[0] + w, word);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int q = threadIdx.x + 256 * k;
      if (s_on[q / 3]) multimem_st4(a.rgb[0] + block_q + q, s_q[q]);
    }
  } else {
    for (int r = 0; r < a.n_dst; ++r) {
      if (send_word) a.vis[r][w] = word;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int q = threadIdx.x + 256 * k;
        if (s_on[q / 3]) a.rgb[r][block_q + q] = s_q[q];
      }
    }
  }
  if (a.comm.world > 1 && last_block(a.comm, L_TICKET_PUSH)) signal_peers(a.comm, CH_PUSHED, step_target(a.comm, a.sub));
}

// ---- expansion ----------------------------------------------------------------------------------------------------
constexpr int kExpBlock = 128;
constexpr int kExpRow = 13;  // float4 per staged row: 12 used + 1 pad (conflict-free lane-strided STS.128)

struct ExpandArgs {
  int first, end, D, V;  // Gaussians [first, end)
  const float* means3D;
  const float* cam_pos;         // device [V,3]: camera centres of the step's views, in summation order
  const float* rgb[kMaxViews];  // gathered masked dRGB, [P,3] per view
  const uint32_t* vis[kMaxViews];
  float* dL_dsh;                // [P,16,3]
  uint32_t* row_state;          // optional [ceil(P/32)]: bit = the row may hold non-zero data (see expand_sh_kernel)
};

// dL_dsh[g] = sum over the views that rendered g of basis(dir) (x) dRGB, accumulated in view order with one FMA per
// term. The basis is the one geom_bwd uses (same sh_basis, same bits), so with one view the rows equal geom_bwd's bit
// for bit (fma(b, c, 0) is the rounded product); with more views they differ from the dense sum of the views' rounded
// rows by the usual last-place effects of a different summation. Rows nobody rendered are written as zeros -- or, with
// row_state, left alone when they are known to hold zeros already (a third of the bench's Gaussians is rendered by no
// view of a step: 71 MB of zero rows per step). The state word is requested at the top of the kernel and looked at only
// when the rows go out: as a dependent load in front of the stores the same idea was SLOWER (107 vs 57 us alone).
__global__ void __launch_bounds__(kExpBlock) expand_sh_kernel(const ExpandArgs a) {
  __shared__ __align__(16) float4 s_row[kExpBlock * kExpRow];
  __shared__ float s_cam[kMaxViews * 3];
  const int tid = threadIdx.x;
  if (tid < 3 * a.V) s_cam[tid] = __ldg(a.cam_pos + tid);
  __syncthreads();
  const int base = a.first + blockIdx.x * kExpBlock;
  const int n = min(kExpBlock, a.end - base);
  const int idx = base + tid;
  const uint32_t may_be_nonzero = (a.row_state && tid < n) ? a.row_state[idx >> 5] : 0xffffffffu;
  float acc[48];
#pragma unroll
  for (int i = 0; i < 48; ++i) acc[i] = 0.f;
  uint32_t seen = 0;
  if (tid < n) {
    for (int v = 0; v < a.V; ++v) seen |= ((__ldg(a.vis[v] + (idx >> 5)) >> (idx & 31)) & 1u) << v;
    if (seen) {
      const float mx = __ldg(a.means3D + 3 * (size_t)idx), my = __ldg(a.means3D + 3 * (size_t)idx + 1),
                  mz = __ldg(a.means3D + 3 * (size_t)idx + 2);
      for (int v0 = 0; v0 < a.V; v0 += 4) {  // four views' loads in flight at a time
        float c[4][3];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool on = v0 + k < a.V && ((seen >> (v0 + k)) & 1u);
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) c[k][ch] = on ? __ldg(a.rgb[min(v0 + k, a.V - 1)] + 3 * (size_t)idx + ch) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int v = v0 + k;
          if (v >= a.V || !((seen >> v) & 1u)) continue;
          float ox, oy, oz, x, y, z, b[16];
          sh_direction(mx, my, mz, s_cam[3 * v], s_cam[3 * v + 1], s_cam[3 * v + 2], ox, oy, oz, x, y, z);
          sh_basis_values(a.D, x, y, z, b);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            acc[3 * j] = __fmaf_rn(b[j], c[k][0], acc[3 * j]);
            acc[3 * j + 1] = __fmaf_rn(b[j], c[k][1], acc[3 * j + 1]);
            acc[3 * j + 2] = __fmaf_rn(b[j], c[k][2], acc[3 * j + 2]);
          }
        }
      }
    }
  }
  // rows that go out: rendered by some view, or still holding an earlier step's data
  const uint32_t now = __ballot_sync(0xffffffffu, seen != 0);
  const bool out = tid < n && (((now | may_be_nonzero) >> (tid & 31)) & 1u);
  if (a.row_state && tid < n && (tid & 31) == 0 && now != may_be_nonzero) a.row_state[idx >> 5] = now;
  __shared__ uint8_t s_out[kExpBlock];
  s_out[tid] = out;
  if (out) {
#pragma unroll
    for (int i = 0; i < 12; ++i)
      s_row[tid * kExpRow + i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
  }
  __syncthreads();
  float4* d4 = reinterpret_cast<float4*>(a.dL_dsh) + (size_t)base * 12;
  for (int q = tid, row = tid / 12, rem = tid % 12; q < n * 12; q += kExpBlock) {
    if (s_out[row]) d4[q] = s_row[row * kExpRow + rem];
    { rem += kExpBlock % 12, row += kExpBlock / 12; if (rem >= 12) rem -= 12, ++row; }
  }
}

// ---- all-reduce of the other 11 floats ----------------------------------------------------------------------------
// The four arrays (16-byte aligned, in one symmetric buffer) as float4 words; group j = Gaussians 4j .. 4j+3 owns words
// 3j..3j+2 of mean3D and of scale, 4j..4j+3 of rotation and word j of opacity: 11 words.
struct RestLayout {
  size_t mean4, scale4, rot4, opac4;  // first float4 word of each array in the buffer
  uint32_t n_mean4, n_rot4, n_opac4;  // words per array (mean3D and scale alike)
  uint32_t g_first, g_end;            // the chunk's groups of four Gaussians
  int sub, pushed_sub;                // READY / DONE use sub; the visibility words are complete at PUSHED(pushed_sub)
  int V;
  const uint32_t* vis[kMaxViews];     // gathered visibility words of the step's views (V == 0: dense)
};
__device__ __forceinline__ bool group_rendered(const RestLayout& L, uint32_t j) {
  if (L.V == 0) return true;
  uint32_t u = 0;
  for (int v = 0; v < L.V; ++v) u |= __ldg(L.vis[v] + (j >> 3));
  return ((u >> (4 * (j & 7))) & 0xfu) != 0;
}
// entry of the reducing kernels: every rank's gradients of the chunk are complete (and, for the sparse exchange, every
// rank's visibility words of the chunk have landed here)
__device__ __forceinline__ uint32_t rest_entry(const RestLayout& L, const Comm& c) {
  const uint32_t e = step_target(c, L.sub);
  if (blockIdx.x == 0) signal_peers(c, CH_READY, e);
  wait_peers(c, CH_READY, e);
  if (L.V > 0) wait_peers(c, CH_PUSHED, step_target(c, L.pushed_sub));
  return e;
}

// Ownership: the groups are dealt to the ranks in blocks of 256 (1024 Gaussians), block bk -> rank bk % world. (A
// contiguous slice per rank leaves the ranks with very different amounts of work: the Gaussians some view rendered are
// clustered in index space, and the exchange takes as long as its busiest rank.)
constexpr uint32_t kOwnShift = 8;
// Every float4 word of `owner`'s groups that takes part, array by array so that the lanes of a warp touch consecutive
// words (512 contiguous bytes per warp instruction: what peer and multicast accesses need), four words per thread in
// flight: v = load(word) for all four, then store(word, v).
constexpr int kRestInFlight = 4;
template <typename Load, typename Store>
__device__ __forceinline__ void rest_for_each_word(const RestLayout& L, int owner, int world, Load load, Store store) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t bk_lo = L.g_first >> kOwnShift, bk_hi = (L.g_end + (1u << kOwnShift) - 1u) >> kOwnShift;
  const uint32_t first_bk = bk_lo + (uint32_t)((owner + world - (int)(bk_lo % (uint32_t)world)) % world);
  if (first_bk >= bk_hi) return;
  const uint32_t n_blocks = (bk_hi - first_bk + (uint32_t)world - 1u) / (uint32_t)world;
#pragma unroll 1
  for (int arr = 0; arr < 4; ++arr) {
    const uint32_t per = arr < 2 ? 3u : (arr == 2 ? 4u : 1u);  // words per group
    const size_t first = arr == 0 ? L.mean4 : (arr == 1 ? L.scale4 : (arr == 2 ? L.rot4 : L.opac4));
    const size_t n_arr = arr < 2 ? L.n_mean4 : (arr == 2 ? L.n_rot4 : L.n_opac4);
    const size_t block_words = (size_t)per << kOwnShift, total = (size_t)n_blocks * block_words;
    for (size_t u = t0; u < total; u += kRestInFlight * stride) {
      float4 v[kRestInFlight];
      size_t w[kRestInFlight];
      bool on[kRestInFlight];
#pragma unroll
      for (int k = 0; k < kRestInFlight; ++k) {
        const size_t uk = u + k * stride;
        const size_t i = (uk >> kOwnShift) / per;  // which of the owner's blocks
        w[k] = ((size_t)first_bk + i * (size_t)world) * block_words + (uk - i * block_words);
        const uint32_t g = (uint32_t)(w[k] / per);
        on[k] = uk < total && g >= L.g_first && g < L.g_end && w[k] < n_arr && group_rendered(L, g);
      }
#pragma unroll
      for (int k = 0; k < kRestInFlight; ++k)
        if (on[k]) v[k] = load(first + w[k]);
#pragma unroll
      for (int k = 0; k < kRestInFlight; ++k)
        if (on[k]) store(first + w[k], v[k]);
    }
  }
}

// NVSwitch path: rank r pulls the words of ITS groups already reduced by the switch and multicasts them back.
// Entry: READY signalled / awaited (every rank's gradients are complete); exit: DONE signalled by the last block.
__global__ void __launch_bounds__(512) rest_multimem_kernel(float4* mc, RestLayout L, Comm c) {
  const uint32_t e = rest_entry(L, c);
  rest_for_each_word(
      L, c.rank, c.world, [&](size_t w) { return multimem_ld_reduce_add(mc + w); },
      [&](size_t w, const float4& v) { multimem_st4(mc + w, v); });
  if (last_block(c, L_TICKET_A)) signal_peers(c, CH_DONE, e);
}

struct PeerBuf {
  float4* p[kMaxRanks];
};
// peer-to-peer path, phase 1: rank r sums its groups over all ranks' buffers (fixed rank order) into its own buffer
__global__ void __launch_bounds__(512) rest_reduce_scatter_kernel(PeerBuf peers, RestLayout L, Comm c) {
  const uint32_t e = rest_entry(L, c);
  float4* own = peers.p[c.rank];
  rest_for_each_word(
      L, c.rank, c.world,
      [&](size_t w) {
        float4 acc = peers.p[0][w];
        for (int r = 1; r < c.world; ++r) {
          const float4 v = peers.p[r][w];
          acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        }
        return acc;
      },
      [&](size_t w, const float4& v) { own[w] = v; });
  if (last_block(c, L_TICKET_A)) signal_peers(c, CH_PART, e);
}
// phase 2: every rank copies the other ranks' groups from their owners
__global__ void __launch_bounds__(512) rest_all_gather_kernel(PeerBuf peers, RestLayout L, Comm c) {
  const uint32_t e = step_target(c, L.sub);
  wait_peers(c, CH_PART, e);
  float4* own = peers.p[c.rank];
  for (int o = 1; o < c.world; ++o) {
    const int src = (c.rank + o) % c.world;  // everybody starts on a different owner
    const float4* from = peers.p[src];
    rest_for_each_word(
        L, src, c.world, [&](size_t w) { return from[w]; }, [&](size_t w, const float4& v) { own[w] = v; });
  }
  if (last_block(c, L_TICKET_B)) signal_peers(c, CH_DONE, e);
}

bool make_comm(const gsr_window_comm* in, Comm& c) {
  c.rank = 0, c.world = 1;
  for (int r = 0; r < kMaxRanks; ++r) c.pad[r] = nullptr;
  if (!in) return true;
  if (in->world < 1 || in->world > kMaxRanks || in->rank < 0 || in->rank >= in->world) return false;
  c.rank = in->rank, c.world = in->world;
  for (int r = 0; r < c.world; ++r) {
    c.pad[r] = static_cast<uint32_t*>(in->pads[r]);
    if (c.world > 1 && !c.pad[r]) return false;
  }
  return true;
}

}  // namespace
}  // namespace gsr

using namespace gsr;

extern "C" {

size_t gsr_window_slot_floats(int P) { return P > 0 ? 12 * (((size_t)P + 3) / 4) : 0; }
size_t gsr_window_pad_words(void) { return kLocalAt + 8; }

int gsr_window_status(const gsr_window_comm* comm, uint32_t* timed_out_channel) {
  Comm c;
  if (!make_comm(comm, c) || !timed_out_channel) return GSR_ERR_INVALID;
  *timed_out_channel = 0;
  if (c.world == 1) return GSR_OK;
  if (cudaMemcpy(timed_out_channel, c.pad[c.rank] + kLocalAt + L_TIMEOUT, sizeof(uint32_t), cudaMemcpyDeviceToHost) !=
      cudaSuccess)
    return GSR_ERR_CUDA;
  return GSR_OK;
}

static bool range_ok(int P, int first, int count) {
  return P > 0 && first >= 0 && count > 0 && (first % 128) == 0 && (long long)first + count <= P;
}

int gsr_window_wait(const gsr_window_comm* comm, int channel, int sub, int advance_epoch, void* stream) {
  Comm c;
  if (!make_comm(comm, c) || channel < 0 || channel >= GSR_WINDOW_FLAG_CHANNELS || sub < 0 || sub >= 32) return GSR_ERR_INVALID;
  if (c.world == 1) return GSR_OK;
  wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(c, channel, sub, advance_epoch);
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

int gsr_window_push_rgb(char* geom_buffer, int P, int first, int count, int sub, const int* radii, const float* dL_dcolor,
                        void* const* dst_rgb, void* const* dst_vis, int n_dst, int multicast, const gsr_window_comm* comm,
                        void* stream) {
  if (!geom_buffer || !range_ok(P, first, count) || sub < 0 || sub >= 32 || !dL_dcolor || !dst_rgb || !dst_vis || n_dst < 1 ||
      n_dst > kMaxRanks)
    return GSR_ERR_INVALID;
  const GeomState geom = carve_geom(geom_buffer, P);
  PushArgs a;
  if (!make_comm(comm, a.comm)) return GSR_ERR_INVALID;
  a.first = first, a.end = first + count, a.sub = sub, a.multicast = multicast ? 1 : 0, a.n_dst = multicast ? 1 : n_dst;
  a.dL_dcolor = dL_dcolor, a.clamped = geom.clamped, a.radii = radii ? radii : geom.internal_radii;
  for (int r = 0; r < kMaxRanks; ++r) {
    a.rgb[r] = r < a.n_dst ? static_cast<float4*>(dst_rgb[r]) : nullptr;
    a.vis[r] = r < a.n_dst ? static_cast<uint32_t*>(dst_vis[r]) : nullptr;
    if (r < a.n_dst && (!a.rgb[r] || !a.vis[r] || (reinterpret_cast<uintptr_t>(a.rgb[r]) & 15u))) return GSR_ERR_INVALID;
  }
  const int threads = (count + 3) / 4;
  const int blocks = (threads + 255) / 256;  // whole warps, so that the nibble exchange sees every lane
  const bool aligned = !(reinterpret_cast<uintptr_t>(a.dL_dcolor) & 15u) && !(reinterpret_cast<uintptr_t>(a.radii) & 15u) &&
                       !(reinterpret_cast<uintptr_t>(a.clamped) & 3u);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (aligned) push_rgb_kernel<true><<<blocks, 256, 0, s>>>(a);
  else push_rgb_kernel<false><<<blocks, 256, 0, s>>>(a);
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

int gsr_window_expand_sh(int P, int first, int count, int D, int M, int n_views, const float* means3D, const float* cam_pos,
                         const void* const* rgb_slots, const void* const* vis_slots, float* dL_dsh, uint32_t* row_state,
                         void* stream) {
  if (!range_ok(P, first, count) || D < 0 || D > 3 || M != 16 || n_views < 1 || n_views > kMaxViews || !means3D || !cam_pos ||
      !rgb_slots || !vis_slots || !dL_dsh || (reinterpret_cast<uintptr_t>(dL_dsh) & 15u))
    return GSR_ERR_INVALID;
  ExpandArgs a;
  a.first = first, a.end = first + count, a.D = D, a.V = n_views, a.means3D = means3D, a.cam_pos = cam_pos, a.dL_dsh = dL_dsh;
  a.row_state = row_state;
  for (int v = 0; v < kMaxViews; ++v) {
    const bool on = v < n_views;
    a.rgb[v] = on ? static_cast<const float*>(rgb_slots[v]) : nullptr;
    a.vis[v] = on ? static_cast<const uint32_t*>(vis_slots[v]) : nullptr;
    if (on && (!a.rgb[v] || !a.vis[v])) return GSR_ERR_INVALID;
  }
  expand_sh_kernel<<<(count + kExpBlock - 1) / kExpBlock, kExpBlock, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

int gsr_window_allreduce_rest(const gsr_window_comm* comm, void* const* buffers, int multicast, int P, int first, int count,
                              int sub, int pushed_sub, const size_t* first_float, int n_views, const void* const* vis_slots,
                              void* stream) {
  Comm c;
  if (!make_comm(comm, c) || !comm || !buffers || !range_ok(P, first, count) || sub < 0 || sub >= 32 || pushed_sub < 0 ||
      pushed_sub >= 32 || !first_float || n_views < 0 || n_views > kMaxViews)
    return GSR_ERR_INVALID;
  if (c.world == 1) return GSR_OK;
  RestLayout L;
  for (int k = 0; k < 4; ++k)
    if (first_float[k] & 3u) return GSR_ERR_INVALID;
  L.mean4 = first_float[0] / 4, L.scale4 = first_float[1] / 4, L.rot4 = first_float[2] / 4, L.opac4 = first_float[3] / 4;
  L.n_mean4 = (uint32_t)((3 * (size_t)P + 3) / 4), L.n_rot4 = (uint32_t)P, L.n_opac4 = (uint32_t)(((size_t)P + 3) / 4);
  L.g_first = (uint32_t)(first / 4), L.g_end = (uint32_t)(((size_t)first + count + 3) / 4), L.sub = sub;
  L.pushed_sub = pushed_sub;
  L.V = vis_slots ? n_views : 0;
  for (int v = 0; v < kMaxViews; ++v) {
    L.vis[v] = (vis_slots && v < n_views) ? static_cast<const uint32_t*>(vis_slots[v]) : nullptr;
    if (vis_slots && v < n_views && !L.vis[v]) return GSR_ERR_INVALID;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = device_sm_count();  // one 512-thread CTA per SM: the links bound this, and the SMs stay open for the
                                       // kernels running beside it
  if (multicast) {
    if (!buffers[0] || (reinterpret_cast<uintptr_t>(buffers[0]) & 15u)) return GSR_ERR_INVALID;
    rest_multimem_kernel<<<grid, 512, 0, s>>>(static_cast<float4*>(buffers[0]), L, c);
  } else {
    PeerBuf pb;
    for (int r = 0; r < kMaxRanks; ++r) {
      pb.p[r] = r < c.world ? static_cast<float4*>(buffers[r]) : nullptr;
      if (r < c.world && (!pb.p[r] || (reinterpret_cast<uintptr_t>(pb.p[r]) & 15u))) return GSR_ERR_INVALID;
    }
    rest_reduce_scatter_kernel<<<grid, 512, 0, s>>>(pb, L, c);
    rest_all_gather_kernel<<<grid, 512, 0, s>>>(pb, L, c);
  }
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

}  // extern "C"
