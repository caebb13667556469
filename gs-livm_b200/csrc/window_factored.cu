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
//   2. all-reduces only the remaining 11 floats (mean3D, scale, rotation, opacity; window_allreduce.cu);
//   3. rebuilds  dL_dsh = sum_v basis(dir_v) (x) dRGB_v  locally (expand_sh_kernel), views in a fixed order, so every
//      rank ends with identical bits. The expansion is HBM-bound local work and runs beside the all-reduce.
// Wire traffic per Gaussian and rank: 44 B all-reduced + 12 B x (views rendered elsewhere that saw the Gaussian)
// gathered, instead of 236 B all-reduced.
//
// The cross-rank barrier is ours as well (barrier_kernel): flag words in the symmetric allocation, put/wait with
// system-scope CAS, replay-safe (every flag returns to 0), so the whole step -- render, push, barriers, exchange,
// expansion -- is captured in ONE CUDA graph.
#include "gsr_common.cuh"
#include "sh_basis.cuh"

namespace gsr {
namespace {

struct PeerU32 {
  uint32_t* p[16];
};
struct PeerF4 {
  float4* p[16];
};

// ---- barrier ------------------------------------------------------------------------------------------------------
// signals[r] = rank r's pad, [kBarrierChannels][16] words, all zero between barriers. Thread p of rank `rank`:
//   put : flip word [channel][rank] of peer p's pad 0 -> 1 (waits for 0: the peer consumed the previous round)
//   wait: flip word [channel][p] of the own pad 1 -> 0
// release/acquire at system scope around them orders every earlier write of this rank (previous kernels of the stream
// included) before every later read of the peers.
__global__ void __launch_bounds__(32) barrier_kernel(PeerU32 signals, int rank, int world, int channel) {
  const int p = threadIdx.x;
  if (p >= world || p == rank) return;
  __threadfence_system();
  uint32_t* put = signals.p[p] + channel * 16 + rank;
  while (atomicCAS_system(put, 0u, 1u) != 0u) {
  }
  uint32_t* wait = signals.p[rank] + channel * 16 + p;
  while (atomicCAS_system(wait, 1u, 0u) != 1u) {
  }
  __threadfence_system();
}

// ---- push ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void multimem_st4(float4* mc, const float4& v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void multimem_st1(uint32_t* mc, uint32_t v) {
  asm volatile("multimem.st.weak.global.b32 [%0], %1;" ::"l"(mc), "r"(v) : "memory");
}

struct PushArgs {
  int P, world, multicast;
  const float* dL_dcolor;  // [P,3] as accumulated by the blend backward
  const uint8_t* clamped;  // [P] bit c: channel c clamped by the forward
  const int* radii;        // [P]
  PeerF4 rgb;              // slot of THIS rank's view in every rank's gather area ([0] = multicast mapping if multicast)
  PeerU32 vis;             // its visibility words
};

// One thread per four Gaussians = 12 floats = three 16-byte words of the slot; a warp covers 128 Gaussians = four
// visibility words. Words whose four Gaussians were all culled are not sent: nobody reads them (expand_sh_kernel looks
// at the visibility bit first).
template <bool kAligned>
__global__ void __launch_bounds__(256) push_rgb_kernel(const PushArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int g0 = 4 * t;
  float v[12];
  uint32_t nib = 0;
  if (g0 + 3 < a.P) {
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
      const bool on = g < a.P && a.radii[g] > 0;
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
  const int n_words = (a.P + 31) / 32;
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
  const size_t block_q = 3 * (size_t)blockIdx.x * blockDim.x;
  if (a.multicast) {
    if (send_word) multimem_st1(a.vis.p[0] + w, word);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int q = threadIdx.x + 256 * k;
      if (s_on[q / 3]) multimem_st4(a.rgb.p[0] + block_q + q, s_q[q]);
    }
  } else {
    for (int r = 0; r < a.world; ++r) {
      if (send_word) a.vis.p[r][w] = word;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int q = threadIdx.x + 256 * k;
        if (s_on[q / 3]) a.rgb.p[r][block_q + q] = s_q[q];
      }
    }
  }
}

// ---- expansion ----------------------------------------------------------------------------------------------------
constexpr int kExpBlock = 128;
constexpr int kExpRow = 13;  // float4 per staged row: 12 used + 1 pad (conflict-free lane-strided STS.128)
constexpr int kMaxViews = 16;

struct ExpandArgs {
  int P, D, V;
  const float* means3D;
  const float* cam_pos;         // device [V,3]: camera centres of the step's views, in summation order
  const float* rgb[kMaxViews];  // gathered masked dRGB, [P,3] per view
  const uint32_t* vis[kMaxViews];
  float* dL_dsh;                // [P,16,3]
};

// dL_dsh[g] = sum over the views that rendered g of basis(dir) (x) dRGB: the products are the ones geom_bwd forms
// (geom_bwd.cu, same sh_basis), rounded, then added in view order -- so with one view the rows equal geom_bwd's bit
// for bit and with two they equal the dense all-reduce's. Rows nobody rendered are written as zeros.
__global__ void __launch_bounds__(kExpBlock) expand_sh_kernel(const ExpandArgs a) {
  __shared__ __align__(16) float4 s_row[kExpBlock * kExpRow];
  __shared__ float s_cam[kMaxViews * 3];
  const int tid = threadIdx.x;
  if (tid < 3 * a.V) s_cam[tid] = __ldg(a.cam_pos + tid);
  __syncthreads();
  const int base = blockIdx.x * kExpBlock;
  const int n = min(kExpBlock, a.P - base);
  const int idx = base + tid;
  float acc[48];
#pragma unroll
  for (int i = 0; i < 48; ++i) acc[i] = 0.f;
  if (tid < n) {
    uint32_t seen = 0;
    for (int v = 0; v < a.V; ++v) seen |= ((__ldg(a.vis[v] + (idx >> 5)) >> (idx & 31)) & 1u) << v;
    if (seen) {
      const float mx = __ldg(a.means3D + 3 * (size_t)idx), my = __ldg(a.means3D + 3 * (size_t)idx + 1),
                  mz = __ldg(a.means3D + 3 * (size_t)idx + 2);
      for (int v = 0; v < a.V; ++v) {
        if (!((seen >> v) & 1u)) continue;
        const float* c = a.rgb[v] + 3 * (size_t)idx;
        const float dr = __ldg(c), dg = __ldg(c + 1), db = __ldg(c + 2);
        float ox, oy, oz, x, y, z, b[16];
        sh_direction(mx, my, mz, s_cam[3 * v], s_cam[3 * v + 1], s_cam[3 * v + 2], ox, oy, oz, x, y, z);
        sh_basis_values(a.D, x, y, z, b);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          acc[3 * k] = __fadd_rn(acc[3 * k], __fmul_rn(b[k], dr));
          acc[3 * k + 1] = __fadd_rn(acc[3 * k + 1], __fmul_rn(b[k], dg));
          acc[3 * k + 2] = __fadd_rn(acc[3 * k + 2], __fmul_rn(b[k], db));
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 12; ++i)
    s_row[tid * kExpRow + i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
  __syncthreads();
  float4* d4 = reinterpret_cast<float4*>(a.dL_dsh) + (size_t)base * 12;
  for (int q = tid; q < n * 12; q += kExpBlock) d4[q] = s_row[(q / 12) * kExpRow + (q % 12)];
}

}  // namespace
}  // namespace gsr

using namespace gsr;

extern "C" {

size_t gsr_window_slot_floats(int P) { return P > 0 ? 12 * (((size_t)P + 3) / 4) : 0; }

int gsr_window_barrier(void* const* peer_signals, int rank, int world, int channel, void* stream) {
  if (!peer_signals || world < 1 || world > 16 || rank < 0 || rank >= world || channel < 0 ||
      channel >= GSR_WINDOW_BARRIER_CHANNELS)
    return GSR_ERR_INVALID;
  if (world == 1) return GSR_OK;
  PeerU32 s;
  for (int r = 0; r < 16; ++r) s.p[r] = r < world ? static_cast<uint32_t*>(peer_signals[r]) : nullptr;
  barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(s, rank, world, channel);
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

int gsr_window_push_rgb(char* geom_buffer, int P, const int* radii, const float* dL_dcolor, void* const* dst_rgb,
                        void* const* dst_vis, int world, int multicast, void* stream) {
  if (!geom_buffer || P <= 0 || !dL_dcolor || !dst_rgb || !dst_vis || world < 1 || world > 16) return GSR_ERR_INVALID;
  const GeomState geom = carve_geom(geom_buffer, P);
  PushArgs a;
  a.P = P, a.world = world, a.multicast = multicast ? 1 : 0;
  a.dL_dcolor = dL_dcolor, a.clamped = geom.clamped, a.radii = radii ? radii : geom.internal_radii;
  const int n_dst = multicast ? 1 : world;
  for (int r = 0; r < 16; ++r) {
    a.rgb.p[r] = r < n_dst ? static_cast<float4*>(dst_rgb[r]) : nullptr;
    a.vis.p[r] = r < n_dst ? static_cast<uint32_t*>(dst_vis[r]) : nullptr;
    if (r < n_dst && (!a.rgb.p[r] || !a.vis.p[r] || (reinterpret_cast<uintptr_t>(a.rgb.p[r]) & 15u))) return GSR_ERR_INVALID;
  }
  const int threads = (P + 3) / 4;
  // whole warps, so that the nibble exchange sees every lane
  const int blocks = (threads + 255) / 256;
  const bool aligned = !(reinterpret_cast<uintptr_t>(a.dL_dcolor) & 15u) && !(reinterpret_cast<uintptr_t>(a.radii) & 15u) &&
                       !(reinterpret_cast<uintptr_t>(a.clamped) & 3u);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (aligned) push_rgb_kernel<true><<<blocks, 256, 0, s>>>(a);
  else push_rgb_kernel<false><<<blocks, 256, 0, s>>>(a);
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

int gsr_window_expand_sh(int P, int D, int M, int n_views, const float* means3D, const float* cam_pos,
                         const void* const* rgb_slots, const void* const* vis_slots, float* dL_dsh, void* stream) {
  if (P <= 0 || D < 0 || D > 3 || M != 16 || n_views < 1 || n_views > kMaxViews || !means3D || !cam_pos ||
      !rgb_slots || !vis_slots || !dL_dsh || (reinterpret_cast<uintptr_t>(dL_dsh) & 15u))
    return GSR_ERR_INVALID;
  ExpandArgs a;
  a.P = P, a.D = D, a.V = n_views, a.means3D = means3D, a.cam_pos = cam_pos, a.dL_dsh = dL_dsh;
  for (int v = 0; v < kMaxViews; ++v) {
    const bool on = v < n_views;
    a.rgb[v] = on ? static_cast<const float*>(rgb_slots[v]) : nullptr;
    a.vis[v] = on ? static_cast<const uint32_t*>(vis_slots[v]) : nullptr;
    if (on && (!a.rgb[v] || !a.vis[v])) return GSR_ERR_INVALID;
  }
  expand_sh_kernel<<<(P + kExpBlock - 1) / kExpBlock, kExpBlock, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

}  // extern "C"
