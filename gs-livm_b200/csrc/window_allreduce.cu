// Gradient exchange of the view-sharded keyframe window: all-reduce (sum, FP32) of the flat
// per-Gaussian parameter-gradient buffer over NVLink 5 / NVSwitch, written against peer memory instead of calling
// a collective library.
//
// The reference has no multi-GPU path at all; this is the one cross-rank step of the design:
// every rank holds the gradients of ITS views in a symmetric buffer (same virtual layout on every GPU, mapped
// peer-to-peer and — where the fabric supports it — as one NVSwitch multicast object), and after the exchange
// every rank holds the same sum, bit for bit.
//
//  * multimem path (NVLS): rank r owns slice r of the buffer. One kernel pulls the slice already REDUCED by the
//    switch (multimem.ld_reduce.add.v4.f32 on the multicast address: the switch reads all replicas and adds) and
//    pushes the result back through the multicast address (multimem.st: the switch writes all replicas). Per GPU
//    that is (N-1)/N of the buffer out and in, both directions of the links busy at once, instead of twice that
//    for a ring or a two-shot peer-to-peer exchange.
//  * peer-to-peer path (no multicast): two kernels, reduce-scatter (rank r sums slice r over all peers' buffers
//    with plain peer loads) and all-gather (rank r copies the other reduced slices from their owners).
// Both need every rank's gradients complete before, and every rank's stores complete after: the caller brackets
// the kernels with the symmetric-memory barrier (gs_livm_b200/window.py), enqueued on the same stream.
#include <cstdlib>

#include "gsr_common.cuh"

namespace gsr {
namespace {

struct PeerPtrs {
  float4* p[16];
};

// .weak: plain (non-synchronising) accesses — ordering against the other ranks comes from the barriers around the
// kernel, not from these instructions.
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.weak.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* mc, const float4& v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// [lo, hi) of rank r's slice, in float4 units
__device__ __forceinline__ void slice_of(size_t n4, int rank, int world, size_t& lo, size_t& hi) {
  const size_t per = (n4 + world - 1) / world;
  lo = min(n4, per * rank);
  hi = min(n4, lo + per);
}

constexpr int kInFlight = 8;  // independent 16-byte requests per thread

__global__ void __launch_bounds__(512) allreduce_multimem_kernel(float4* mc, size_t n4, int rank, int world) {
  size_t lo, hi;
  slice_of(n4, rank, world, lo, hi);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (kInFlight - 1) * stride < hi; i += kInFlight * stride) {
    float4 v[kInFlight];
#pragma unroll
    for (int k = 0; k < kInFlight; ++k) v[k] = multimem_ld_reduce_add(mc + i + k * stride);
#pragma unroll
    for (int k = 0; k < kInFlight; ++k) multimem_st(mc + i + k * stride, v[k]);
  }
  for (; i < hi; i += stride) multimem_st(mc + i, multimem_ld_reduce_add(mc + i));
}

__global__ void __launch_bounds__(512) reduce_scatter_p2p_kernel(PeerPtrs peers, size_t n4, int rank, int world) {
  size_t lo, hi;
  slice_of(n4, rank, world, lo, hi);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // fixed rank order: every slice is summed the same way wherever it is summed
  for (; i + (kInFlight - 1) * stride < hi; i += kInFlight * stride) {
    float4 acc[kInFlight];
#pragma unroll
    for (int k = 0; k < kInFlight; ++k) acc[k] = peers.p[0][i + k * stride];
    for (int r = 1; r < world; ++r) {
      float4 v[kInFlight];
#pragma unroll
      for (int k = 0; k < kInFlight; ++k) v[k] = peers.p[r][i + k * stride];
#pragma unroll
      for (int k = 0; k < kInFlight; ++k) acc[k].x += v[k].x, acc[k].y += v[k].y, acc[k].z += v[k].z, acc[k].w += v[k].w;
    }
#pragma unroll
    for (int k = 0; k < kInFlight; ++k) peers.p[rank][i + k * stride] = acc[k];
  }
  for (; i < hi; i += stride) {
    float4 acc = peers.p[0][i];
    for (int r = 1; r < world; ++r) {
      const float4 v = peers.p[r][i];
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    peers.p[rank][i] = acc;
  }
}

__global__ void __launch_bounds__(512) all_gather_p2p_kernel(PeerPtrs peers, size_t n4, int rank, int world) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int o = 1; o < world; ++o) {
    const int src = (rank + o) % world;  // everybody starts on a different owner
    size_t lo, hi;
    slice_of(n4, src, world, lo, hi);
    size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kInFlight - 1) * stride < hi; i += kInFlight * stride) {
      float4 v[kInFlight];
#pragma unroll
      for (int k = 0; k < kInFlight; ++k) v[k] = peers.p[src][i + k * stride];
#pragma unroll
      for (int k = 0; k < kInFlight; ++k) peers.p[rank][i + k * stride] = v[k];
    }
    for (; i < hi; i += stride) peers.p[rank][i] = peers.p[src][i];
  }
}

int grid_for(int dev) {
  static int sms[64] = {};
  if (dev < 0 || dev >= 64) return 148 * 2;
  if (!sms[dev] && cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms[dev] = 148;
  static const int mult = [] {
    const char* e = getenv("GSR_AR_GRID");  // tuning knob: CTAs per SM of the exchange kernels
    return e ? atoi(e) : 2;
  }();
  return sms[dev] * (mult > 0 ? mult : 2);
}

}  // namespace
}  // namespace gsr

using namespace gsr;

extern "C" {

// phase 0 = the whole exchange through the multicast mapping; `buffer` is the MULTICAST pointer of the symmetric
// buffer. n_floats must be a multiple of 4 and the buffer 16-byte aligned.
int gsr_window_allreduce_multimem(void* multicast_buffer, size_t n_floats, int rank, int world, void* stream) {
  if (!multicast_buffer || (n_floats & 3) || world < 1 || world > 16 || rank < 0 || rank >= world) return GSR_ERR_INVALID;
  int dev = 0;
  cudaGetDevice(&dev);
  allreduce_multimem_kernel<<<grid_for(dev), 512, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<float4*>(multicast_buffer), n_floats / 4, rank, world);
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

// peer-to-peer exchange: phase 1 = reduce-scatter, phase 2 = all-gather (a symmetric-memory barrier between them).
// `peer_buffers` = host array of `world` device pointers, entry r = rank r's buffer as mapped into this process.
int gsr_window_allreduce_p2p(void* const* peer_buffers, size_t n_floats, int rank, int world, int phase, void* stream) {
  if (!peer_buffers || (n_floats & 3) || world < 1 || world > 16 || rank < 0 || rank >= world || (phase != 1 && phase != 2))
    return GSR_ERR_INVALID;
  PeerPtrs pp;
  for (int r = 0; r < 16; ++r) pp.p[r] = r < world ? static_cast<float4*>(peer_buffers[r]) : nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (phase == 1)
    reduce_scatter_p2p_kernel<<<grid_for(dev), 512, 0, s>>>(pp, n_floats / 4, rank, world);
  else
    all_gather_p2p_kernel<<<grid_for(dev), 512, 0, s>>>(pp, n_floats / 4, rank, world);
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

}  // extern "C"
