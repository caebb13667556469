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

// Optional row mask: float4 words [lo4, hi4) of the buffer are rows of row4 words (the 192-byte dL_dsh rows: 81 % of
// the buffer) and row r takes part in the exchange only if bit r of `bits` is set -- the union over the ranks of
// "this Gaussian was visible in my view". A Gaussian visible in no view of the window has an all-zero row on every
// rank (geom_bwd writes the zeros), so skipping it changes nothing and saves its share of the link traffic.
struct RowMask {
  const uint32_t* bits;  // nullptr: dense exchange
  size_t lo4, hi4;
  uint32_t row4;
};
__device__ __forceinline__ bool takes_part(const RowMask& m, size_t i) {
  if (!m.bits || i < m.lo4 || i >= m.hi4) return true;
  const uint32_t r = (uint32_t)((i - m.lo4) / m.row4);
  return (m.bits[r >> 5] >> (r & 31)) & 1u;
}

__global__ void __launch_bounds__(512) allreduce_multimem_kernel(float4* mc, size_t n4, int rank, int world, RowMask m) {
  size_t lo, hi;
  slice_of(n4, rank, world, lo, hi);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (kInFlight - 1) * stride < hi; i += kInFlight * stride) {
    float4 v[kInFlight];
    bool on[kInFlight];
#pragma unroll
    for (int k = 0; k < kInFlight; ++k) on[k] = takes_part(m, i + k * stride);
#pragma unroll
    for (int k = 0; k < kInFlight; ++k)
      if (on[k]) v[k] = multimem_ld_reduce_add(mc + i + k * stride);
#pragma unroll
    for (int k = 0; k < kInFlight; ++k)
      if (on[k]) multimem_st(mc + i + k * stride, v[k]);
  }
  for (; i < hi; i += stride)
    if (takes_part(m, i)) multimem_st(mc + i, multimem_ld_reduce_add(mc + i));
}

__global__ void __launch_bounds__(512) reduce_scatter_p2p_kernel(PeerPtrs peers, size_t n4, int rank, int world, RowMask m) {
  size_t lo, hi;
  slice_of(n4, rank, world, lo, hi);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // fixed rank order: every slice is summed the same way wherever it is summed
  for (; i + (kInFlight - 1) * stride < hi; i += kInFlight * stride) {
    float4 acc[kInFlight];
    bool on[kInFlight];
#pragma unroll
    for (int k = 0; k < kInFlight; ++k) on[k] = takes_part(m, i + k * stride);
#pragma unroll
    for (int k = 0; k < kInFlight; ++k)
      if (on[k]) acc[k] = peers.p[0][i + k * stride];
    for (int r = 1; r < world; ++r) {
      float4 v[kInFlight];
#pragma unroll
      for (int k = 0; k < kInFlight; ++k)
        if (on[k]) v[k] = peers.p[r][i + k * stride];
#pragma unroll
      for (int k = 0; k < kInFlight; ++k)
        if (on[k]) acc[k].x += v[k].x, acc[k].y += v[k].y, acc[k].z += v[k].z, acc[k].w += v[k].w;
    }
#pragma unroll
    for (int k = 0; k < kInFlight; ++k)
      if (on[k]) peers.p[rank][i + k * stride] = acc[k];
  }
  for (; i < hi; i += stride) {
    if (!takes_part(m, i)) continue;
    float4 acc = peers.p[0][i];
    for (int r = 1; r < world; ++r) {
      const float4 v = peers.p[r][i];
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    peers.p[rank][i] = acc;
  }
}

__global__ void __launch_bounds__(512) all_gather_p2p_kernel(PeerPtrs peers, size_t n4, int rank, int world, RowMask m) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int o = 1; o < world; ++o) {
    const int src = (rank + o) % world;  // everybody starts on a different owner
    size_t lo, hi;
    slice_of(n4, src, world, lo, hi);
    size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kInFlight - 1) * stride < hi; i += kInFlight * stride) {
      float4 v[kInFlight];
      bool on[kInFlight];
#pragma unroll
      for (int k = 0; k < kInFlight; ++k) on[k] = takes_part(m, i + k * stride);
#pragma unroll
      for (int k = 0; k < kInFlight; ++k)
        if (on[k]) v[k] = peers.p[src][i + k * stride];
#pragma unroll
      for (int k = 0; k < kInFlight; ++k)
        if (on[k]) peers.p[rank][i + k * stride] = v[k];
    }
    for (; i < hi; i += stride)
      if (takes_part(m, i)) peers.p[rank][i] = peers.p[src][i];
  }
}

struct PeerWords {
  const uint32_t* p[16];
};

// word w, bit b: Gaussian 32 w + b was rendered by this rank's view (radius > 0)
__global__ void __launch_bounds__(256) visibility_words_kernel(const int* __restrict__ radii, int P, uint32_t* words) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t b = __ballot_sync(0xffffffffu, idx < P && radii[idx] > 0);
  if ((threadIdx.x & 31) == 0 && idx < P) words[idx >> 5] = b;
}

__global__ void __launch_bounds__(256) union_words_kernel(PeerWords peers, int n_words, int world, uint32_t* out) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  uint32_t u = 0;
  for (int r = 0; r < world; ++r) u |= peers.p[r][w];
  out[w] = u;
}

int grid_for(int /*dev*/) { return device_sm_count() * tuning().ar_grid; }  // CTAs per SM: Tuning::ar_grid

}  // namespace
}  // namespace gsr

using namespace gsr;

extern "C" {

// phase 0 = the whole exchange through the multicast mapping; `buffer` is the MULTICAST pointer of the symmetric
// buffer. n_floats must be a multiple of 4 and the buffer 16-byte aligned.
static bool make_mask(const gsr_window_rows* rows, size_t n_floats, RowMask& m) {
  m.bits = nullptr, m.lo4 = m.hi4 = 0, m.row4 = 1;
  if (!rows || !rows->union_words) return true;  // dense
  if ((rows->first_float & 3) || (rows->floats_per_row & 3) || rows->floats_per_row == 0 ||
      rows->first_float + (size_t)rows->n_rows * rows->floats_per_row > n_floats)
    return false;
  m.bits = rows->union_words, m.lo4 = rows->first_float / 4, m.row4 = rows->floats_per_row / 4;
  m.hi4 = m.lo4 + (size_t)rows->n_rows * m.row4;
  return true;
}

int gsr_window_allreduce_multimem(void* multicast_buffer, size_t n_floats, int rank, int world, void* stream) {
  return gsr_window_allreduce_multimem_rows(multicast_buffer, n_floats, rank, world, nullptr, stream);
}

int gsr_window_allreduce_multimem_rows(void* multicast_buffer, size_t n_floats, int rank, int world,
                                       const gsr_window_rows* rows, void* stream) {
  if (!multicast_buffer || (n_floats & 3) || world < 1 || world > 16 || rank < 0 || rank >= world) return GSR_ERR_INVALID;
  RowMask m;
  if (!make_mask(rows, n_floats, m)) return GSR_ERR_INVALID;
  int dev = 0;
  cudaGetDevice(&dev);
  allreduce_multimem_kernel<<<grid_for(dev), 512, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<float4*>(multicast_buffer), n_floats / 4, rank, world, m);
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

int gsr_window_visibility_words(const int* radii, int P, uint32_t* words, void* stream) {
  if (!radii || !words || P <= 0) return GSR_ERR_INVALID;
  visibility_words_kernel<<<(P + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(radii, P, words);
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

int gsr_window_union_words(void* const* peer_words, int n_words, int world, uint32_t* out, void* stream) {
  if (!peer_words || !out || n_words <= 0 || world < 1 || world > 16) return GSR_ERR_INVALID;
  PeerWords pw;
  for (int r = 0; r < 16; ++r) pw.p[r] = r < world ? static_cast<const uint32_t*>(peer_words[r]) : nullptr;
  union_words_kernel<<<(n_words + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(pw, n_words, world, out);
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

// peer-to-peer exchange: phase 1 = reduce-scatter, phase 2 = all-gather (a symmetric-memory barrier between them).
// `peer_buffers` = host array of `world` device pointers, entry r = rank r's buffer as mapped into this process.
int gsr_window_allreduce_p2p(void* const* peer_buffers, size_t n_floats, int rank, int world, int phase, void* stream) {
  return gsr_window_allreduce_p2p_rows(peer_buffers, n_floats, rank, world, phase, nullptr, stream);
}

int gsr_window_allreduce_p2p_rows(void* const* peer_buffers, size_t n_floats, int rank, int world, int phase,
                                  const gsr_window_rows* rows, void* stream) {
  RowMask m;
  if (!make_mask(rows, n_floats, m)) return GSR_ERR_INVALID;
  if (!peer_buffers || (n_floats & 3) || world < 1 || world > 16 || rank < 0 || rank >= world || (phase != 1 && phase != 2))
    return GSR_ERR_INVALID;
  PeerPtrs pp;
  for (int r = 0; r < 16; ++r) pp.p[r] = r < world ? static_cast<float4*>(peer_buffers[r]) : nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (phase == 1)
    reduce_scatter_p2p_kernel<<<grid_for(dev), 512, 0, s>>>(pp, n_floats / 4, rank, world, m);
  else
    all_gather_p2p_kernel<<<grid_for(dev), 512, 0, s>>>(pp, n_floats / 4, rank, world, m);
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

}  // extern "C"
