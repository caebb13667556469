// Tile-bucketed binning for sm_100a: the (tile | depth) sort of the reference done as an exact bucketing by
// tile followed by in-shared-memory radix sorts of bounded segments.
//
// Replaces, on the fast path, cub InclusiveSum + duplicateWithKeys + cub SortPairs + cudaMemset +
// identifyTileRanges (reference src/cuda_rasterizer/rasterizer_impl.cu:270-316, 64-125) with
//
//   preprocess              counts instances per tile (RED.ADD into tile_count)            [preprocess.cu]
//   tile_scan_kernel        one CTA: exclusive scan of the counts -> ranges (== identifyTileRanges' result),
//                           scatter cursors, launch order of the blend CTAs, depth buckets per long tile,
//                           {R, longest list, #long lists, overflow, #depth buckets}
//   scatter_instances       one thread per Gaussian: (depth bits, id) appended to each touched tile's segment
//                           through the tile's cursor (order inside a segment is arbitrary here)
//   partition_long_kernel   one CTA per LONG list (more than split_len instances): splitters from a sorted
//                           sample of the segment's depths cut it into K depth buckets of ~bucket_target
//                           instances; the segment is rewritten bucket by bucket (order inside a bucket still
//                           arbitrary). Any list length: the work is linear in it, and what follows is spread
//                           over all SMs.
//   segment_sort_kernel     one CTA per short list / per depth bucket: LSD radix sort by depth bits in shared
//                           memory (only over the bits that differ inside the segment), ties put in ascending-id
//                           order, then the outputs in final order: point_list_keys (tile << 32 | depth), point_list
//                           and the gathered 48-byte records. A bucket too large for shared memory (possible only
//                           with thousands of EQUAL depths in one tile, or lists of several 100 000 instances) is
//                           sorted by the same passes in global memory.
//
// Why this is bit-exact. The reference sorts instances by (tile, depth bits) with a STABLE radix sort of a
// list written in Gaussian-id order (rows of a Gaussian's rectangle are different tiles), so its result is
// the unique ascending order of the triple (tile, depth bits, id). Any correct sort of that triple gives the
// same arrays; stability and the unsorted intermediate order are artefacts. Buckets are ranges of depth bits
// (equal depths always share a bucket), laid out in ascending order inside the tile's range, so sorting every
// bucket by (depth bits, id) sorts the tile. Why it is faster at frame scale: a device-wide onesweep pass over
// ~1 M pairs is bound by its decoupled look-back chain (~20 us per pass, 6 passes + 2 set-up kernels for 43 key
// bits, ncu in profiles/); here no pass crosses a CTA boundary, and no list is sorted by a single SM.
#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr uint32_t kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------------------
// One CTA: ranges, cursors, launch order, depth buckets of the long lists, summary.
__global__ void __launch_bounds__(1024) tile_scan_kernel(int tiles, uint32_t* __restrict__ tile_count,
                                                        uint2* __restrict__ ranges, uint32_t* __restrict__ order,
                                                        uint32_t* __restrict__ tile_last,
                                                        uint32_t* __restrict__ bucket_base,
                                                        uint32_t* __restrict__ summary, uint32_t split_len,
                                                        uint32_t target, uint32_t r_cap) {
  __shared__ uint32_t s_warp[32], s_warp_b[32];
  __shared__ uint32_t s_count[33], s_base[33];
  __shared__ uint32_t s_carry, s_carry_b, s_max, s_long, s_overflow;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 33) s_count[tid] = 0;
  if (tid == 0) s_carry = 0, s_carry_b = 0, s_max = 0, s_long = 0, s_overflow = 0;
  __syncthreads();
  for (int base = 0; base < tiles; base += 1024) {
    const int t = base + tid;
    const uint32_t c = t < tiles ? tile_count[t] : 0u;
    const uint32_t kb = bucket_count(c, split_len, target);
    uint32_t incl = c, incl_b = kb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(kFull, incl, d), vb = __shfl_up_sync(kFull, incl_b, d);
      if (lane >= d) incl += v, incl_b += vb;
    }
    if (lane == 31) s_warp[warp] = incl, s_warp_b[warp] = incl_b;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane], wb = s_warp_b[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(kFull, w, d), vb = __shfl_up_sync(kFull, wb, d);
        if (lane >= d) w += v, wb += vb;
      }
      s_warp[lane] = w, s_warp_b[lane] = wb;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t off = s_carry + (warp ? s_warp[warp - 1] : 0u) + incl - c;
    const uint32_t off_b = s_carry_b + (warp ? s_warp_b[warp - 1] : 0u) + incl_b - kb;
    if (t < tiles) {
      ranges[t] = c ? make_uint2(off, off + c) : make_uint2(0u, 0u);  // untouched tiles stay (0,0): memset at :311
      tile_count[t] = off;                                            // from here on: the tile's scatter cursor
      tile_last[t] = 0;  // blend_fwd accumulates the tile's last contributor with atomicMax
      bucket_base[t] = off_b;
      if (kb) atomicAdd(&s_long, 1u);
      atomicMax(&s_max, c);
    }
    __syncthreads();
    if (tid == 0) s_carry += s_warp[31], s_carry_b += s_warp_b[31];
    __syncthreads();
  }
  // r_cap != ~0: sync-free forward into a presized binning buffer. If the frame does not fit (more instances than
  // the buffer holds) every range is emptied, so all later stages find nothing to do, and the overflow is
  // reported to the host with the other counters.
  if (tid == 0) {
    const uint32_t overflow = (r_cap != 0xffffffffu && s_carry > r_cap) ? 1u : 0u;
    summary[0] = s_carry, summary[1] = s_max, summary[2] = overflow ? 0u : s_long, summary[3] = overflow;
    summary[4] = overflow ? 0u : s_carry_b;
#pragma unroll
    for (int i = 5; i < 12; ++i) summary[i] = 0u;  // [6..9]: the blend forward's work counters
    s_overflow = overflow;
  }
  __syncthreads();
  const bool overflow = s_overflow != 0;
  if (overflow)
    for (int t = tid; t < tiles; t += 1024) ranges[t] = make_uint2(0u, 0u);
  // launch order of the per-tile CTAs: bucketed by floor(log2(length)), longest bucket first (on overflow every
  // list is empty: one bucket, identity order)
  for (int t = tid; t < tiles; t += 1024) {
    const uint2 r = ranges[t];  // this thread's own store above
    atomicAdd(&s_count[32 - __clz(r.y - r.x)], 1u);
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t acc = 0;
    for (int b = 32; b >= 0; --b) {
      s_base[b] = acc;
      acc += s_count[b];
    }
  }
  __syncthreads();
  for (int t = tid; t < tiles; t += 1024) {
    const uint2 r = ranges[t];
    order[atomicAdd(&s_base[32 - __clz(r.y - r.x)], 1u)] = (uint32_t)t;
  }
}

// ---------------------------------------------------------------------------------------------------------
// One thread per Gaussian: append (depth bits, id) to every touched tile's segment. The rectangle is
// recomputed from the stored pixel mean and the integer radius with the same float ops as in preprocess
// (getRect, auxiliary.h:39-46), exactly like duplicate_with_keys_kernel.
__global__ void __launch_bounds__(256) scatter_instances_kernel(int P, const SplatRecord* __restrict__ rec,
                                                               const int* __restrict__ radii,
                                                               uint32_t* __restrict__ cursor,
                                                               uint32_t* __restrict__ seg_keys,
                                                               uint32_t* __restrict__ seg_vals, int grid_x,
                                                               int grid_y, const uint32_t* __restrict__ summary) {
  if (summary[3]) return;  // presized binning buffer too small for this frame: nothing may be written
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int radius = idx < P ? radii[idx] : 0;
  int rminx = 0, rminy = 0, rw = 1;
  uint32_t ntiles = 0, depth_bits = 0;
  if (radius > 0) {
    const float4 q0 = rec[idx].q0;
    depth_bits = __float_as_uint(rec[idx].q2.w);
    const float rf = (float)radius;
    rminx = min(grid_x, max(0, __float2int_rz(__fmul_rn(__fadd_rn(q0.x, -rf), 0.0625f))));
    rminy = min(grid_y, max(0, __float2int_rz(__fmul_rn(__fadd_rn(q0.y, -rf), 0.0625f))));
    const int rmaxx =
        min(grid_x, max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(q0.x, rf), 16.0f), -1.0f), 0.0625f))));
    const int rmaxy =
        min(grid_y, max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(q0.y, rf), 16.0f), -1.0f), 0.0625f))));
    rw = rmaxx - rminx;
    ntiles = (uint32_t)rw * (uint32_t)(rmaxy - rminy);
  }
  // row-major walk of the Gaussian's tile rectangle without dividing: (tile, column) -> next
  const uint32_t t_first = (uint32_t)(rminy * grid_x + rminx);
  auto next_tile = [&](uint32_t& t, uint32_t& rx) {
    ++t;
    if (++rx == (uint32_t)rw) rx = 0, t += (uint32_t)(grid_x - rw);
  };
  // The 256 Gaussians of a CTA are neighbours in the map and fall into a few dozen tiles, and L2 serialises
  // returning atomics to one address (~15 ns each: a 10 000-instance list would cost its tile's cursor ~40 us even
  // with warp-merged requests). So the CTA first counts its appends per tile in a shared-memory hash table, one
  // thread per occupied slot then reserves the CTA's share of the tile's segment with ONE global atomic, and the
  // appends take their positions from the slot's shared-memory cursor. An append whose tile found no slot within
  // kProbes probes (never seen on real frames) goes to the global cursor directly; the table only grows, so the
  // same probe sequence fails again in the second walk and nothing is counted twice.
  constexpr uint32_t kSlots = 1024, kProbes = 16, kEmpty = 0xffffffffu;
  __shared__ __align__(16) uint32_t h_key[kSlots], h_cnt[kSlots];
  static_assert(kSlots == 4 * 256, "one 16-byte store per thread and table");
  reinterpret_cast<uint4*>(h_key)[threadIdx.x] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
  reinterpret_cast<uint4*>(h_cnt)[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  auto find = [&](uint32_t t, bool insert) -> int {
    uint32_t slot = (t * 2654435761u) >> 22;  // top 10 bits
    for (uint32_t pr = 0; pr < kProbes; ++pr, slot = (slot + 1) & (kSlots - 1)) {
      const uint32_t seen = insert ? atomicCAS(&h_key[slot], kEmpty, t) : h_key[slot];
      if (seen == t || (insert && seen == kEmpty)) return (int)slot;
    }
    return -1;
  };
  for (uint32_t k = 0, t = t_first, rx = 0; k < ntiles; ++k, next_tile(t, rx)) {
    const int slot = find(t, true);
    if (slot >= 0) atomicAdd(&h_cnt[slot], 1u);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < kSlots; i += blockDim.x) {
    const uint32_t c = h_cnt[i];
    if (c) h_cnt[i] = atomicAdd(&cursor[h_key[i]], c);  // from here on: the CTA's cursor inside the tile's segment
  }
  __syncthreads();
  for (uint32_t k = 0, t = t_first, rx = 0; k < ntiles; ++k, next_tile(t, rx)) {
    const int slot = find(t, false);
    const uint32_t pos = slot >= 0 ? atomicAdd(&h_cnt[slot], 1u) : atomicAdd(&cursor[t], 1u);
    seg_keys[pos] = depth_bits;
    seg_vals[pos] = (uint32_t)idx;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Depth partition of the long lists. The first summary[2] entries of `order` are exactly the lists with more than
// split_len instances (split_len + 1 is a power of two and `order` is bucketed by floor(log2(length))). One CTA per
// list, persistent over the lists:
//   1. ~16 depths per bucket, sampled at a regular stride over the (arbitrarily ordered) segment, are sorted
//      (bitonic network in shared memory); every 32nd of them is a splitter;
//   2. every instance finds its bucket (binary search over the K - 1 splitters: equal depths share a bucket) and is
//      counted with a shared-memory atomic;
//   3. bucket offsets by a scan; one descriptor {first instance, instances, tile} per bucket;
//   4. every instance is written to its bucket's part of the second segment buffer.
// With 16-fold oversampling a bucket's size has a relative spread of ~25 %: kSegCap is more than eight standard
// deviations above the default target. One that is larger anyway (thousands of equal depths cannot be separated)
// is sorted by segment_sort_kernel's global-memory path.
struct PartitionArgs {
  const uint2* ranges;
  const uint32_t* order;
  const uint32_t* summary;
  const uint32_t* bucket_base;
  const uint32_t* seg_keys;
  const uint32_t* seg_vals;
  uint32_t* part_keys;
  uint32_t* part_vals;
  uint4* seg_desc;
  uint32_t split_len, target;
};

constexpr int kPartThreads = 1024, kMaxSample = 2048, kOversample = 16, kPartUnroll = 4;

__global__ void __launch_bounds__(kPartThreads) partition_long_kernel(const PartitionArgs a) {
  __shared__ uint32_t s_sample[kMaxSample];
  __shared__ uint32_t s_split[kMaxBuckets], s_cnt[kMaxBuckets], s_base[kMaxBuckets];
  __shared__ uint32_t s_wsum[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_long = a.summary[2];
  for (uint32_t li = blockIdx.x; li < n_long; li += gridDim.x) {
    const uint32_t tile = a.order[li];
    const uint2 range = a.ranges[tile];
    const uint32_t n = range.y - range.x;
    const uint32_t K = bucket_count(n, a.split_len, a.target);
    if (K == 0) continue;  // (cannot happen: see above)
    const uint32_t* keys = a.seg_keys + range.x;
    const uint32_t* vals = a.seg_vals + range.x;
    uint32_t m = (uint32_t)kOversample * K;
    m = m > (uint32_t)kMaxSample ? (uint32_t)kMaxSample : m;
    m = m < n ? m : n;                       // samples
    uint32_t m_pad = 64;                     // power of two for the network, padded with the largest key
    while (m_pad < m) m_pad <<= 1;
    __syncthreads();  // the previous list of this CTA is done with the shared arrays
    for (uint32_t j = tid; j < m_pad; j += kPartThreads)
      s_sample[j] = j < m ? keys[(uint32_t)(((uint64_t)j * n) / m)] : 0xffffffffu;
    if ((uint32_t)tid < K) s_cnt[tid] = 0;
    __syncthreads();
    for (uint32_t size = 2; size <= m_pad; size <<= 1) {
      for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
        for (uint32_t t = tid; t < (m_pad >> 1); t += kPartThreads) {
          const uint32_t i = 2 * t - (t & (stride - 1));  // lower index of the pair
          const uint32_t j = i + stride;
          const uint32_t x = s_sample[i], y = s_sample[j];
          const bool up = (i & size) == 0;
          if ((x > y) == up) s_sample[i] = y, s_sample[j] = x;
        }
        __syncthreads();
      }
    }
    if (tid >= 1 && (uint32_t)tid < K) s_split[tid - 1] = s_sample[(uint32_t)(((uint64_t)tid * m) / K)];
    __syncthreads();
    auto bucket_of = [&](uint32_t key) -> uint32_t {  // number of splitters <= key
      uint32_t lo = 0, hi = K - 1;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s_split[mid] <= key) lo = mid + 1;
        else hi = mid;
      }
      return lo;
    };
    // count (shared-memory atomics: ~4 SM-cycles per warp, MATCH.ANY would be ~58)
    for (uint32_t base = 0; base < n; base += kPartUnroll * kPartThreads) {
      uint32_t k[kPartUnroll];
      bool ok[kPartUnroll];
#pragma unroll
      for (int u = 0; u < kPartUnroll; ++u) {
        const uint32_t i = base + u * kPartThreads + tid;
        ok[u] = i < n;
        k[u] = ok[u] ? keys[i] : 0u;
      }
#pragma unroll
      for (int u = 0; u < kPartUnroll; ++u)
        if (ok[u]) atomicAdd(&s_cnt[bucket_of(k[u])], 1u);
    }
    __syncthreads();
    // exclusive scan of the K <= 256 counts; descriptors; counts become cursors
    uint32_t v = 0, incl = 0;
    if (tid < kMaxBuckets) {
      v = (uint32_t)tid < K ? s_cnt[tid] : 0u;
      incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += t;
      }
      if (lane == 31) s_wsum[warp] = incl;
    }
    __syncthreads();
    if (tid < kMaxBuckets) {
      uint32_t off = incl - v;
      for (int w = 0; w < warp; ++w) off += s_wsum[w];
      s_base[tid] = off;
      if ((uint32_t)tid < K) {
        a.seg_desc[a.bucket_base[tile] + tid] = make_uint4(range.x + off, v, tile, 0u);
        s_cnt[tid] = 0;
      }
    }
    __syncthreads();
    // scatter (the order inside a bucket is arbitrary: whatever the atomics hand out)
    for (uint32_t base = 0; base < n; base += kPartUnroll * kPartThreads) {
      uint32_t k[kPartUnroll], id[kPartUnroll];
      bool ok[kPartUnroll];
#pragma unroll
      for (int u = 0; u < kPartUnroll; ++u) {
        const uint32_t i = base + u * kPartThreads + tid;
        ok[u] = i < n;
        k[u] = ok[u] ? keys[i] : 0u;
        id[u] = ok[u] ? vals[i] : 0u;
      }
#pragma unroll
      for (int u = 0; u < kPartUnroll; ++u) {
        if (ok[u]) {
          const uint32_t b = bucket_of(k[u]);
          const uint32_t pos = range.x + s_base[b] + atomicAdd(&s_cnt[b], 1u);
          a.part_keys[pos] = k[u];
          a.part_vals[pos] = id[u];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
struct SegmentSortArgs {
  const uint2* ranges;
  const uint32_t* order;
  const uint32_t* summary;  // [4] = number of depth buckets
  const uint4* seg_desc;
  uint32_t* seg_keys;   // short lists: (depth bits, id) as scattered
  uint32_t* seg_vals;
  uint32_t* part_keys;  // depth buckets of the long lists
  uint32_t* part_vals;
  const SplatRecord* rec;
  uint64_t* keys_out;
  uint32_t* vals_out;
  SplatRecord* inst;
  uint32_t split_len;
};

// Stable LSD pass over bits [lo, lo + nbits) of `kin` (nbits <= kBits): (kin, vin) -> (kout, vout).
// Every warp owns a contiguous chunk; digits are counted with shared-memory atomics into the warp's own histogram,
// ranks across chunks come from the digit-major scan of the per-warp histograms, ranks inside a group of 32
// consecutive elements from one ballot per digit bit. (MATCH.ANY would do the latter in one instruction, but it costs
// ~58 SM-cycles per warp on sm_100a against ~4 for a shared-memory atomic and ~3 per ballot: tools/micro/match_bench.cu.)
// The arrays may live in shared or in global memory.
template <int kWarps, int kBits>
__device__ __forceinline__ void radix_pass(const uint32_t* kin, const uint32_t* vin, uint32_t* kout, uint32_t* vout,
                                           int n, int lo, int nbits, bool by_val, uint32_t (*hist)[1 << kBits],
                                           uint32_t* s_scan) {
  constexpr int kDigits = 1 << kBits;
  static_assert(kDigits <= kWarps * 32, "one thread per digit in the scan");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t dmask = (1u << nbits) - 1u, lt = (1u << lane) - 1u;
  const int chunk = (((n + kWarps - 1) / kWarps) + 31) & ~31;
  const int beg = min(n, warp * chunk), end = min(n, beg + chunk);
  const uint32_t* src = by_val ? vin : kin;
  for (int i = tid; i < kWarps * kDigits; i += kWarps * 32) hist[0][i] = 0;
  __syncthreads();
  // count
  for (int i = beg + lane; i < end; i += 32) atomicAdd(&hist[warp][(src[i] >> lo) & dmask], 1u);
  __syncthreads();
  // digit-major exclusive scan: thread d walks the warps of digit d, then the digit totals are scanned
  if (tid < kDigits) {
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t c = hist[w][tid];
      hist[w][tid] = s;
      s += c;
    }
    uint32_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) s_scan[warp] = incl;
    s_scan[32 + tid] = incl - s;  // exclusive inside the warp
  }
  __syncthreads();
  if (tid < kDigits) {
    uint32_t base = s_scan[32 + tid];
    for (int w = 0; w < warp; ++w) base += s_scan[w];  // warp < kDigits / 32
#pragma unroll
    for (int w = 0; w < kWarps; ++w) hist[w][tid] += base;
  }
  __syncthreads();
  // scatter
  for (int i = beg + lane; i - lane < end; i += 32) {
    const bool valid = i < end;
    const uint32_t k = valid ? kin[i] : 0u, v = valid ? vin[i] : 0u;
    const uint32_t d = ((by_val ? v : k) >> lo) & dmask;
    uint32_t peers = __ballot_sync(kFull, valid);  // lanes of this group with the same digit
    for (int b = 0; b < nbits; ++b) {
      const bool bit = (d >> b) & 1u;
      const uint32_t m = __ballot_sync(kFull, bit);
      peers &= bit ? m : ~m;
    }
    const uint32_t rank = __popc(peers & lt);
    uint32_t pos = 0;
    if (valid) pos = hist[warp][d] + rank;
    __syncwarp();  // every lane has read its digit's cursor
    if (valid && rank == 0u) hist[warp][d] = pos + __popc(peers);
    if (valid) kout[pos] = k, vout[pos] = v;
    __syncwarp();
  }
  __syncthreads();
}

constexpr int kSortThreads = 256, kSortWarps = kSortThreads / 32, kSortBits = 8;

// Sort one segment of n (depth bits, id) pairs found at src_k / src_v, destined for instances [start, start + n)
// of `tile`, and write the final arrays. n <= kSegCap: in shared memory. Larger (an oversized depth bucket): the
// same passes ping-pong between the segment itself and its own part of keys_out (8 B per instance = one
// (depth, id) pair), which is only overwritten with final keys after the sorted pairs are back in the segment.
__device__ __forceinline__ void sort_segment(const SegmentSortArgs& a, uint32_t* src_k, uint32_t* src_v, uint32_t start,
                                             int n, uint32_t tile, unsigned char* smem_raw) {
  uint32_t* k0 = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* v0 = k0 + kSegCap;
  uint32_t* k1 = v0 + kSegCap;
  uint32_t* v1 = k1 + kSegCap;
  uint32_t(*hist)[1 << kSortBits] = reinterpret_cast<uint32_t(*)[1 << kSortBits]>(v1 + kSegCap);
  __shared__ uint32_t s_scan[32 + (1 << kSortBits)];
  __shared__ uint32_t s_or_k, s_or_v, s_flag;

  const int tid = threadIdx.x;
  const bool in_smem = n <= kSegCap;
  if (tid == 0) s_or_k = 0, s_or_v = 0, s_flag = 0;
  __syncthreads();
  {
    const uint32_t kf = src_k[0], vf = src_v[0];
    uint32_t ok = 0, ov = 0;
    for (int i = tid; i < n; i += kSortThreads) {
      const uint32_t k = src_k[i], v = src_v[i];
      if (in_smem) k0[i] = k, v0[i] = v;
      ok |= k ^ kf, ov |= v ^ vf;
    }
    ok = __reduce_or_sync(kFull, ok), ov = __reduce_or_sync(kFull, ov);
    if ((tid & 31) == 0) {
      if (ok) atomicOr(&s_or_k, ok);
      if (ov) atomicOr(&s_or_v, ov);
    }
  }
  __syncthreads();
  const uint32_t diff_k = s_or_k, diff_v = s_or_v;
  if (!in_smem) {
    k0 = src_k, v0 = src_v;
    k1 = reinterpret_cast<uint32_t*>(a.keys_out + start);
    v1 = k1 + n;
  }

  uint32_t *kc = k0, *vc = v0, *kn = k1, *vn = v1;
  auto sort_bits = [&](uint32_t diff, bool by_val) {  // LSD over the bits that differ inside the segment
    if (diff == 0) return;
    const int lo0 = __ffs(diff) - 1, hi = 32 - __clz(diff);
    const int passes = (hi - lo0 + kSortBits - 1) / kSortBits;
    const int width = (hi - lo0 + passes - 1) / passes;
    for (int p = 0, lo = lo0; p < passes; ++p, lo += width) {
      radix_pass<kSortWarps, kSortBits>(kc, vc, kn, vn, n, lo, min(width, hi - lo), by_val, hist, s_scan);
      uint32_t* t = kc;
      kc = kn, kn = t;
      t = vc, vc = vn, vn = t;
    }
  };
  sort_bits(diff_k, false);
  // Equal depths must come out in ascending id order (the reference's stable sort of an id-ordered list);
  // the scatter order inside a segment is arbitrary. Lattice rows at the same view depth make such runs
  // common on voxelised maps, so this is done in parallel: every warp finds the run bounds of its contiguous
  // share with ballots, then every element ranks its id inside its run. A run longer than kRunMax makes the
  // CTA redo the segment as a two-key LSD sort (ids first).
  constexpr int kRunMax = 256;
  {
    const int lane = tid & 31, warp = tid >> 5;
    const int chunk = (((n + kSortWarps - 1) / kSortWarps) + 31) & ~31;
    const int beg = min(n, warp * chunk), end = min(n, beg + chunk);
    uint32_t* run_lo = kn;  // the spare key/value buffers hold the bounds
    uint32_t* run_hi = vn;
    if (beg < end) {
      int carry = beg;  // start of the run that contains element beg
      {
        const uint32_t k = kc[beg];
        while (carry > 0 && beg - carry <= kRunMax && kc[carry - 1] == k) --carry;
      }
      for (int c = beg; c < end; c += 32) {
        const int i = c + lane;
        const bool valid = i < end;
        const bool is_start = valid && (i == 0 || kc[i - 1] != kc[i]);
        const uint32_t b = __ballot_sync(kFull, is_start);
        const uint32_t mine = b & (kFull >> (31 - lane));
        if (valid) run_lo[i] = (uint32_t)(mine ? c + 31 - __clz(mine) : carry);
        if (b) carry = c + 31 - __clz(b);
      }
      int carry_hi = end;  // one past the run that contains element end - 1
      {
        const uint32_t k = kc[end - 1];
        while (carry_hi < n && carry_hi - end <= kRunMax && kc[carry_hi] == k) ++carry_hi;
      }
      for (int c = beg + ((end - beg - 1) & ~31); c >= beg; c -= 32) {
        const int i = c + lane;
        const bool valid = i < end;
        const bool is_last = valid && (i == n - 1 || kc[i + 1] != kc[i]);
        const uint32_t b = __ballot_sync(kFull, is_last);
        const uint32_t mine = b & (kFull << lane);
        if (valid) run_hi[i] = (uint32_t)(mine ? c + __ffs(mine) : carry_hi);
        if (b) carry_hi = c + __ffs(b);
      }
    }
    __syncthreads();
    for (int i = tid; i < n; i += kSortThreads) {
      const int lo = (int)run_lo[i], hi = (int)run_hi[i];
      int pos = i;
      if (hi - lo > kRunMax) {
        s_flag = 1;
      } else if (hi - lo > 1) {
        const uint32_t v = vc[i];
        int rank = 0;
        for (int j = lo; j < hi; ++j) rank += vc[j] < v;
        pos = lo + rank;
      }
      run_lo[i] = (uint32_t)pos;
    }
    __syncthreads();
    if (!s_flag) {
      for (int i = tid; i < n; i += kSortThreads) run_hi[run_lo[i]] = vc[i];
      __syncthreads();
      uint32_t* t = vc;  // ids now live in the spare value buffer; keys are unchanged (equal inside a run)
      vc = vn, vn = t;
    } else {
      sort_bits(diff_v, true);
      sort_bits(diff_k, false);
    }
  }
  if (!in_smem) {
    // the sorted pairs must not sit in keys_out's own memory while it is overwritten with the final keys
    if (kc != src_k) {
      for (int i = tid; i < n; i += kSortThreads) src_k[i] = kc[i];
      __syncthreads();
      kc = src_k;
    }
    if (vc != src_v) {
      for (int i = tid; i < n; i += kSortThreads) src_v[i] = vc[i];
      __syncthreads();
      vc = src_v;
    }
  }

  // outputs in final order
  const uint64_t hi_key = (uint64_t)tile << 32;
  for (int i = tid; i < n; i += kSortThreads) {
    const size_t g = (size_t)start + i;
    a.keys_out[g] = hi_key | kc[i];
    a.vals_out[g] = vc[i];
  }
  auto emit = [&](int i, const float4& q0, const float4& q1, const float4& q2) {
    float4* dst = reinterpret_cast<float4*>(a.inst + (size_t)start + i);
    dst[0] = q0, dst[1] = q1, dst[2] = q2;
  };
  for (int i = tid; i < n; i += 2 * kSortThreads) {  // two independent record gathers in flight per thread
    const int i2 = i + kSortThreads;
    const uint32_t id = vc[i], id2 = i2 < n ? vc[i2] : id;
    const float4* src = reinterpret_cast<const float4*>(a.rec + id);
    const float4* src2 = reinterpret_cast<const float4*>(a.rec + id2);
    const float4 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2);
    const float4 r0 = __ldg(src2), r1 = __ldg(src2 + 1), r2 = __ldg(src2 + 2);
    emit(i, q0, q1, q2);
    if (i2 < n) emit(i2, r0, r1, r2);
  }
}

// kBuckets = false: CTA b sorts the list of tile order[b] if it is short (<= split_len instances);
// kBuckets = true:  the CTAs stride over the depth buckets of the long lists (partition_long_kernel).
template <bool kBuckets>
__global__ void __launch_bounds__(kSortThreads) segment_sort_kernel(const SegmentSortArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (kBuckets) {
    const uint32_t nb = a.summary[4];
    for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {  // persistent: the grid covers the SMs, not the buckets
      const uint4 d = a.seg_desc[b];
      if (d.y == 0) continue;
      __syncthreads();  // the previous segment of this CTA is completely written out
      sort_segment(a, a.part_keys + d.x, a.part_vals + d.x, d.x, (int)d.y, d.z, smem_raw);
    }
  } else {
    const uint32_t tile = a.order[blockIdx.x];
    const uint2 range = a.ranges[tile];
    const uint32_t n = range.y - range.x;
    if (n == 0 || n > a.split_len) return;
    sort_segment(a, a.seg_keys + range.x, a.seg_vals + range.x, range.x, (int)n, tile, smem_raw);
  }
}

constexpr size_t kSortSmem = (size_t)kSegCap * 16 + (size_t)kSortWarps * (1 << kSortBits) * 4;

}  // namespace

cudaError_t launch_tile_scan(int tiles, const ImageState& img, uint32_t r_cap, cudaStream_t s) {
  const Tuning t = tuning();
  tile_scan_kernel<<<1, 1024, 0, s>>>(tiles, img.tile_count, img.ranges, img.tile_order, img.tile_last, img.bucket_base,
                                      img.summary, (uint32_t)t.split_len, (uint32_t)t.bucket_target, r_cap);
  return cudaGetLastError();
}

cudaError_t launch_scatter_instances(int P, const GeomState& g, const int* radii, const ImageState& img,
                                     const BinningState& b, int grid_x, int grid_y, cudaStream_t s) {
  uint32_t* seg_keys = reinterpret_cast<uint32_t*>(b.keys_unsorted);
  scatter_instances_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.rec, radii, img.tile_count, seg_keys, b.vals_unsorted,
                                                          grid_x, grid_y, img.summary);
  return cudaGetLastError();
}

cudaError_t launch_tile_sort(int tiles, int n_long, int n_buckets, int r_cap, const ImageState& img,
                             const BinningState& b, const GeomState& g, int grid_x, cudaStream_t s) {
  (void)grid_x;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(segment_sort_kernel<false>), (int)kSortSmem);
  if (e == cudaSuccess) e = ensure_dynamic_smem(reinterpret_cast<const void*>(segment_sort_kernel<true>), (int)kSortSmem);
  if (e != cudaSuccess) return e;
  const Tuning t = tuning();
  const int sms = device_sm_count();
  SegmentSortArgs a;
  a.ranges = img.ranges, a.order = img.tile_order, a.summary = img.summary, a.seg_desc = b.seg_desc;
  a.seg_keys = reinterpret_cast<uint32_t*>(b.keys_unsorted), a.seg_vals = b.vals_unsorted;
  a.part_keys = a.seg_keys + (size_t)r_cap, a.part_vals = b.vals_tmp;
  a.rec = g.rec, a.keys_out = b.keys, a.vals_out = b.vals, a.inst = b.inst, a.split_len = (uint32_t)t.split_len;
  // Two independent chains: the short lists are sorted on a side stream while the long ones are partitioned and
  // their depth buckets sorted on `s`. n_long / n_buckets < 0: not known on the host (sync-free forward) -- the
  // kernels read the counts on the device and the (persistent) grids cover the SMs, not the work.
  struct Side {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
  };
  static thread_local Side side[64];
  int dev = 0;
  cudaGetDevice(&dev);
  Side* sd = nullptr;
  if (n_long != 0 && dev >= 0 && dev < 64) {
    sd = &side[dev];
    if (!sd->stream) {
      const bool ok = cudaEventCreateWithFlags(&sd->fork, cudaEventDisableTiming) == cudaSuccess &&
                      cudaStreamCreateWithFlags(&sd->stream, cudaStreamNonBlocking) == cudaSuccess &&
                      cudaEventCreateWithFlags(&sd->join, cudaEventDisableTiming) == cudaSuccess;
      if (!ok) {
        sd->stream = nullptr;
        sd = nullptr;
        (void)cudaGetLastError();
      }
    }
  }
  cudaStream_t short_stream = s;
  if (sd) {
    e = cudaEventRecord(sd->fork, s);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(sd->stream, sd->fork, 0);
    if (e != cudaSuccess) return e;
    short_stream = sd->stream;
  }
  if (n_long != 0) {
    PartitionArgs p;
    p.ranges = img.ranges, p.order = img.tile_order, p.summary = img.summary, p.bucket_base = img.bucket_base;
    p.seg_keys = a.seg_keys, p.seg_vals = a.seg_vals, p.part_keys = a.part_keys, p.part_vals = a.part_vals;
    p.seg_desc = b.seg_desc, p.split_len = (uint32_t)t.split_len, p.target = (uint32_t)t.bucket_target;
    const int lists = n_long > 0 ? n_long : tiles;
    const int pgrid = lists < 2 * sms ? lists : 2 * sms;  // two 1024-thread CTAs per SM
    partition_long_kernel<<<pgrid, kPartThreads, 0, s>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t worst = max_depth_buckets((size_t)r_cap), resident = (size_t)5 * sms;  // 5 CTAs of 40 KB per SM
    size_t bgrid = n_buckets > 0 ? (size_t)n_buckets : worst;
    bgrid = bgrid < resident ? bgrid : resident;
    segment_sort_kernel<true><<<(unsigned)bgrid, kSortThreads, kSortSmem, s>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  segment_sort_kernel<false><<<tiles, kSortThreads, kSortSmem, short_stream>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (sd) {
    e = cudaEventRecord(sd->join, sd->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s, sd->join, 0);
  }
  return e;
}

}  // namespace gsr
