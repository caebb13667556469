// Tile-bucketed binning for sm_100a: the (tile | depth) sort of the reference done as an exact bucketing by
// tile followed by an in-shared-memory radix sort of every tile's segment.
//
// Replaces, on the fast path, cub InclusiveSum + duplicateWithKeys + cub SortPairs + cudaMemset +
// identifyTileRanges (reference src/cuda_rasterizer/rasterizer_impl.cu:270-316, 64-125) with
//
//   preprocess              counts instances per tile (RED.ADD into tile_count)            [preprocess.cu]
//   tile_scan_kernel        one CTA: exclusive scan of the counts -> ranges (== identifyTileRanges' result),
//                           scatter cursors, launch order of the blend CTAs, {R, longest list, #large lists}
//   scatter_instances       one thread per Gaussian: (depth bits, id) appended to each touched tile's segment
//                           through the tile's cursor (order inside a segment is arbitrary here)
//   tile_sort_kernel        one CTA per tile: LSD radix sort of the segment by depth bits in shared memory
//                           (only over the bits that differ inside the segment), ties put in ascending-id
//                           order, then the outputs in final order: point_list_keys (tile << 32 | depth),
//                           point_list and the gathered 48-byte records.
//
// Why this is bit-exact. The reference sorts instances by (tile, depth bits) with a STABLE radix sort of a
// list written in Gaussian-id order (rows of a Gaussian's rectangle are different tiles), so its result is
// the unique ascending order of the triple (tile, depth bits, id). Any correct sort of that triple gives the
// same arrays; stability and the unsorted intermediate order are artefacts. Why it is faster at frame scale:
// a device-wide onesweep pass over ~1 M pairs is bound by its decoupled look-back chain (~20 us per pass,
// 6 passes + 2 set-up kernels for 43 key bits, ncu in profiles/); here no pass crosses a CTA boundary.
//
// Lists longer than kLargeCap do not fit one CTA's shared memory: gsr_forward then takes the device-wide
// path (binning.cu), which is the bandwidth-efficient choice at those sizes anyway.
#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr uint32_t kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------------------
// One CTA: ranges, cursors, launch order, summary.
__global__ void __launch_bounds__(1024) tile_scan_kernel(int tiles, uint32_t* __restrict__ tile_count,
                                                        uint2* __restrict__ ranges, uint32_t* __restrict__ order,
                                                        uint32_t* __restrict__ tile_last,
                                                        uint32_t* __restrict__ summary, uint32_t small_cap,
                                                        uint32_t r_cap, uint32_t list_cap) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_count[33], s_base[33];
  __shared__ uint32_t s_carry, s_max, s_large;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 33) s_count[tid] = 0;
  if (tid == 0) s_carry = 0, s_max = 0, s_large = 0;
  __syncthreads();
  for (int base = 0; base < tiles; base += 1024) {
    const int t = base + tid;
    const uint32_t c = t < tiles ? tile_count[t] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(kFull, w, d);
        if (lane >= d) w += v;
      }
      s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t off = s_carry + (warp ? s_warp[warp - 1] : 0u) + incl - c;
    if (t < tiles) {
      ranges[t] = c ? make_uint2(off, off + c) : make_uint2(0u, 0u);  // untouched tiles stay (0,0): memset at :311
      tile_count[t] = off;                                            // from here on: the tile's scatter cursor
      tile_last[t] = 0;  // blend_fwd accumulates the tile's last contributor with atomicMax
      atomicAdd(&s_count[32 - __clz(c)], 1u);
      if (c > small_cap) atomicAdd(&s_large, 1u);
      atomicMax(&s_max, c);
    }
    __syncthreads();
    if (tid == 0) s_carry += s_warp[31];
    __syncthreads();
  }
  // launch order of the per-tile CTAs: bucketed by floor(log2(length)), longest bucket first
  if (tid == 0) {
    uint32_t acc = 0;
    for (int b = 32; b >= 0; --b) {
      s_base[b] = acc;
      acc += s_count[b];
    }
    // r_cap != ~0: sync-free forward into a presized binning buffer. If the frame does not fit (more instances
    // than the buffer holds, or a list too long for the per-tile sort) every range is emptied, so all later
    // stages find nothing to do, and the overflow is reported to the host with the other counters.
    const uint32_t overflow = r_cap == 0xffffffffu ? 0u : (s_carry > r_cap ? 1u : (s_max > list_cap ? 2u : 0u));
    summary[0] = s_carry, summary[1] = s_max, summary[2] = overflow ? 0u : s_large, summary[3] = overflow;
    s_large = overflow;
  }
  __syncthreads();
  if (s_large != 0 && r_cap != 0xffffffffu) {  // (s_large now holds the overflow code)
    for (int t = tid; t < tiles; t += 1024) ranges[t] = make_uint2(0u, 0u);
    __syncthreads();
  }
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
  auto tile_of = [&](uint32_t k) -> uint32_t {
    const int ry = (int)k / rw;
    return (uint32_t)((rminy + ry) * grid_x + rminx + ((int)k - ry * rw));
  };
  // The 256 Gaussians of a CTA are neighbours in the map and fall into a few dozen tiles, and L2 serialises
  // returning atomics to one address (~15 ns each: a 10 000-instance list would cost its tile's cursor ~40 us even
  // with warp-merged requests). So the CTA first counts its appends per tile in a shared-memory hash table, one
  // thread per occupied slot then reserves the CTA's share of the tile's segment with ONE global atomic, and the
  // appends take their positions from the slot's shared-memory cursor. An append whose tile found no slot within
  // kProbes probes (never seen on real frames) goes to the global cursor directly; the table only grows, so the
  // same probe sequence fails again in the second walk and nothing is counted twice.
  constexpr uint32_t kSlots = 1024, kProbes = 16, kEmpty = 0xffffffffu;
  __shared__ uint32_t h_key[kSlots], h_cnt[kSlots];
  for (uint32_t i = threadIdx.x; i < kSlots; i += blockDim.x) h_key[i] = kEmpty, h_cnt[i] = 0;
  __syncthreads();
  auto find = [&](uint32_t t, bool insert) -> int {
    uint32_t slot = (t * 2654435761u) >> 22;  // top 10 bits
    for (uint32_t pr = 0; pr < kProbes; ++pr, slot = (slot + 1) & (kSlots - 1)) {
      const uint32_t seen = insert ? atomicCAS(&h_key[slot], kEmpty, t) : h_key[slot];
      if (seen == t || (insert && seen == kEmpty)) return (int)slot;
    }
    return -1;
  };
  for (uint32_t k = 0; k < ntiles; ++k) {
    const int slot = find(tile_of(k), true);
    if (slot >= 0) atomicAdd(&h_cnt[slot], 1u);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < kSlots; i += blockDim.x) {
    const uint32_t c = h_cnt[i];
    if (c) h_cnt[i] = atomicAdd(&cursor[h_key[i]], c);  // from here on: the CTA's cursor inside the tile's segment
  }
  __syncthreads();
  for (uint32_t k = 0; k < ntiles; ++k) {
    const uint32_t t = tile_of(k);
    const int slot = find(t, false);
    const uint32_t pos = slot >= 0 ? atomicAdd(&h_cnt[slot], 1u) : atomicAdd(&cursor[t], 1u);
    seg_keys[pos] = depth_bits;
    seg_vals[pos] = (uint32_t)idx;
  }
}

struct TileSortArgs {
  const uint2* ranges;
  const uint32_t* order;
  const uint32_t* summary;  // [2] = number of long lists (they come first in `order`)
  const uint32_t* seg_keys;
  const uint32_t* seg_vals;
  const SplatRecord* rec;
  uint64_t* keys_out;
  uint32_t* vals_out;
  SplatRecord* inst;
  int grid_x;
  uint32_t min_len, max_len;  // this launch handles lists with min_len < length <= max_len
};

template <int kBits>
struct DigitCount {
  using type = uint32_t;
};
template <>
struct DigitCount<9> {
  using type = uint16_t;
};

// Stable LSD pass over bits [lo, lo + nbits) of `kin` (nbits <= kBits): (kin, vin) -> (kout, vout).
// Every warp owns a contiguous chunk; ranks inside a chunk come from match_any peers, ranks across chunks
// from the digit-major scan of the per-warp histograms. The 1024-thread classes use 9-bit digits with 16-bit
// counters (positions stay below kLargeCap < 65536): 27 depth bits in three passes instead of four, in the same
// 32 KB of histograms.
template <int kWarps, int kBits, typename Count>
__device__ __forceinline__ void radix_pass(const uint32_t* kin, const uint32_t* vin, uint32_t* kout, uint32_t* vout,
                                           int n, int lo, int nbits, bool by_val, Count (*hist)[1 << kBits],
                                           uint32_t* s_scan) {
  constexpr int kDigits = 1 << kBits;
  static_assert(kDigits <= kWarps * 32, "one thread per digit in the scan");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t dmask = (1u << nbits) - 1u;
  const int chunk = (((n + kWarps - 1) / kWarps) + 31) & ~31;
  const int beg = min(n, warp * chunk), end = min(n, beg + chunk);
  const uint32_t* src = by_val ? vin : kin;
  for (int i = tid; i < kWarps * kDigits; i += kWarps * 32) hist[0][i] = 0;
  __syncthreads();
  // count
  for (int i = beg + lane; i - lane < end; i += 32) {
    const bool valid = i < end;
    const uint32_t d = valid ? ((src[i] >> lo) & dmask) : 0xffffffffu;
    const uint32_t peers = __match_any_sync(kFull, d);
    if (valid && lane == __ffs(peers) - 1) hist[warp][d] = (Count)(hist[warp][d] + __popc(peers));
    __syncwarp();
  }
  __syncthreads();
  // digit-major exclusive scan: thread d walks the warps of digit d, then the digit totals are scanned
  if (tid < kDigits) {
    uint32_t s = 0;
#pragma unroll 4
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t c = hist[w][tid];
      hist[w][tid] = (Count)s;
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
#pragma unroll 4
    for (int w = 0; w < kWarps; ++w) hist[w][tid] = (Count)(hist[w][tid] + base);
  }
  __syncthreads();
  // scatter
  for (int i = beg + lane; i - lane < end; i += 32) {
    const bool valid = i < end;
    const uint32_t k = valid ? kin[i] : 0u, v = valid ? vin[i] : 0u;
    const uint32_t d = valid ? (((by_val ? v : k) >> lo) & dmask) : 0xffffffffu;
    const uint32_t peers = __match_any_sync(kFull, d);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (valid && lane == leader) {
      base = hist[warp][d];
      hist[warp][d] = (Count)(base + __popc(peers));
    }
    base = __shfl_sync(kFull, base, leader);
    if (valid) {
      const uint32_t pos = base + __popc(peers & ((1u << lane) - 1u));
      kout[pos] = k;
      vout[pos] = v;
    }
    __syncwarp();
  }
  __syncthreads();
}

// kGather: also gather the 48-byte records into sorted order. The long-list classes leave that to
// gather_long_kernel: ~1 MB of random 16-byte accesses per 10 000-instance list is bound by the bandwidth of the ONE
// SM that sorts the list (~20 us of the CTA's ~60), while all SMs together do it for every long list in a few us.
template <int kCap, int kThreads, bool kGather>
__device__ __forceinline__ void sort_one_tile(const TileSortArgs& a, uint32_t tile, unsigned char* smem_raw) {
  constexpr int kWarps = kThreads / 32;
  constexpr int kBits = kThreads >= 512 ? 9 : 8;  // digit width (see radix_pass)
  using Count = typename DigitCount<kBits>::type;
  uint32_t* k0 = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* v0 = k0 + kCap;
  uint32_t* k1 = v0 + kCap;
  uint32_t* v1 = k1 + kCap;
  Count(*hist)[1 << kBits] = reinterpret_cast<Count(*)[1 << kBits]>(v1 + kCap);
  __shared__ uint32_t s_scan[32 + (1 << kBits)];
  __shared__ uint32_t s_or_k, s_or_v, s_flag;

  const int tid = threadIdx.x;
  const uint2 range = a.ranges[tile];
  const int n = (int)(range.y - range.x);
  if ((uint32_t)n <= a.min_len || (uint32_t)n > a.max_len) return;

  __syncthreads();  // the previous tile of this CTA is completely written out
  if (tid == 0) s_or_k = 0, s_or_v = 0, s_flag = 0;
  __syncthreads();
  {
    const uint32_t kf = a.seg_keys[range.x], vf = a.seg_vals[range.x];
    uint32_t ok = 0, ov = 0;
    for (int i = tid; i < n; i += kThreads) {
      const uint32_t k = a.seg_keys[range.x + i], v = a.seg_vals[range.x + i];
      k0[i] = k, v0[i] = v;
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

  uint32_t *kc = k0, *vc = v0, *kn = k1, *vn = v1;
  auto sort_bits = [&](uint32_t diff, bool by_val) {  // LSD over the bits that differ inside the segment
    if (diff == 0) return;
    const int lo0 = __ffs(diff) - 1, hi = 32 - __clz(diff);
    const int passes = (hi - lo0 + kBits - 1) / kBits;
    const int width = (hi - lo0 + passes - 1) / passes;
    for (int p = 0, lo = lo0; p < passes; ++p, lo += width) {
      radix_pass<kWarps, kBits, Count>(kc, vc, kn, vn, n, lo, min(width, hi - lo), by_val, hist, s_scan);
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
    const int chunk = (((n + kWarps - 1) / kWarps) + 31) & ~31;
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
    for (int i = tid; i < n; i += kThreads) {
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
      for (int i = tid; i < n; i += kThreads) run_hi[run_lo[i]] = vc[i];
      __syncthreads();
      uint32_t* t = vc;  // ids now live in the spare value buffer; keys are unchanged (equal inside a run)
      vc = vn, vn = t;
    } else {
      sort_bits(diff_v, true);
      sort_bits(diff_k, false);
    }
  }

  // outputs in final order
  const uint64_t hi_key = (uint64_t)tile << 32;
  for (int i = tid; i < n; i += kThreads) {
    const size_t g = (size_t)range.x + i;
    a.keys_out[g] = hi_key | kc[i];
    a.vals_out[g] = vc[i];
  }
  if (!kGather) return;
  auto emit = [&](int i, uint32_t id, const float4& q0, const float4& q1, const float4& q2) {
    const size_t g = (size_t)range.x + i;
    float4* dst = reinterpret_cast<float4*>(a.inst + g);
    dst[0] = q0, dst[1] = q1, dst[2] = q2;
  };
  for (int i = tid; i < n; i += 2 * kThreads) {  // two independent record gathers in flight per thread
    const int i2 = i + kThreads;
    const uint32_t id = vc[i], id2 = i2 < n ? vc[i2] : id;
    const float4* src = reinterpret_cast<const float4*>(a.rec + id);
    const float4* src2 = reinterpret_cast<const float4*>(a.rec + id2);
    const float4 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2);
    const float4 r0 = __ldg(src2), r1 = __ldg(src2 + 1), r2 = __ldg(src2 + 2);
    emit(i, id, q0, q1, q2);
    if (i2 < n) emit(i2, id2, r0, r1, r2);
  }
}

// Record gather of the long lists (see sort_one_tile): work item = 1024 consecutive sorted instances of one of the
// first summary[2] tiles of the order whose length lies in (min_len, max_len]; the CTAs stride over the items.
constexpr int kGatherSlice = 1024, kGatherSlices = (kLargeCap + kGatherSlice - 1) / kGatherSlice;
__global__ void __launch_bounds__(256) gather_long_kernel(const TileSortArgs a) {
  const uint32_t work = a.summary[2] * kGatherSlices;
  for (uint32_t w = blockIdx.x; w < work; w += gridDim.x) {
    const uint32_t p = w / kGatherSlices, sl = w % kGatherSlices;
    const uint2 range = a.ranges[a.order[p]];
    const uint32_t n = range.y - range.x;
    if (n <= a.min_len || n > a.max_len || sl * kGatherSlice >= n) continue;
    uint32_t id[4];
    size_t g[4];
    bool on[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t i = sl * kGatherSlice + u * 256 + threadIdx.x;
      on[u] = i < n;
      g[u] = (size_t)range.x + (on[u] ? i : 0u);
      id[u] = a.vals_out[g[u]];
    }
    float4 q[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4* src = reinterpret_cast<const float4*>(a.rec + id[u]);
      q[u][0] = __ldg(src), q[u][1] = __ldg(src + 1), q[u][2] = __ldg(src + 2);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (on[u]) {
        float4* dst = reinterpret_cast<float4*>(a.inst + g[u]);
        dst[0] = q[u][0], dst[1] = q[u][1], dst[2] = q[u][2];
      }
    }
  }
}

// Short lists: one CTA per tile. Long lists: a few persistent CTAs (each takes a whole SM's shared memory) walk
// the first summary[2] entries of the launch order.
template <int kCap, int kThreads, bool kPersistent>
__global__ void __launch_bounds__(kThreads) tile_sort_kernel(const TileSortArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (kPersistent) {
    const uint32_t n_long = a.summary[2];
    for (uint32_t t = blockIdx.x; t < n_long; t += gridDim.x)
      sort_one_tile<kCap, kThreads, false>(a, a.order[t], smem_raw);
  } else {
    sort_one_tile<kCap, kThreads, true>(a, a.order[blockIdx.x], smem_raw);
  }
}

constexpr int kSmallThreads = 256, kLargeThreads = 1024;
constexpr size_t sort_smem(int cap, int threads) {  // keys/ids ping-pong + per-warp histograms (sort_one_tile)
  return (size_t)cap * 16 + (size_t)(threads / 32) * (threads >= 512 ? 512 * 2 : 256 * 4);
}

}  // namespace

cudaError_t launch_tile_scan(int tiles, const ImageState& img, uint32_t r_cap, cudaStream_t s) {
  tile_scan_kernel<<<1, 1024, 0, s>>>(tiles, img.tile_count, img.ranges, img.tile_order, img.tile_last, img.summary,
                                      (uint32_t)kSmallMax, r_cap, (uint32_t)kLargeCap);
  return cudaGetLastError();
}

cudaError_t launch_scatter_instances(int P, const GeomState& g, const int* radii, const ImageState& img,
                                     const BinningState& b, int grid_x, int grid_y, cudaStream_t s) {
  uint32_t* seg_keys = reinterpret_cast<uint32_t*>(b.keys_unsorted);
  scatter_instances_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.rec, radii, img.tile_count, seg_keys, b.vals_unsorted,
                                                          grid_x, grid_y, img.summary);
  return cudaGetLastError();
}

cudaError_t launch_tile_sort(int tiles, int n_large, const ImageState& img, const BinningState& b, const GeomState& g,
                             int grid_x, cudaStream_t s) {
  constexpr int kMidCap = 4096;  // a power of two: class boundaries coincide with tile_order's buckets
  constexpr size_t kSmemSmall = sort_smem(kSmallCap, kSmallThreads), kSmemLarge = sort_smem(kLargeCap, kLargeThreads);
  constexpr size_t kSmemMid = sort_smem(kMidCap, kLargeThreads);
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tile_sort_kernel<kSmallCap, kSmallThreads, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemSmall);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tile_sort_kernel<kLargeCap, kLargeThreads, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLarge);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tile_sort_kernel<kMidCap, kLargeThreads, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMid);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  static int sm_count[64] = {};
  int sms = 148;
  if (dev >= 0 && dev < 64) {
    if (!sm_count[dev] && cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      sm_count[dev] = 148;
    sms = sm_count[dev];
  }
  TileSortArgs a;
  a.ranges = img.ranges, a.order = img.tile_order, a.summary = img.summary;
  a.seg_keys = reinterpret_cast<const uint32_t*>(b.keys_unsorted), a.seg_vals = b.vals_unsorted;
  a.rec = g.rec, a.keys_out = b.keys, a.vals_out = b.vals, a.inst = b.inst, a.grid_x = grid_x;
  // Three independent classes, three streams: lists of 4096+ instances (one 1024-thread CTA with all of an SM's
  // shared memory each), lists of 2048..4095 (1024-thread CTAs, two per SM) and the many short ones (256-thread
  // CTAs). The longest lists come first in tile_order (bucketed by floor(log2(length))): both persistent kernels
  // walk its first summary[2] entries and skip the lists of the other class.
  struct Side {
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
  };
  static thread_local Side side[64];
  Side* sd = nullptr;
  if (n_large != 0 && dev >= 0 && dev < 64) {  // n_large < 0: not known on the host (sync-free forward)
    sd = &side[dev];
    if (!sd->stream[0]) {
      bool ok = cudaEventCreateWithFlags(&sd->fork, cudaEventDisableTiming) == cudaSuccess;
      for (int i = 0; i < 2 && ok; ++i)
        ok = cudaStreamCreateWithFlags(&sd->stream[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&sd->join[i], cudaEventDisableTiming) == cudaSuccess;
      if (!ok) {
        sd->stream[0] = nullptr;
        sd = nullptr;
        (void)cudaGetLastError();
      }
    }
  }
  cudaStream_t mid_stream = s, small_stream = s;
  if (sd) {
    cudaError_t e = cudaEventRecord(sd->fork, s);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(sd->stream[i], sd->fork, 0);
    if (e != cudaSuccess) return e;
    mid_stream = sd->stream[0], small_stream = sd->stream[1];
  }
  if (n_large != 0) {
    const int cap = n_large > 0 ? n_large : tiles;
    a.min_len = (uint32_t)kMidCap - 1u, a.max_len = (uint32_t)kLargeCap;
    const int gwork = cap * kGatherSlices, ggrid = gwork < 8 * sms ? gwork : 8 * sms;
    tile_sort_kernel<kLargeCap, kLargeThreads, true><<<cap < sms ? cap : sms, kLargeThreads, kSmemLarge, s>>>(a);
    gather_long_kernel<<<ggrid, 256, 0, s>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    a.min_len = (uint32_t)kSmallMax, a.max_len = (uint32_t)kMidCap - 1u;
    tile_sort_kernel<kMidCap, kLargeThreads, true><<<cap < 2 * sms ? cap : 2 * sms, kLargeThreads, kSmemMid, mid_stream>>>(a);
    gather_long_kernel<<<ggrid, 256, 0, mid_stream>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  a.min_len = 0, a.max_len = (uint32_t)kSmallMax;
  tile_sort_kernel<kSmallCap, kSmallThreads, false><<<tiles, kSmallThreads, kSmemSmall, small_stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (sd) {
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
      e = cudaEventRecord(sd->join[i], sd->stream[i]);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(s, sd->join[i], 0);
    }
  }
  return e;
}

}  // namespace gsr
