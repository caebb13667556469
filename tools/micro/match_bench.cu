// Micro-benchmark: throughput of __match_any_sync against the ballot-per-bit formulation and shared-memory atomics.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o match_bench match_bench.cu && ./match_bench
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(const unsigned* in, unsigned* out, int iters) {
  __shared__ unsigned hist[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) hist[0][i] = 0;
  __syncthreads();
  unsigned x = in[blockIdx.x * blockDim.x + threadIdx.x], acc = 0;
  for (int it = 0; it < iters; ++it) {
    const unsigned d = (x >> (it & 7)) & 255u;
    unsigned peers;
    if (MODE == 0) {
      peers = __match_any_sync(0xffffffffu, d);
    } else if (MODE == 1) {
      peers = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const unsigned v = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        peers &= ((d >> b) & 1u) ? v : ~v;
      }
    } else {
      peers = atomicAdd(&hist[warp][d], 1u);
    }
    acc += __popc(peers & ((1u << lane) - 1u)) + (MODE == 2 ? peers : 0u);
    x = x * 1664525u + 1013904223u + acc;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  const int blocks = 148 * 8, threads = 256, iters = 2000;
  unsigned *in, *out;
  cudaMalloc(&in, blocks * threads * 4), cudaMalloc(&out, blocks * threads * 4);
  cudaMemset(in, 0x5a, blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const char* names[3] = {"match_any", "8 ballots", "smem atomicAdd"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<blocks, threads>>>(in, out, iters);
      else if (mode == 1) k<1><<<blocks, threads>>>(in, out, iters);
      else k<2><<<blocks, threads>>>(in, out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep) {
        const double warp_ops = (double)blocks * threads / 32 * iters;
        printf("%-16s %8.3f ms  %.2f cycles per warp-op per SM (at 1.9 GHz, 148 SMs)\n", names[mode], ms,
               ms * 1e-3 * 1.9e9 * 148 / warp_ops);
      }
    }
  }
  return 0;
}
