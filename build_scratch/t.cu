__global__ void k(float* g, const float* v, const int* idx) {
  __shared__ float s[256];
  s[threadIdx.x] = 0.f;
  __syncthreads();
  atomicAdd(&s[idx[threadIdx.x]], v[threadIdx.x]);
  __syncthreads();
  g[threadIdx.x] = s[threadIdx.x];
}
