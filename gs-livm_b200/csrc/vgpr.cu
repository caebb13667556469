// Voxel-GP regression (VGPR) + variance-centred Gaussian initialisation on the GPU, batched properly (SURVEY.md
// 8f-f4). sm_100a.
//
// The producer side of GS-LIVM's map: for every voxel that holds enough LiDAR returns, a Gaussian process regresses
// the surface height over a num_gp_side x num_gp_side lattice from the voxel's training points, and every
// neighbour x neighbour block of that lattice becomes one 3D Gaussian whose mean and covariance are the
// inverse-variance weighted moments of the block's lattice points. Reference: src/gp3d/gpprocess.cu:8-60 (lattice,
// kernel k = exp(-kernel_size |d|) + sigma^2 on the diagonal, mean-centred targets), :602-668 (Ky^-1, Ky^-1 K*, f*,
// K*^T Ky^-1 K*), :420-458 (fastInitial3DGS), src/gs/gaussian.cu:325-386 (scale = sqrt(diag(cov) * scale_factor)).
// The reference runs this per frame with per-call cudaMalloc / cublasCreate, a round trip to the host and tensors built
// point by point (gpprocess.cu:523-559,772-863); here it is ONE launch, one CTA per voxel, everything in shared memory
// and registers:
//   1. the T training points (T <= 32) and the T x T kernel matrix Ky; its Cholesky factor by one warp;
//   2. thread n = lattice point n: k*(n) (T values), x = Ky^-1 k*(n) by two triangular solves in registers,
//      f*(n) = tf . x + mean, variance entry kvar(n) = k*(n) . x;
//   3. thread b < (side / neighbour)^2 = Gaussian b: weighted mean, covariance, scales of its block.
// No cuBLAS, no per-voxel allocation: ~0.45 MFLOP and < 6 KB of shared memory per voxel.
#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kMaxTrain = 32, kMaxSide = 16;

struct VgprArgs {
  int V, T, side, neighbour, full_cover;
  float grid, kernel_size, variance_sensor, scale_factor;
  const float* train;     // [V,T,3]
  const int* voxel;       // [V,3] integer voxel coordinates: the voxel spans [voxel * grid, (voxel + 1) * grid)
  const int* direction;   // [V] regressed axis 0 / 1 / 2 (cell.cpp:5-30)
  float* means;           // [V * g * g, 3]
  float* covs;            // [V * g * g, 3, 3]
  float* scales;          // [V * g * g, 3]
  float* var_mean;        // [V] 1 - mean(kvar): the reference's "needs more points" gate (gpprocess.cu:63-122)
  float* test_points;     // optional [V, side^2, 3]
  float* kvar;            // optional [V, side^2]
};

__global__ void __launch_bounds__(256) vgpr_kernel(const VgprArgs a) {
  __shared__ float s_tx[kMaxTrain], s_ty[kMaxTrain], s_tf[kMaxTrain];
  __shared__ float s_L[kMaxTrain * kMaxTrain];  // Ky, then its lower Cholesky factor (row-major, stride T)
  __shared__ float s_p[kMaxSide * kMaxSide][3];
  __shared__ float s_w[kMaxSide * kMaxSide];
  __shared__ float s_mean, s_sum;
  const int v = blockIdx.x, tid = threadIdx.x, T = a.T, n2 = a.side * a.side;
  const int dir = a.direction[v];
  // direction -> (train x axis, train y axis, regressed axis): pointMatrix.h:105-125
  const int ax = dir == 0 ? 1 : (dir == 1 ? 2 : 0), ay = dir == 0 ? 2 : (dir == 1 ? 0 : 1), af = dir;
  if (tid < T) {
    const float* p = a.train + ((size_t)v * T + tid) * 3;
    s_tx[tid] = p[ax], s_ty[tid] = p[ay], s_tf[tid] = p[af];
  }
  __syncthreads();
  if (tid == 0) {
    float m = 0.f;
    for (int i = 0; i < T; ++i) m += s_tf[i];
    s_mean = m / (float)T;
    s_sum = 0.f;
  }
  for (int e = tid; e < T * T; e += blockDim.x) {
    const int i = e / T, j = e - i * T;
    const float dx = s_tx[j] - s_tx[i], dy = s_ty[j] - s_ty[i];
    s_L[e] = expf(-a.kernel_size * sqrtf(dx * dx + dy * dy)) + (i == j ? a.variance_sensor * a.variance_sensor : 0.f);
  }
  __syncthreads();
  if (tid < T) s_tf[tid] -= s_mean;
  // Cholesky of the T x T matrix by warp 0, column by column (Ky is symmetric positive definite: an exponential kernel
  // plus sigma^2 I)
  if (tid < 32) {
    for (int k = 0; k < T; ++k) {
      float d = s_L[k * T + k];
      for (int j = 0; j < k; ++j) d -= s_L[k * T + j] * s_L[k * T + j];
      d = sqrtf(fmaxf(d, 1e-20f));
      __syncwarp();
      if (tid == 0) s_L[k * T + k] = d;
      if (tid > k && tid < T) {
        float s = s_L[tid * T + k];
        for (int j = 0; j < k; ++j) s -= s_L[tid * T + j] * s_L[k * T + j];
        s_L[tid * T + k] = s / d;
      }
      __syncwarp();
    }
  }
  __syncthreads();

  const float interval = a.grid / (float)(a.full_cover ? a.side - 1 : a.side);
  const float off = a.full_cover ? 0.f : 0.5f;
  const float lo_x = (float)a.voxel[3 * v + ax] * a.grid, lo_y = (float)a.voxel[3 * v + ay] * a.grid;
  if (tid < n2) {
    const float gx = lo_x + interval * ((float)(tid / a.side) + off), gy = lo_y + interval * ((float)(tid % a.side) + off);
    float ks[kMaxTrain], x[kMaxTrain];
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      const float dx = gx - s_tx[t], dy = gy - s_ty[t];
      ks[t] = expf(-a.kernel_size * sqrtf(dx * dx + dy * dy));
    }
    // L y = k*, L^T x = y
#pragma unroll 1
    for (int i = 0; i < T; ++i) {
      float s = ks[i];
      for (int j = 0; j < i; ++j) s -= s_L[i * T + j] * x[j];
      x[i] = s / s_L[i * T + i];
    }
#pragma unroll 1
    for (int i = T - 1; i >= 0; --i) {
      float s = x[i];
      for (int j = i + 1; j < T; ++j) s -= s_L[j * T + i] * x[j];
      x[i] = s / s_L[i * T + i];
    }
    float f = 0.f, kv = 0.f;
    for (int t = 0; t < T; ++t) f += s_tf[t] * x[t], kv += ks[t] * x[t];
    f += s_mean;
    s_p[tid][ax] = gx, s_p[tid][ay] = gy, s_p[tid][af] = f;
    s_w[tid] = kv;
    atomicAdd(&s_sum, kv);
    if (a.test_points) {
      float* o = a.test_points + ((size_t)v * n2 + tid) * 3;
      o[ax] = gx, o[ay] = gy, o[af] = f;
    }
    if (a.kvar) a.kvar[(size_t)v * n2 + tid] = kv;
  }
  __syncthreads();
  if (tid == 0) a.var_mean[v] = 1.f - s_sum / (float)n2;

  // variance-centred initialisation: Gaussian b = block (bi, bj) of neighbour x neighbour lattice points
  const int g = a.side / a.neighbour;
  if (tid < g * g) {
    const int bi = tid / g, bj = tid - bi * g;
    float wsum = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f;
    for (int u = 0; u < a.neighbour; ++u)
      for (int w = 0; w < a.neighbour; ++w) {
        const int n = (bi * a.neighbour + u) * a.side + bj * a.neighbour + w;
        const float wt = 1.f / s_w[n];
        wsum += wt, m0 += wt * s_p[n][0], m1 += wt * s_p[n][1], m2 += wt * s_p[n][2];
      }
    m0 /= wsum, m1 /= wsum, m2 /= wsum;
    float c00 = 0.f, c01 = 0.f, c02 = 0.f, c11 = 0.f, c12 = 0.f, c22 = 0.f;
    for (int u = 0; u < a.neighbour; ++u)
      for (int w = 0; w < a.neighbour; ++w) {
        const int n = (bi * a.neighbour + u) * a.side + bj * a.neighbour + w;
        const float wt = 1.f / s_w[n];
        const float d0 = s_p[n][0] - m0, d1 = s_p[n][1] - m1, d2 = s_p[n][2] - m2;
        c00 += wt * d0 * d0, c01 += wt * d0 * d1, c02 += wt * d0 * d2, c11 += wt * d1 * d1, c12 += wt * d1 * d2,
            c22 += wt * d2 * d2;
      }
    c00 /= wsum, c01 /= wsum, c02 /= wsum, c11 /= wsum, c12 /= wsum, c22 /= wsum;
    const size_t o = (size_t)v * g * g + tid;
    a.means[3 * o] = m0, a.means[3 * o + 1] = m1, a.means[3 * o + 2] = m2;
    float* c = a.covs + 9 * o;
    c[0] = c00, c[1] = c01, c[2] = c02, c[3] = c01, c[4] = c11, c[5] = c12, c[6] = c02, c[7] = c12, c[8] = c22;
    a.scales[3 * o] = sqrtf(fmaxf(c00, 0.f) * a.scale_factor);
    a.scales[3 * o + 1] = sqrtf(fmaxf(c11, 0.f) * a.scale_factor);
    a.scales[3 * o + 2] = sqrtf(fmaxf(c22, 0.f) * a.scale_factor);
  }
}

}  // namespace
}  // namespace gsr

using namespace gsr;

extern "C" int gsr_vgpr_gaussians(int V, int T, const float* train, const int* voxel, const int* direction,
                                  const gsr_vgpr_params* prm, float* means, float* covs, float* scales, float* var_mean,
                                  float* test_points, float* kvar, void* stream) {
  if (V < 0 || !prm || T < 1 || T > kMaxTrain || prm->num_gp_side < 1 || prm->num_gp_side > kMaxSide || prm->neighbour < 1 ||
      prm->num_gp_side % prm->neighbour != 0 || !(prm->grid > 0.f))
    return GSR_ERR_INVALID;
  if (V == 0) return GSR_OK;
  if (!train || !voxel || !direction || !means || !covs || !scales || !var_mean) return GSR_ERR_INVALID;
  VgprArgs a;
  a.V = V, a.T = T, a.side = prm->num_gp_side, a.neighbour = prm->neighbour, a.full_cover = prm->full_cover;
  a.grid = prm->grid, a.kernel_size = prm->kernel_size, a.variance_sensor = prm->variance_sensor;
  a.scale_factor = prm->scale_factor;
  a.train = train, a.voxel = voxel, a.direction = direction;
  a.means = means, a.covs = covs, a.scales = scales, a.var_mean = var_mean, a.test_points = test_points, a.kvar = kvar;
  vgpr_kernel<<<V, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}
