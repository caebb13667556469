// The two per-iteration GPU steps that sit right behind the rasterizer in GS-LIVM's mapping loop,
// fused: the image loss producing dL/dC directly in the [3,H,W] layout gsr_backward reads, and Adam over flat
// parameter / gradient buffers. sm_100a.
//
//  gsr_image_loss  <- gaussian_splatting::l1_loss + ssim (include/gs/gs/loss_utils.cuh:11-70) and their combination
//                     (1 - lambda) L1 + lambda (1 - SSIM) (src/liw/lioOptimization.cpp:1729-1734), forward AND backward.
//                     The reference runs five grouped 11x11 conv2d (plus their autograd backwards) per view; here the
//                     window is applied separably in shared memory, two kernels per view.
//                     The window is the reference's own, NON-standard one (loss_utils.cuh:24-30):
//                     g[x] = exp(-floor((x - 11) / 2)^2 / (2 * 1.5^2)), x = 0..10, normalised -- not symmetric.
//  gsr_adam_step   <- torch::optim::Adam::step as GaussianModel::Training_setup configures it
//                     (src/gs/gaussian.cu:396-428: six groups, eps 1e-15, betas 0.9 / 0.999, no weight decay, no amsgrad)
//                     over ONE flat buffer with a per-group learning rate, one launch.
#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kWin = 11, kPad = kWin / 2;
constexpr int kLT = 16;                // pixels per CTA edge
constexpr int kHalo = kLT + 2 * kPad;  // 26
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

struct Window {
  float w[kWin];
};

// loss_utils.cuh:24-30, evaluated in float like the reference does (std::pow / std::floor / std::exp on floats)
Window make_window() {
  Window win;
  float sum = 0.f;
  for (int x = 0; x < kWin; ++x) {
    const float f = floorf(static_cast<float>(x - kWin) / 2.f);
    win.w[x] = expf(-(f * f) / (2.f * 1.5f * 1.5f));
    sum += win.w[x];
  }
  for (int x = 0; x < kWin; ++x) win.w[x] /= sum;
  return win;
}

struct LossArgs {
  const float* image;   // [3,H,W]
  const float* target;  // [3,H,W]
  float* maps;          // [3][3,H,W]: dm/dmu1, dm/ds1, dm/ds12 of the SSIM map
  float* dL_dimage;     // [3,H,W]
  float* sums;          // [2]: sum |x - y|, sum ssim_map
  int W, H;
  float lambda;
  Window win;
};

__device__ __forceinline__ float block_sum(float v, float* s_red) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (kLT * kLT) / 32 ? s_red[threadIdx.x] : 0.f;
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  }
  __syncthreads();
  return t;  // valid in thread 0
}

// Pass 1: per pixel and channel the five windowed moments -> SSIM map value and its three partial derivatives;
// per CTA the partial sums of |x - y| and of the map.
__global__ void __launch_bounds__(kLT* kLT) ssim_stats_kernel(const LossArgs a) {
  __shared__ float s_x[kHalo][kHalo + 1], s_y[kHalo][kHalo + 1];
  __shared__ float s_h[5][kHalo][kLT + 1];
  __shared__ float s_red[8];
  const int tx = threadIdx.x % kLT, ty = threadIdx.x / kLT;
  const int x0 = blockIdx.x * kLT, y0 = blockIdx.y * kLT, ch = blockIdx.z;
  const size_t plane = (size_t)a.W * a.H;
  const float* img = a.image + ch * plane;
  const float* tgt = a.target + ch * plane;
  for (int i = threadIdx.x; i < kHalo * kHalo; i += kLT * kLT) {  // zero padding, as conv2d(padding = 5)
    const int ly = i / kHalo, lx = i % kHalo;
    const int gx = x0 + lx - kPad, gy = y0 + ly - kPad;
    const bool in = gx >= 0 && gx < a.W && gy >= 0 && gy < a.H;
    s_x[ly][lx] = in ? img[(size_t)gy * a.W + gx] : 0.f;
    s_y[ly][lx] = in ? tgt[(size_t)gy * a.W + gx] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kHalo * kLT; i += kLT * kLT) {  // horizontal
    const int ly = i / kLT, lx = i % kLT;
    float mx = 0.f, my = 0.f, mxx = 0.f, myy = 0.f, mxy = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const float w = a.win.w[k], x = s_x[ly][lx + k], y = s_y[ly][lx + k];
      mx = fmaf(w, x, mx), my = fmaf(w, y, my);
      mxx = fmaf(w, x * x, mxx), myy = fmaf(w, y * y, myy), mxy = fmaf(w, x * y, mxy);
    }
    s_h[0][ly][lx] = mx, s_h[1][ly][lx] = my, s_h[2][ly][lx] = mxx, s_h[3][ly][lx] = myy, s_h[4][ly][lx] = mxy;
  }
  __syncthreads();
  float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
  for (int k = 0; k < kWin; ++k) {  // vertical
    const float w = a.win.w[k];
    mu1 = fmaf(w, s_h[0][ty + k][tx], mu1), mu2 = fmaf(w, s_h[1][ty + k][tx], mu2);
    e11 = fmaf(w, s_h[2][ty + k][tx], e11), e22 = fmaf(w, s_h[3][ty + k][tx], e22);
    e12 = fmaf(w, s_h[4][ty + k][tx], e12);
  }
  const int gx = x0 + tx, gy = y0 + ty;
  const bool in = gx < a.W && gy < a.H;
  const float sig1 = e11 - mu1 * mu1, sig2 = e22 - mu2 * mu2, sig12 = e12 - mu1 * mu2;
  const float A1 = 2.f * mu1 * mu2 + kC1, A2 = 2.f * sig12 + kC2;
  const float B1 = mu1 * mu1 + mu2 * mu2 + kC1, B2 = sig1 + sig2 + kC2;
  const float inv = 1.f / (B1 * B2);
  const float m = A1 * A2 * inv;
  float l1 = 0.f, ms = 0.f;
  if (in) {
    const size_t p = (size_t)gy * a.W + gx;
    // total derivatives w.r.t. the windowed moments of x: mu1, E[x^2], E[xy]
    a.maps[(0 * 3 + ch) * plane + p] = 2.f * mu2 * (A2 - A1) * inv - 2.f * mu1 * m * (B2 - B1) * inv;
    a.maps[(1 * 3 + ch) * plane + p] = -m / B2;
    a.maps[(2 * 3 + ch) * plane + p] = 2.f * A1 * inv;
    l1 = fabsf(s_x[ty + kPad][tx + kPad] - s_y[ty + kPad][tx + kPad]);
    ms = m;
  }
  const float t1 = block_sum(l1, s_red);
  const float t2 = block_sum(ms, s_red);
  if (threadIdx.x == 0) {
    atomicAdd(a.sums + 0, t1);
    atomicAdd(a.sums + 1, t2);
  }
}

// Pass 2: the three derivative maps through the transposed window, combined with the L1 term into dL/dimage.
__global__ void __launch_bounds__(kLT* kLT) ssim_grad_kernel(const LossArgs a) {
  __shared__ float s_m[3][kHalo][kHalo + 1];
  __shared__ float s_h[3][kHalo][kLT + 1];
  const int tx = threadIdx.x % kLT, ty = threadIdx.x / kLT;
  const int x0 = blockIdx.x * kLT, y0 = blockIdx.y * kLT, ch = blockIdx.z;
  const size_t plane = (size_t)a.W * a.H;
  for (int i = threadIdx.x; i < kHalo * kHalo; i += kLT * kLT) {
    const int ly = i / kHalo, lx = i % kHalo;
    const int gx = x0 + lx - kPad, gy = y0 + ly - kPad;
    const bool in = gx >= 0 && gx < a.W && gy >= 0 && gy < a.H;
    const size_t p = (size_t)gy * a.W + gx;
#pragma unroll
    for (int q = 0; q < 3; ++q) s_m[q][ly][lx] = in ? a.maps[(q * 3 + ch) * plane + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kHalo * kLT; i += kLT * kLT) {  // horizontal, window flipped
    const int ly = i / kLT, lx = i % kLT;
    float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const float w = a.win.w[kWin - 1 - k];
      h0 = fmaf(w, s_m[0][ly][lx + k], h0), h1 = fmaf(w, s_m[1][ly][lx + k], h1), h2 = fmaf(w, s_m[2][ly][lx + k], h2);
    }
    s_h[0][ly][lx] = h0, s_h[1][ly][lx] = h1, s_h[2][ly][lx] = h2;
  }
  __syncthreads();
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
  for (int k = 0; k < kWin; ++k) {  // vertical, window flipped
    const float w = a.win.w[kWin - 1 - k];
    g0 = fmaf(w, s_h[0][ty + k][tx], g0), g1 = fmaf(w, s_h[1][ty + k][tx], g1), g2 = fmaf(w, s_h[2][ty + k][tx], g2);
  }
  const int gx = x0 + tx, gy = y0 + ty;
  if (gx < a.W && gy < a.H) {
    const size_t p = (size_t)gy * a.W + gx;
    const float x = a.image[ch * plane + p], y = a.target[ch * plane + p];
    const float inv_n = 1.f / (3.f * (float)plane);
    const float d = x - y;
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    a.dL_dimage[ch * plane + p] =
        (1.f - a.lambda) * inv_n * sgn - a.lambda * inv_n * (g0 + 2.f * x * g1 + y * g2);
  }
}

// lambda == 0: plain L1, one pass
__global__ void __launch_bounds__(256) l1_loss_kernel(const float* __restrict__ image, const float* __restrict__ target,
                                                      float* __restrict__ dL_dimage, float* __restrict__ sums, size_t n,
                                                      float inv_n) {
  __shared__ float s_red[8];
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = image[i] - target[i];
    acc += fabsf(d);
    dL_dimage[i] = d > 0.f ? inv_n : (d < 0.f ? -inv_n : 0.f);
  }
  const float t = block_sum(acc, s_red);
  if (threadIdx.x == 0) atomicAdd(sums, t);
}

__global__ void loss_finish_kernel(const float* __restrict__ sums, float* __restrict__ loss, float lambda, float inv_n) {
  // (1 - lambda) * mean|x - y| + lambda * (1 - mean(ssim_map))   (lioOptimization.cpp:1734)
  const float l1 = sums[0] * inv_n, ss = sums[1] * inv_n;
  loss[0] = (1.f - lambda) * l1 + lambda * (1.f - ss);
  loss[1] = l1;
  loss[2] = ss;
}

// ---- Adam -------------------------------------------------------------------------------------------------------
struct AdamGroups {
  unsigned long long begin[8], end[8];  // element ranges of the flat buffers
  float lr[8], lr_other[8];
  unsigned period[8], lo[8], hi[8];     // period 0: lr everywhere; else lr on [lo, hi) of every row, lr_other elsewhere
  int n;
};

__device__ __forceinline__ void adam_element(float& p, float grad, float& m, float& v, float step_size, float beta1,
                                             float beta2, float eps, float sqrt_bc2) {
  m = m * beta1 + grad * (1.f - beta1);           // exp_avg.mul_(beta1).add_(grad, 1 - beta1)
  v = v * beta2 + (grad * grad) * (1.f - beta2);  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / sqrt_bc2 + eps;  // (exp_avg_sq.sqrt() / sqrt(bc2)).add_(eps)
  p = p - step_size * (m / denom);                // param.addcdiv_(exp_avg, denom, -step_size)
}

// 16-byte accesses where the group's range allows it (7 streams of 4 bytes per element: the step is HBM-bound), scalar
// head / tail otherwise.
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        const AdamGroups groups, float beta1, float beta2, float eps,
                                                        float bc1, float sqrt_bc2) {
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (int k = 0; k < groups.n; ++k) {
    // torch/csrc/api/src/optim/adam.cpp: step_size = lr / bias_correction1
    const float step_a = groups.lr[k] / bc1, step_b = groups.lr_other[k] / bc1;
    const unsigned period = groups.period[k], lo = groups.lo[k], hi = groups.hi[k];
    const size_t begin = groups.begin[k], end = groups.end[k];
    auto step_of = [&](size_t i) {
      if (period == 0) return step_a;
      const unsigned ph = (unsigned)((i - begin) % period);
      return (ph >= lo && ph < hi) ? step_a : step_b;
    };
    const size_t vb = (begin + 3) & ~(size_t)3, ve = end & ~(size_t)3;  // the 16-byte aligned middle
    if (vb < ve) {
      for (size_t i = vb + 4 * t0; i < ve; i += 4 * stride) {
        float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i),
               vv = *reinterpret_cast<float4*>(v + i);
        const float4 gg = *reinterpret_cast<const float4*>(g + i);
        adam_element(pp.x, gg.x, mm.x, vv.x, step_of(i), beta1, beta2, eps, sqrt_bc2);
        adam_element(pp.y, gg.y, mm.y, vv.y, step_of(i + 1), beta1, beta2, eps, sqrt_bc2);
        adam_element(pp.z, gg.z, mm.z, vv.z, step_of(i + 2), beta1, beta2, eps, sqrt_bc2);
        adam_element(pp.w, gg.w, mm.w, vv.w, step_of(i + 3), beta1, beta2, eps, sqrt_bc2);
        *reinterpret_cast<float4*>(m + i) = mm, *reinterpret_cast<float4*>(v + i) = vv;
        *reinterpret_cast<float4*>(p + i) = pp;
      }
    }
    // head [begin, vb) and tail [ve, end): at most three elements each (or the whole of a tiny range)
    const size_t head_end = vb < end ? vb : end, tail_begin = ve > head_end ? ve : head_end;
    for (size_t i = begin + t0; i < head_end; i += stride) adam_element(p[i], g[i], m[i], v[i], step_of(i), beta1, beta2, eps, sqrt_bc2);
    for (size_t i = tail_begin + t0; i < end; i += stride) adam_element(p[i], g[i], m[i], v[i], step_of(i), beta1, beta2, eps, sqrt_bc2);
  }
}

}  // namespace
}  // namespace gsr

using namespace gsr;

extern "C" {

size_t gsr_image_loss_scratch_floats(int width, int height) {
  if (width <= 0 || height <= 0) return 0;
  return (size_t)9 * width * height + 8;  // three derivative maps [3,H,W] + the two partial sums (16-byte aligned)
}

int gsr_image_loss(const float* image, const float* target, int width, int height, float lambda_dssim,
                   float* dL_dimage, float* loss_out, float* scratch, void* stream_v) {
  if (!image || !target || !dL_dimage || !loss_out || !scratch || width <= 0 || height <= 0 || lambda_dssim < 0.f ||
      lambda_dssim > 1.f)
    return GSR_ERR_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream_v);
  const size_t plane = (size_t)width * height;
  float* sums = scratch + 9 * plane;
  if (cudaMemsetAsync(sums, 0, 2 * sizeof(float), s) != cudaSuccess) return GSR_ERR_CUDA;
  const float inv_n = 1.f / (3.f * (float)plane);
  if (lambda_dssim == 0.f) {
    const int grid = device_sm_count() * 4;
    l1_loss_kernel<<<grid, 256, 0, s>>>(image, target, dL_dimage, sums, 3 * plane, inv_n);
  } else {
    static const Window win = make_window();
    LossArgs a;
    a.image = image, a.target = target, a.maps = scratch, a.dL_dimage = dL_dimage, a.sums = sums;
    a.W = width, a.H = height, a.lambda = lambda_dssim, a.win = win;
    const dim3 grid((width + kLT - 1) / kLT, (height + kLT - 1) / kLT, 3);
    ssim_stats_kernel<<<grid, kLT * kLT, 0, s>>>(a);
    ssim_grad_kernel<<<grid, kLT * kLT, 0, s>>>(a);
  }
  loss_finish_kernel<<<1, 1, 0, s>>>(sums, loss_out, lambda_dssim, inv_n);
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

int gsr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const gsr_adam_group* groups,
                  int n_groups, float beta1, float beta2, float eps, int step, void* stream_v) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !groups || n_groups < 1 || n_groups > 8 || step < 1)
    return GSR_ERR_INVALID;
  AdamGroups g;
  g.n = n_groups;
  size_t total = 0;
  for (int k = 0; k < n_groups; ++k) {
    g.begin[k] = groups[k].first, g.end[k] = groups[k].first + groups[k].count, g.lr[k] = groups[k].lr;
    g.period[k] = groups[k].period, g.lo[k] = groups[k].phase_lo, g.hi[k] = groups[k].phase_hi;
    g.lr_other[k] = groups[k].lr_other;
    if (g.period[k] && (g.lo[k] > g.hi[k] || g.hi[k] > g.period[k])) return GSR_ERR_INVALID;
    total += groups[k].count;
  }
  if (total == 0) return GSR_OK;
  // bias corrections in double on the host like LibTorch (std::pow on doubles), applied in float on the device
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  size_t blocks = (total / n_groups / 4 + 255) / 256;
  const size_t cap = (size_t)device_sm_count() * 8;
  blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
  adam_step_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
      params, grads, exp_avg, exp_avg_sq, g, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2));
  return cudaGetLastError() == cudaSuccess ? GSR_OK : GSR_ERR_CUDA;
}

}  // extern "C"
