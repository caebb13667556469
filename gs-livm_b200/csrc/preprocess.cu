// Per-Gaussian forward preprocess + markVisible for sm_100a.
//
// Replaces FORWARD::preprocess / preprocessCUDA<3> (reference src/cuda_rasterizer/forward.cu:179-286,
// 443-505) and checkFrustum (rasterizer_impl.cu:52-60).
//
// Numerical contract. `depths`, `means2D`, `radii`, `tiles_touched` feed the sort keys and the tile
// ranges, which must be bit-identical to the reference. The reference is compiled by nvcc with FMA
// contraction, so "the same formula" is not enough: the operation DAG below (which product is a plain
// FMUL, which is fused into an FFMA, where the FADDs are) was read off the reference's sm_100a SASS
//  and is written with explicit round-to-nearest intrinsics so that this
// compiler cannot re-contract it. Divisions / square roots are IEEE (__frcp_rn, __fdiv_rn, __fsqrt_rn),
// ndc2Pix is evaluated in FP64 exactly like the reference's double literals force it to be
// (include/gs/cuda_rasterizer/auxiliary.h:35-37).
#include "gsr_common.cuh"

namespace gsr {
namespace {

// SH basis constants (reference auxiliary.h:22-33).
__device__ constexpr float kC0 = 0.28209479177387814f;
__device__ constexpr float kC1 = 0.4886025119029199f;
__device__ constexpr float kC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                     -1.0925484305920792f, 0.5462742152960396f};
__device__ constexpr float kC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                     0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                     -0.5900435899266435f};

struct PreArgs {
  int P, D, M, W, H, grid_x, grid_y;
  float scale_modifier, tan_fovx, tan_fovy, focal_x, focal_y;
  const float* means3D;
  const float* scales;
  const float* rotations;
  const float* opacities;
  const float* shs;
  const float* cov3D_precomp;
  const float* colors_precomp;
  const float* viewmatrix;
  const float* projmatrix;
  const float* cam_pos;
  int* radii;
  SplatRecord* rec;
  uint32_t* tiles_touched;
  uint8_t* clamped;
  uint32_t* tile_count;  // [tiles] instances per tile, accumulated here (tile_sort.cu)
  int vec_ok;  // means3D/scales/shs 16-byte aligned -> float4 staging
  int act;     // GSR_ACT_* bits: opacities / scales / rotations are raw parameters, activated on load
};

// a*b + c*d + e*f as nvcc contracts the reference's left-to-right sums:
// FMUL(c*d) -> FFMA(a,b,.) -> FFMA(e,f,.)
__device__ __forceinline__ float dot3_ref(float a, float b, float c, float d, float e, float f) {
  return __fmaf_rn(e, f, __fmaf_rn(a, b, __fmul_rn(c, d)));
}

// row r of a column-major 4x4 applied to (x,y,z,1): transformPoint4x3/4x4 (auxiliary.h:48-64).
__device__ __forceinline__ float xform_row(const float* m, int r, float x, float y, float z) {
  return __fadd_rn(dot3_ref(x, m[r], y, m[4 + r], z, m[8 + r]), m[12 + r]);
}

// Stage `count` floats starting at g (block-contiguous slab) into shared memory, float4 when aligned.
__device__ __forceinline__ void stage_floats(float* s, const float* g, int count, bool vec_ok) {
  const int tid = threadIdx.x;
  if (vec_ok) {
    const int nvec = count >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* s4 = reinterpret_cast<float4*>(s);
    for (int i = tid; i < nvec; i += blockDim.x) s4[i] = __ldg(g4 + i);
    for (int i = (nvec << 2) + tid; i < count; i += blockDim.x) s[i] = __ldg(g + i);
  } else {
    for (int i = tid; i < count; i += blockDim.x) s[i] = __ldg(g + i);
  }
}

// SH -> RGB for one Gaussian (reference computeColorFromSH, forward.cu:29-76). `sh` points at this
// Gaussian's coefficients, channel-interleaved [M][3]. Not on the bit-exact path (<= 1e-5 on C).
template <typename Load>
__device__ __forceinline__ void eval_sh(int deg, float3 dir, Load sh, float out[3]) {
  const float x = dir.x, y = dir.y, z = dir.z;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float r = kC0 * sh(0, c);
    if (deg > 0) {
      r = r - kC1 * y * sh(1, c) + kC1 * z * sh(2, c) - kC1 * x * sh(3, c);
      if (deg > 1) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        r = r + kC2[0] * xy * sh(4, c) + kC2[1] * yz * sh(5, c) + kC2[2] * (2.0f * zz - xx - yy) * sh(6, c) +
            kC2[3] * xz * sh(7, c) + kC2[4] * (xx - yy) * sh(8, c);
        if (deg > 2) {
          r = r + kC3[0] * y * (3.0f * xx - yy) * sh(9, c) + kC3[1] * xy * z * sh(10, c) +
              kC3[2] * y * (4.0f * zz - xx - yy) * sh(11, c) +
              kC3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh(12, c) +
              kC3[4] * x * (4.0f * zz - xx - yy) * sh(13, c) + kC3[5] * z * (xx - yy) * sh(14, c) +
              kC3[6] * x * (xx - 3.0f * yy) * sh(15, c);
        }
      }
    }
    out[c] = r + 0.5f;
  }
}

constexpr int kBlock = 64;  // Gaussians per CTA: many small CTAs per SM interleave staging, math and SH phases better than three big ones
constexpr int kShRow = 13;  // float4 per staged SH row: 12 used + 1 pad, so lane-strided LDS.128 are conflict-free
constexpr uint32_t kTileSlotBits = 8, kTileSlots = 1u << kTileSlotBits, kTileProbes = 8;

// kVec16: M == 16 and shs 16-byte aligned. The 192-byte SH rows of the block's SURVIVING Gaussians are staged
// into shared memory with coalesced 16-byte cp.async (LDGSTS) transfers and read back row-wise; rows of culled
// Gaussians (behind the camera, oversized, off-screen) are never fetched.
template <bool kVec16>
__global__ void __launch_bounds__(kBlock) preprocess_fwd_kernel(const PreArgs a) {
  __shared__ float s_view[16], s_proj[16], s_cam[3];
  __shared__ __align__(16) float s_means[kBlock * 3];
  __shared__ __align__(16) float s_scales[kBlock * 3];
  __shared__ uint8_t s_survive[kBlock];
  __shared__ __align__(16) uint32_t s_tile_key[kTileSlots], s_tile_cnt[kTileSlots];  // per-CTA tile -> instance count (see below)
  extern __shared__ __align__(16) float4 s_sh[];  // [kBlock][kShRow] when kVec16

  const int tid = threadIdx.x;
  const uint32_t sh0 = kVec16 ? smem_u32(s_sh) : 0u;  // (once: the conversion is not hoisted out of the staging loops)
  static_assert(kTileSlots == 4 * kBlock, "one 16-byte store per thread and table");
  reinterpret_cast<uint4*>(s_tile_key)[tid] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  reinterpret_cast<uint4*>(s_tile_cnt)[tid] = make_uint4(0u, 0u, 0u, 0u);
  const int base = blockIdx.x * kBlock;
  const int n = min(kBlock, a.P - base);
  for (int i = tid; i < 35; i += kBlock) {  // the camera: 16 + 16 + 3 floats
    if (i < 16) s_view[i] = __ldg(a.viewmatrix + i);
    else if (i < 32) s_proj[i - 16] = __ldg(a.projmatrix + i - 16);
    else s_cam[i - 32] = __ldg(a.cam_pos + i - 32);
  }
  stage_floats(s_means, a.means3D + (size_t)base * 3, n * 3, a.vec_ok);
  if (a.scales) stage_floats(s_scales, a.scales + (size_t)base * 3, n * 3, a.vec_ok);
  const int idx = base + tid;
  // the two per-Gaussian inputs that are not staged are requested before the barrier as well, so their latency
  // overlaps the staging instead of being exposed in the middle of the covariance math
  float4 rot_in = make_float4(0.f, 0.f, 0.f, 0.f);
  float op_in = 0.f;
  if (tid < n) {
    if (!a.cov3D_precomp) rot_in = __ldg(reinterpret_cast<const float4*>(a.rotations) + idx);
    op_in = __ldg(a.opacities + idx);
  }
  __syncthreads();
  // GaussianModel's activations (gaussian.cuh:40-54) on load, written like LibTorch's kernels compute them
  if (a.act) {
    if ((a.act & GSR_ACT_SCALE_EXP) && a.scales && tid < n) {  // own three values only: no hazard
#pragma unroll
      for (int c = 0; c < 3; ++c) s_scales[3 * tid + c] = expf(s_scales[3 * tid + c]);
    }
    if (a.act & GSR_ACT_OPACITY_SIGMOID) op_in = sigmoid_ref(op_in);
    if (a.act & GSR_ACT_ROTATION_NORMALIZE) rot_in = normalize_quat(rot_in);
  }

  int out_radius = 0;
  int rect_x0 = 0, rect_y0 = 0, rect_w = 1;
  uint32_t out_tiles = 0;
  bool survive = false;
  float4 rq0 = make_float4(0.f, 0.f, 0.f, 0.f), rq1 = rq0;
  float rdepth = 0.f;
  const float px = tid < n ? s_means[3 * tid] : 0.f, py = tid < n ? s_means[3 * tid + 1] : 0.f,
              pz = tid < n ? s_means[3 * tid + 2] : 0.f;

  // SH rows of the Gaussians that will most likely survive are requested NOW (coalesced cp.async), so the 192-byte
  // rows arrive while the covariance math runs: in front of the near plane, not oversized, and projected within a
  // generous bound of the image. The test is only a prefetch heuristic: a survivor it missed is fetched after the
  // exact tests (second pass below), a culled Gaussian it let through costs one wasted row.
  bool prefetched = false;
  if (kVec16) {
    if (tid < n) {
      const float zc = xform_row(s_view, 2, px, py, pz);
      bool likely = zc > 0.2f;
      float smax = 0.3f;
      if (a.scales) {
        smax = fmaxf(fmaxf(s_scales[3 * tid], s_scales[3 * tid + 1]), s_scales[3 * tid + 2]) * a.scale_modifier;
        likely = likely && !(smax > 0.3f);
      }
      if (likely) {
        const float w = xform_row(s_proj, 3, px, py, pz) + 0.0000001f;
        const float cxp = 0.5f * ((xform_row(s_proj, 0, px, py, pz) / w + 1.0f) * (float)a.W - 1.0f);
        const float cyp = 0.5f * ((xform_row(s_proj, 1, px, py, pz) / w + 1.0f) * (float)a.H - 1.0f);
        const float rb = 3.0f * sqrtf(2.25f * smax * smax * fmaxf(a.focal_x, a.focal_y) * fmaxf(a.focal_x, a.focal_y) / (zc * zc) + 0.3f) + 17.0f;
        likely = (cxp + rb > 0.f) && (cxp - rb < (float)a.W) && (cyp + rb > 0.f) && (cyp - rb < (float)a.H);
      }
      prefetched = likely;
    }
    s_survive[tid] = prefetched;
    __syncthreads();
    const float4* g4 = reinterpret_cast<const float4*>(a.shs) + (size_t)base * 12;
    for (int q = tid, row = tid / 12, rem = tid % 12; q < n * 12; q += kBlock) {  // row = q / 12 without dividing
      if (s_survive[row]) {
        const uint32_t dst = sh0 + (uint32_t)(row * kShRow + rem) * 16u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g4 + q) : "memory");
      }
      { rem += kBlock % 12, row += kBlock / 12; if (rem >= 12) rem -= 12, ++row; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  if (tid < n) do {  // single-pass block; `break` == the reference's early `return` with radius/tiles left at 0
    const float* V = s_view;
    const float* PM = s_proj;

    // p_view.z (forward.cu:222); near cull at z <= 0.2 (forward.cu:223)
    const float depth = xform_row(V, 2, px, py, pz);
    if (depth <= 0.2f) break;

    // scale_abnormal (forward.cu:19-25,227): cull when any activated scale * modifier > 0.3
    float sx = 0.f, sy = 0.f, sz = 0.f;
    if (a.scales) {
      sx = __fmul_rn(s_scales[3 * tid], a.scale_modifier);
      sy = __fmul_rn(s_scales[3 * tid + 1], a.scale_modifier);
      sz = __fmul_rn(s_scales[3 * tid + 2], a.scale_modifier);
      if (sx > 0.3f || sy > 0.3f || sz > 0.3f) break;
    }

    // p_hom / p_w / p_proj (forward.cu:232-234)
    const float hom_x = xform_row(PM, 0, px, py, pz);
    const float hom_y = xform_row(PM, 1, px, py, pz);
    const float hom_w = __fadd_rn(xform_row(PM, 3, px, py, pz), 0.0000001f);
    const float p_w = __frcp_rn(hom_w);
    const float proj_x = __fmul_rn(hom_x, p_w);
    const float proj_y = __fmul_rn(hom_y, p_w);

    // 3D covariance (forward.cu:138-176): M = S*R with the quaternion used as given, Sigma = M^T M.
    float c0, c1, c2, c3, c4, c5;
    if (a.cov3D_precomp) {
      const float* cp = a.cov3D_precomp + (size_t)idx * 6;
      c0 = __ldg(cp), c1 = __ldg(cp + 1), c2 = __ldg(cp + 2), c3 = __ldg(cp + 3), c4 = __ldg(cp + 4), c5 = __ldg(cp + 5);
    } else {
      const float4 q = rot_in;
      const float r = q.x, x = q.y, y = q.z, z = q.w;
      const float xz = __fmul_rn(x, z), rx = __fmul_rn(r, x), rz = __fmul_rn(r, z);
      const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
      const float xz_p_ry = __fmaf_rn(r, y, xz);
      const float yz_p_rx = __fmaf_rn(y, z, rx);
      const float xz_m_ry = __fmaf_rn(-r, y, xz);
      const float xy_m_rz = __fmaf_rn(x, y, -rz);
      const float xy_p_rz = __fmaf_rn(x, y, rz);
      const float yz_m_rx = __fmaf_rn(y, z, -rx);
      const float xx_p_yy = __fmaf_rn(x, x, yy);
      const float yy_p_zz = __fadd_rn(yy, zz);
      const float xx_p_zz = __fmaf_rn(x, x, zz);
      const float R00 = __fadd_rn(-__fadd_rn(yy_p_zz, yy_p_zz), 1.0f);
      const float R11 = __fadd_rn(-__fadd_rn(xx_p_zz, xx_p_zz), 1.0f);
      const float R22 = __fadd_rn(-__fadd_rn(xx_p_yy, xx_p_yy), 1.0f);
      // rows of M: every entry is one rounded product s_i * R_ij
      const float A0 = __fmul_rn(sx, R00);
      const float A1 = __fmul_rn(sx, __fadd_rn(xy_p_rz, xy_p_rz));
      const float A2 = __fmul_rn(sx, __fadd_rn(xz_m_ry, xz_m_ry));
      const float B0 = __fmul_rn(sy, __fadd_rn(xy_m_rz, xy_m_rz));
      const float B1 = __fmul_rn(sy, R11);
      const float B2 = __fmul_rn(sy, __fadd_rn(yz_p_rx, yz_p_rx));
      const float C0 = __fmul_rn(sz, __fadd_rn(xz_p_ry, xz_p_ry));
      const float C1 = __fmul_rn(sz, __fadd_rn(yz_m_rx, yz_m_rx));
      const float C2 = __fmul_rn(sz, R22);
      c0 = dot3_ref(A0, A0, B0, B0, C0, C0);
      c1 = dot3_ref(A0, A1, B0, B1, C0, C1);
      c2 = dot3_ref(A0, A2, B0, B2, C0, C2);
      c3 = dot3_ref(A1, A1, B1, B1, C1, C1);
      c4 = dot3_ref(A1, A2, B1, B2, C1, C2);
      c5 = dot3_ref(A2, A2, B2, B2, C2, C2);
    }

    // 2D covariance, EWA (forward.cu:79-133)
    const float tz = depth;
    const float tx0 = xform_row(V, 0, px, py, pz);
    const float ty0 = xform_row(V, 1, px, py, pz);
    const float limx = __fmul_rn(a.tan_fovx, 1.3f);
    const float limy = __fmul_rn(a.tan_fovy, 1.3f);
    const float txc = fminf(fmaxf(__fdiv_rn(tx0, tz), -limx), limx);
    const float tyc = fminf(fmaxf(__fdiv_rn(ty0, tz), -limy), limy);
    const float J00 = __fdiv_rn(a.focal_x, tz);
    const float J11 = __fdiv_rn(a.focal_y, tz);
    const float tz2 = __fmul_rn(tz, tz);
    const float J02 = __fdiv_rn(__fmul_rn(__fmul_rn(tz, -txc), a.focal_x), tz2);
    const float J12 = __fdiv_rn(__fmul_rn(__fmul_rn(tz, -tyc), a.focal_y), tz2);
    // T = W*J : T0i = J02*W2i + W0i*J00 ; T1i = J12*W2i + W1i*J11  (W0 = V[0,4,8], W1 = V[1,5,9], W2 = V[2,6,10])
    const float T00 = __fmaf_rn(J02, V[2], __fmul_rn(V[0], J00));
    const float T01 = __fmaf_rn(J02, V[6], __fmul_rn(V[4], J00));
    const float T02 = __fmaf_rn(J02, V[10], __fmul_rn(V[8], J00));
    const float T10 = __fmaf_rn(J12, V[2], __fmul_rn(V[1], J11));
    const float T11 = __fmaf_rn(J12, V[6], __fmul_rn(V[5], J11));
    const float T12 = __fmaf_rn(J12, V[10], __fmul_rn(V[9], J11));
    const float u0 = dot3_ref(T00, c0, T01, c1, T02, c2);
    const float u1 = dot3_ref(T00, c1, T01, c3, T02, c4);
    const float u2 = dot3_ref(T00, c2, T01, c4, T02, c5);
    const float v0 = dot3_ref(T10, c0, T11, c1, T12, c2);
    const float v1 = dot3_ref(T10, c1, T11, c3, T12, c4);
    const float v2 = dot3_ref(T10, c2, T11, c4, T12, c5);
    const float cov_x = __fadd_rn(dot3_ref(T00, u0, T01, u1, T02, u2), 0.3f);
    const float cov_y = dot3_ref(T00, v0, T01, v1, T02, v2);
    const float cov_z = __fadd_rn(dot3_ref(T10, v0, T11, v1, T12, v2), 0.3f);

    // conic, eigenvalues, radius (forward.cu:250-263)
    const float det = __fmaf_rn(cov_x, cov_z, -__fmul_rn(cov_y, cov_y));
    if (det == 0.0f) break;
    const float det_inv = __frcp_rn(det);
    const float con_x = __fmul_rn(cov_z, det_inv);
    const float con_y = __fmul_rn(cov_y, -det_inv);
    const float con_z = __fmul_rn(cov_x, det_inv);
    const float mid = __fmul_rn(__fadd_rn(cov_x, cov_z), 0.5f);
    const float disc = __fsqrt_rn(fmaxf(__fmaf_rn(mid, mid, -det), 0.1f));
    const float lam = fmaxf(__fadd_rn(mid, disc), __fadd_rn(mid, -disc));
    const int radius = __float2int_ru(__fmul_rn(__fsqrt_rn(lam), 3.0f));
    const float rf = (float)radius;

    // ndc2Pix in double (auxiliary.h:35-37): ((v + 1.0) * S - 1.0) * 0.5 with the mul-sub fused (DFMA)
    const float pix_x = (float)__dmul_rn(__fma_rn(__dadd_rn((double)proj_x, 1.0), (double)a.W, -1.0), 0.5);
    const float pix_y = (float)__dmul_rn(__fma_rn(__dadd_rn((double)proj_y, 1.0), (double)a.H, -1.0), 0.5);

    // getRect (auxiliary.h:39-46): float arithmetic, truncation toward zero, clamp to the tile grid
    const int rminx = min(a.grid_x, max(0, __float2int_rz(__fmul_rn(__fadd_rn(pix_x, -rf), 0.0625f))));
    const int rminy = min(a.grid_y, max(0, __float2int_rz(__fmul_rn(__fadd_rn(pix_y, -rf), 0.0625f))));
    const int rmaxx = min(
        a.grid_x, max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(pix_x, rf), 16.0f), -1.0f), 0.0625f))));
    const int rmaxy = min(
        a.grid_y, max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(pix_y, rf), 16.0f), -1.0f), 0.0625f))));
    const uint32_t tiles = (uint32_t)(rmaxx - rminx) * (uint32_t)(rmaxy - rminy);
    if (tiles == 0) break;
    rect_x0 = rminx, rect_y0 = rminy, rect_w = rmaxx - rminx;

    // Conservative extent of {pixels where the reference can reach alpha >= 1/255} (ours).
    // power = -q(d), q = 1/2 d^T Q d, Q = conic, Q^-1 = cov2D  =>  |dx| <= sqrt(2 tau cov_x).
    const float op = op_in;
    float ex = INFINITY, ey = INFINITY;
    if (op < 1.0f / 255.0f) {
      ex = ey = -INFINITY;  // alpha = op*exp(power<=0) <= op < 1/255 for every pixel
    } else {
      const float d2 = fmaxf(__fmaf_rn(mid, mid, -det), 0.0f);
      const float sq = __fsqrt_rn(d2);
      const float l_hi = mid + sq, l_lo = mid - sq;
      // condition number gate: float error of the reference's `power` is <= ~7e-7 * kappa relative
      if (l_lo > 0.f && l_hi < 1000.0f * l_lo) {
        const float tau = (logf(255.0f * op) + 1.0e-3f) * 1.004f + 1.0e-3f;
        ex = __fsqrt_rn(2.0f * tau * cov_x) * 1.001f + 1.0e-3f;
        ey = __fsqrt_rn(2.0f * tau * cov_z) * 1.001f + 1.0e-3f;
        if (!(ex >= 0.f) || !(ey >= 0.f)) ex = ey = INFINITY;
      }
    }
    rq0 = make_float4(pix_x, pix_y, ex, ey);
    rq1 = make_float4(con_x, con_y, con_z, op);
    rdepth = depth;
    out_radius = radius;
    out_tiles = tiles;
    survive = true;
  } while (false);

  // Instances per tile (tile_sort.cu). The Gaussians of a CTA are neighbours in the map and fall into a few dozen
  // tiles, and L2 serialises atomics to one address: the CTA counts per tile in a small shared-memory hash table
  // (shared-memory atomics cost ~4 SM-cycles per warp; merging lanes with MATCH.ANY costs ~58 and was a third of
  // this kernel's time) and sends ONE global atomic per tile it touched. A tile that finds no slot within kProbes
  // probes is counted in global memory directly.
  {
    for (uint32_t k = 0, rx = 0, t_row = (uint32_t)(rect_y0 * a.grid_x + rect_x0); k < out_tiles; ++k) {
      const uint32_t t = t_row + rx;  // row-major walk of the rectangle without dividing
      if (++rx == (uint32_t)rect_w) rx = 0, t_row += (uint32_t)a.grid_x;
      uint32_t slot = (t * 2654435761u) >> (32 - kTileSlotBits);
      bool counted = false;
      for (uint32_t pr = 0; pr < kTileProbes && !counted; ++pr, slot = (slot + 1) & (kTileSlots - 1)) {
        const uint32_t seen = atomicCAS(&s_tile_key[slot], 0xffffffffu, t);
        if (seen == 0xffffffffu || seen == t) {
          atomicAdd(&s_tile_cnt[slot], 1u);
          counted = true;
        }
      }
      if (!counted) atomicAdd(&a.tile_count[t], 1u);
    }
    __syncthreads();
    {  // this thread's four slots (one 16-byte load: most are empty)
      const uint4 c = reinterpret_cast<const uint4*>(s_tile_cnt)[tid];
      if (c.x | c.y | c.z | c.w) {
        const uint4 k4 = reinterpret_cast<const uint4*>(s_tile_key)[tid];
        if (c.x) atomicAdd(&a.tile_count[k4.x], c.x);
        if (c.y) atomicAdd(&a.tile_count[k4.y], c.y);
        if (c.z) atomicAdd(&a.tile_count[k4.z], c.z);
        if (c.w) atomicAdd(&a.tile_count[k4.w], c.w);
      }
    }
  }

  if (kVec16) {  // second pass: survivors the prefetch heuristic missed (uniform control flow)
    const bool late = survive && !prefetched;
    if (__syncthreads_or(late)) {  // also: everybody is done reading s_survive in the first pass
      s_survive[tid] = late;
      __syncthreads();
      const float4* g4 = reinterpret_cast<const float4*>(a.shs) + (size_t)base * 12;
      for (int q = tid, row = tid / 12, rem = tid % 12; q < n * 12; q += kBlock) {
        if (s_survive[row]) {
          const uint32_t dst = sh0 + (uint32_t)(row * kShRow + rem) * 16u;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g4 + q) : "memory");
        }
        { rem += kBlock % 12, row += kBlock / 12; if (rem >= 12) rem -= 12, ++row; }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }

  if (survive) {
    // colour: precomputed, or SH evaluated towards the camera (forward.cu:272-277)
    float rgb[3];
    uint32_t clamp_bits = 0;
    if (a.colors_precomp) {
      const float* cp = a.colors_precomp + (size_t)idx * 3;
      rgb[0] = __ldg(cp), rgb[1] = __ldg(cp + 1), rgb[2] = __ldg(cp + 2);
    } else {
      float3 dir = {px - s_cam[0], py - s_cam[1], pz - s_cam[2]};
      const float len = __fsqrt_rn(__fmaf_rn(dir.z, dir.z, __fmaf_rn(dir.x, dir.x, __fmul_rn(dir.y, dir.y))));
      dir.x = __fdiv_rn(dir.x, len), dir.y = __fdiv_rn(dir.y, len), dir.z = __fdiv_rn(dir.z, len);
      if (kVec16) {
        float4 c4v[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) c4v[i] = s_sh[tid * kShRow + i];
        const float* cf = reinterpret_cast<const float*>(c4v);
        eval_sh(a.D, dir, [&](int k, int c) { return cf[k * 3 + c]; }, rgb);
      } else {
        const float* sh = a.shs + (size_t)idx * a.M * 3;
        eval_sh(a.D, dir, [&](int k, int c) { return __ldg(sh + k * 3 + c); }, rgb);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (rgb[c] < 0.f) clamp_bits |= 1u << c;
        rgb[c] = fmaxf(rgb[c], 0.0f);
      }
    }
    SplatRecord rec;
    rec.q0 = rq0;
    rec.q1 = rq1;
    rec.q2 = make_float4(rgb[0], rgb[1], rgb[2], rdepth);
    a.rec[idx] = rec;
    a.clamped[idx] = (uint8_t)clamp_bits;
  }
  if (tid < n) {
    a.radii[idx] = out_radius;
    a.tiles_touched[idx] = out_tiles;
  }
}

__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* __restrict__ means3D,
                                                           const float* __restrict__ viewmatrix,
                                                           unsigned char* __restrict__ present) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const float x = __ldg(means3D + 3 * (size_t)idx), y = __ldg(means3D + 3 * (size_t)idx + 1),
              z = __ldg(means3D + 3 * (size_t)idx + 2);
  const float v2 = __ldg(viewmatrix + 2), v6 = __ldg(viewmatrix + 6), v10 = __ldg(viewmatrix + 10),
              v14 = __ldg(viewmatrix + 14);
  const float depth = __fadd_rn(__fmaf_rn(z, v10, __fmaf_rn(x, v2, __fmul_rn(y, v6))), v14);
  present[idx] = depth <= 0.2f ? 0 : 1;  // in_frustum, auxiliary.h:120-144
}

}  // namespace

cudaError_t launch_preprocess_fwd(const gsr_fwd_args& f, const GeomState& g, int* radii, uint32_t* tile_count,
                                  float focal_x, float focal_y, cudaStream_t s) {
  PreArgs a;
  a.P = f.P, a.D = f.D, a.M = f.M, a.W = f.width, a.H = f.height;
  a.grid_x = (f.width + kTile - 1) / kTile, a.grid_y = (f.height + kTile - 1) / kTile;
  a.scale_modifier = f.scale_modifier, a.tan_fovx = f.tan_fovx, a.tan_fovy = f.tan_fovy;
  a.focal_x = focal_x, a.focal_y = focal_y;
  a.means3D = f.means3D, a.scales = f.scales, a.rotations = f.rotations, a.opacities = f.opacities, a.shs = f.shs;
  a.cov3D_precomp = f.cov3D_precomp, a.colors_precomp = f.colors_precomp;
  a.viewmatrix = f.viewmatrix, a.projmatrix = f.projmatrix, a.cam_pos = f.cam_pos;
  a.radii = radii, a.rec = g.rec, a.tiles_touched = g.tiles_touched, a.clamped = g.clamped;
  a.tile_count = tile_count;
  a.act = f.activations;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  a.vec_ok = al16(f.means3D) && (!f.scales || al16(f.scales));
  const bool rot_ok = f.cov3D_precomp || al16(f.rotations);
  if (!rot_ok) return cudaErrorMisalignedAddress;
  const int blocks = (f.P + kBlock - 1) / kBlock;
  const bool vec16 = !f.colors_precomp && f.M == 16 && al16(f.shs);
  if (vec16) {
    constexpr int kSmem = kBlock * kShRow * (int)sizeof(float4);
    if (cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(preprocess_fwd_kernel<true>), kSmem)) return e;
    preprocess_fwd_kernel<true><<<blocks, kBlock, kSmem, s>>>(a);
  } else {
    preprocess_fwd_kernel<false><<<blocks, kBlock, 0, s>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t launch_mark_visible(int P, const float* means3D, const float* viewmatrix, unsigned char* present,
                                cudaStream_t s) {
  mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, viewmatrix, present);
  return cudaGetLastError();
}

}  // namespace gsr
