// Per-Gaussian backward: dL/d{conic, mean2D, colour} -> dL/d{mean3D, cov3D, SH, scale, rotation}. sm_100a.
//
// One fused kernel replaces the reference's two launches computeCov2DCUDA (src/cuda_rasterizer/
// backward.cu:140-275) and preprocessCUDA<3> (:371-435) with its helpers computeColorFromSH backward
// (:20-135) and computeCov3D backward (:279-366); fusing removes the dL_dmean3D read-modify-write
// and one pass over means/radii (559 instead of 623 algorithmic B per visible Gaussian).
// The 3D covariance is recomputed from scale/rotation with the forward's exact operation DAG instead
// of being stored by the forward and re-read here.
//
// Reference quirks kept on purpose: gradients w.r.t. the quaternion AS PASSED (no normalisation
// Jacobian, backward.cu:361-365), dL_dscale w.r.t. the modifier-scaled scale (:340-342), clamped
// colour channels pass no gradient (:41-43), Gaussians with radii <= 0 are skipped (outputs stay at the
// caller's zeros).
#include "gsr_common.cuh"
#include "sh_basis.cuh"

namespace gsr {
namespace {

struct GeomBwdArgs {
  int P, D, M;
  float scale_modifier, tan_fovx, tan_fovy, focal_x, focal_y;
  const float* means3D;
  const int* radii;
  const float* shs;
  const uint8_t* clamped;
  const float* scales;
  const float* rotations;
  const float* cov3D_precomp;
  const float* viewmatrix;
  const float* projmatrix;
  const float* cam_pos;
  const float* dL_dmean2D;  // [P,3]
  const float* dL_dconic;   // [P,4]
  const float* dL_dcolor;   // [P,3]
  float* dL_dmean3D;
  float* dL_dcov3D;
  float* dL_dsh;
  float* dL_dscale;
  float* dL_drot;
  int first;   // the kernel covers Gaussians [first, P): P is the END of the range here
  int vec_sh;  // M == 16 and shs / dL_dsh 16-byte aligned
  int assign_all;  // GSR_BWD_OUTPUTS_UNINITIALIZED: also write the zeros of Gaussians with radii <= 0
  int act;                 // GSR_ACT_* bits: scales / rotations / opacities are raw parameters (see gsr.h)
  const float* opacities;  // raw, with GSR_ACT_OPACITY_SIGMOID
  float* dL_dopacity;      // accumulated by the blend backward w.r.t. the activated opacity; rescaled in place
  uint8_t* row_state;  // gsr_bwd_args.row_state (with assign_all): != 0 <=> the Gaussian's outputs may be non-zero; a culled
                       // Gaussian whose outputs are known to be zero already is skipped, not rewritten
};

__device__ __forceinline__ float dot3_ref(float a, float b, float c, float d, float e, float f) {
  return __fmaf_rn(e, f, __fmaf_rn(a, b, __fmul_rn(c, d)));
}
__device__ __forceinline__ float xform_row(const float* m, int r, float x, float y, float z) {
  return __fadd_rn(dot3_ref(x, m[r], y, m[4 + r], z, m[8 + r]), m[12 + r]);
}

// Gaussians per CTA. The kernel alternates between a memory phase (SH rows in, dL_dsh rows out) and ~1000
// instructions of covariance / SH math; with 121 registers only 16 warps fit an SM, and eight 64-thread CTAs
// interleave those phases better than two 256-thread ones (97 -> 91 us).
constexpr int kBlock = 64;
constexpr int kShRow = 13;  // float4 per staged SH row: 12 used + 1 pad -> conflict-free lane-strided LDS/STS.128

// kVec16: M == 16 and shs / dL_dsh 16-byte aligned. SH rows of the block's visible Gaussians are staged into
// shared memory with coalesced 16-byte cp.async transfers (issued first, so they overlap the covariance
// math), and the dL_dsh rows leave the same way: written row-wise into shared memory, stored coalesced.
// kFactored: GSR_BWD_SH_FACTORED -- dL_dsh is not written at all (window_factored.cu rebuilds the window's sum of it
// from the views' dRGB); the SH rows are still read for the direction gradient.
template <bool kVec16, bool kFactored>
__global__ void __launch_bounds__(kBlock) geom_bwd_kernel(const GeomBwdArgs a) {
  __shared__ float s_view[16], s_proj[16], s_cam[3];
  __shared__ uint8_t s_active[kBlock];
  extern __shared__ __align__(16) float4 s_sh[];  // [kBlock][kShRow] when kVec16
  const int tid = threadIdx.x;
  for (int i = tid; i < 35; i += kBlock) {  // the camera: 16 + 16 + 3 floats
    if (i < 16) s_view[i] = __ldg(a.viewmatrix + i);
    else if (i < 32) s_proj[i - 16] = __ldg(a.projmatrix + i - 16);
    else s_cam[i - 32] = __ldg(a.cam_pos + i - 32);
  }
  const int base = a.first + blockIdx.x * kBlock;
  const int n = min(kBlock, a.P - base);
  const int idx = base + tid;
  const bool active = tid < n && a.radii[idx] > 0;
  // requested here, looked at only when the outputs are written: the kernel is latency-bound, and a dependent load in
  // front of the SH-row prefetch costs more than the zero rows it saves (measured: 95 vs 88 us)
  const uint8_t may_be_nonzero = (a.row_state && tid < n) ? a.row_state[idx] : (uint8_t)1;
  s_active[tid] = active;
  __syncthreads();
  if (kVec16) {
    const float4* g4 = reinterpret_cast<const float4*>(a.shs) + (size_t)base * 12;
    const uint32_t sh0 = smem_u32(s_sh);  // (once: the conversion is not hoisted out of the loop by itself)
    for (int q = tid, row = tid / 12, rem = tid % 12; q < n * 12; q += kBlock) {  // row = q / 12 without dividing
      if (s_active[row]) {
        const uint32_t dst = sh0 + (uint32_t)(row * kShRow + rem) * 16u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g4 + q) : "memory");
      }
      { rem += kBlock % 12, row += kBlock / 12; if (rem >= 12) rem -= 12, ++row; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  float gmx = 0.f, gmy = 0.f, gmz = 0.f;
  float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float sx = 0.f, sy = 0.f, sz = 0.f, qr = 0.f, qx = 0.f, qy = 0.f, qz = 0.f;
  float A0 = 0, A1 = 0, A2 = 0, B0 = 0, B1 = 0, B2 = 0, C0 = 0, C1 = 0, C2 = 0;  // rows of M = S*R
  float R00 = 0, R01 = 0, R02 = 0, R10 = 0, R11 = 0, R12 = 0, R20 = 0, R21 = 0, R22 = 0;
  float mx = 0.f, my = 0.f, mz = 0.f;
  uint32_t cl_early = 0;
  float dRGB_early[3] = {0.f, 0.f, 0.f};
  float act_s0 = 1.f, act_s1 = 1.f, act_s2 = 1.f, q_norm = 1.f;  // activated scales, |raw quaternion| (GSR_ACT_*)
  if (active) {

  const float* V = s_view;
  const float* PM = s_proj;
  mx = __ldg(a.means3D + 3 * (size_t)idx), my = __ldg(a.means3D + 3 * (size_t)idx + 1),
  mz = __ldg(a.means3D + 3 * (size_t)idx + 2);
  // every other per-Gaussian input is requested here as well, so that one memory latency is exposed, not five
  const float4 dcon = __ldg(reinterpret_cast<const float4*>(a.dL_dconic) + idx);  // x, y, (unused), w
  const float d2x = __ldg(a.dL_dmean2D + 3 * (size_t)idx), d2y = __ldg(a.dL_dmean2D + 3 * (size_t)idx + 1);
  if (a.shs) {
    cl_early = a.clamped[idx];
    dRGB_early[0] = __ldg(a.dL_dcolor + 3 * (size_t)idx), dRGB_early[1] = __ldg(a.dL_dcolor + 3 * (size_t)idx + 1);
    dRGB_early[2] = __ldg(a.dL_dcolor + 3 * (size_t)idx + 2);
  }

  // ---- forward recomputation: 3D covariance (same DAG as preprocess.cu) --------------------------
  float c0, c1, c2, c3, c4, c5;
  if (a.cov3D_precomp) {
    const float* cp = a.cov3D_precomp + (size_t)idx * 6;
    c0 = __ldg(cp), c1 = __ldg(cp + 1), c2 = __ldg(cp + 2), c3 = __ldg(cp + 3), c4 = __ldg(cp + 4), c5 = __ldg(cp + 5);
  } else {
    float s0 = __ldg(a.scales + 3 * (size_t)idx), s1 = __ldg(a.scales + 3 * (size_t)idx + 1),
          s2 = __ldg(a.scales + 3 * (size_t)idx + 2);
    if (a.act & GSR_ACT_SCALE_EXP) s0 = expf(s0), s1 = expf(s1), s2 = expf(s2);
    act_s0 = s0, act_s1 = s1, act_s2 = s2;
    sx = __fmul_rn(s0, a.scale_modifier), sy = __fmul_rn(s1, a.scale_modifier), sz = __fmul_rn(s2, a.scale_modifier);
    float4 q = __ldg(reinterpret_cast<const float4*>(a.rotations) + idx);
    if (a.act & GSR_ACT_ROTATION_NORMALIZE) q = normalize_quat(q, &q_norm);
    qr = q.x, qx = q.y, qy = q.z, qz = q.w;
    const float xz = __fmul_rn(qx, qz), rx = __fmul_rn(qr, qx), rz = __fmul_rn(qr, qz);
    const float yy = __fmul_rn(qy, qy), zz = __fmul_rn(qz, qz);
    const float xz_p_ry = __fmaf_rn(qr, qy, xz), yz_p_rx = __fmaf_rn(qy, qz, rx);
    const float xz_m_ry = __fmaf_rn(-qr, qy, xz), xy_m_rz = __fmaf_rn(qx, qy, -rz);
    const float xy_p_rz = __fmaf_rn(qx, qy, rz), yz_m_rx = __fmaf_rn(qy, qz, -rx);
    const float xx_p_yy = __fmaf_rn(qx, qx, yy), yy_p_zz = __fadd_rn(yy, zz), xx_p_zz = __fmaf_rn(qx, qx, zz);
    R00 = __fadd_rn(-__fadd_rn(yy_p_zz, yy_p_zz), 1.0f), R01 = __fadd_rn(xy_p_rz, xy_p_rz),
    R02 = __fadd_rn(xz_m_ry, xz_m_ry);
    R10 = __fadd_rn(xy_m_rz, xy_m_rz), R11 = __fadd_rn(-__fadd_rn(xx_p_zz, xx_p_zz), 1.0f),
    R12 = __fadd_rn(yz_p_rx, yz_p_rx);
    R20 = __fadd_rn(xz_p_ry, xz_p_ry), R21 = __fadd_rn(yz_m_rx, yz_m_rx),
    R22 = __fadd_rn(-__fadd_rn(xx_p_yy, xx_p_yy), 1.0f);
    A0 = __fmul_rn(sx, R00), A1 = __fmul_rn(sx, R01), A2 = __fmul_rn(sx, R02);
    B0 = __fmul_rn(sy, R10), B1 = __fmul_rn(sy, R11), B2 = __fmul_rn(sy, R12);
    C0 = __fmul_rn(sz, R20), C1 = __fmul_rn(sz, R21), C2 = __fmul_rn(sz, R22);
    c0 = dot3_ref(A0, A0, B0, B0, C0, C0), c1 = dot3_ref(A0, A1, B0, B1, C0, C1);
    c2 = dot3_ref(A0, A2, B0, B2, C0, C2), c3 = dot3_ref(A1, A1, B1, B1, C1, C1);
    c4 = dot3_ref(A1, A2, B1, B2, C1, C2), c5 = dot3_ref(A2, A2, B2, B2, C2, C2);
  }

  // ---- conic backward (reference computeCov2DCUDA) ----------------------------------------------
  const float tz = xform_row(V, 2, mx, my, mz);
  const float tx0 = xform_row(V, 0, mx, my, mz), ty0 = xform_row(V, 1, mx, my, mz);
  const float limx = __fmul_rn(a.tan_fovx, 1.3f), limy = __fmul_rn(a.tan_fovy, 1.3f);
  const float txtz = __fdiv_rn(tx0, tz), tytz = __fdiv_rn(ty0, tz);
  const float tx = __fmul_rn(fminf(fmaxf(txtz, -limx), limx), tz);
  const float ty = __fmul_rn(fminf(fmaxf(tytz, -limy), limy), tz);
  const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
  const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
  const float J00 = __fdiv_rn(a.focal_x, tz), J11 = __fdiv_rn(a.focal_y, tz);
  const float tz_sq = __fmul_rn(tz, tz);
  const float J02 = __fdiv_rn(-__fmul_rn(a.focal_x, tx), tz_sq), J12 = __fdiv_rn(-__fmul_rn(a.focal_y, ty), tz_sq);
  const float T00 = __fmaf_rn(J02, V[2], __fmul_rn(V[0], J00)), T01 = __fmaf_rn(J02, V[6], __fmul_rn(V[4], J00)),
              T02 = __fmaf_rn(J02, V[10], __fmul_rn(V[8], J00));
  const float T10 = __fmaf_rn(J12, V[2], __fmul_rn(V[1], J11)), T11 = __fmaf_rn(J12, V[6], __fmul_rn(V[5], J11)),
              T12 = __fmaf_rn(J12, V[10], __fmul_rn(V[9], J11));
  const float u0 = dot3_ref(T00, c0, T01, c1, T02, c2), u1 = dot3_ref(T00, c1, T01, c3, T02, c4),
              u2 = dot3_ref(T00, c2, T01, c4, T02, c5);
  const float v0 = dot3_ref(T10, c0, T11, c1, T12, c2), v1 = dot3_ref(T10, c1, T11, c3, T12, c4),
              v2 = dot3_ref(T10, c2, T11, c4, T12, c5);
  const float ca = __fadd_rn(dot3_ref(T00, u0, T01, u1, T02, u2), 0.3f);
  const float cb = dot3_ref(T00, v0, T01, v1, T02, v2);
  const float cc = __fadd_rn(dot3_ref(T10, v0, T11, v1, T12, v2), 0.3f);

  const float denom = ca * cc - cb * cb;
  const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
  float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
  if (denom2inv != 0.f) {
    dL_da = denom2inv * (-cc * cc * dcon.x + 2.f * cb * cc * dcon.y + (denom - ca * cc) * dcon.w);
    dL_dc = denom2inv * (-ca * ca * dcon.w + 2.f * ca * cb * dcon.y + (denom - ca * cc) * dcon.x);
    dL_db = denom2inv * 2.f * (cb * cc * dcon.x - (denom + 2.f * cb * cb) * dcon.y + ca * cb * dcon.w);
    dcov[0] = T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc;
    dcov[3] = T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc;
    dcov[5] = T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc;
    dcov[1] = 2.f * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2.f * T10 * T11 * dL_dc;
    dcov[2] = 2.f * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2.f * T10 * T12 * dL_dc;
    dcov[4] = 2.f * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2.f * T11 * T12 * dL_dc;
  }
  {
    float* o = a.dL_dcov3D + (size_t)idx * 6;
#pragma unroll
    for (int i = 0; i < 6; ++i) o[i] = dcov[i];
  }
  // dL/dT (upper 2x3), with u_k = sum_j T0j Vrk[j][k], v_k = sum_j T1j Vrk[j][k]
  const float dT00 = 2.f * u0 * dL_da + v0 * dL_db, dT01 = 2.f * u1 * dL_da + v1 * dL_db,
              dT02 = 2.f * u2 * dL_da + v2 * dL_db;
  const float dT10 = 2.f * v0 * dL_dc + u0 * dL_db, dT11 = 2.f * v1 * dL_dc + u1 * dL_db,
              dT12 = 2.f * v2 * dL_dc + u2 * dL_db;
  const float dJ00 = V[0] * dT00 + V[4] * dT01 + V[8] * dT02;
  const float dJ02 = V[2] * dT00 + V[6] * dT01 + V[10] * dT02;
  const float dJ11 = V[1] * dT10 + V[5] * dT11 + V[9] * dT12;
  const float dJ12 = V[2] * dT10 + V[6] * dT11 + V[10] * dT12;
  const float itz = 1.f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
  const float dtx = x_grad_mul * -a.focal_x * itz2 * dJ02;
  const float dty = y_grad_mul * -a.focal_y * itz2 * dJ12;
  const float dtz = -a.focal_x * itz2 * dJ00 - a.focal_y * itz2 * dJ11 + (2.f * a.focal_x * tx) * itz3 * dJ02 +
                    (2.f * a.focal_y * ty) * itz3 * dJ12;
  gmx = V[0] * dtx + V[1] * dty + V[2] * dtz;
  gmy = V[4] * dtx + V[5] * dty + V[6] * dtz;
  gmz = V[8] * dtx + V[9] * dty + V[10] * dtz;

  // ---- projection backward: dL/dmean2D -> dL/dmean3D (reference preprocessCUDA, backward.cu:399-416) ----
  {
    const float hw = (PM[3] * mx + PM[7] * my + PM[11] * mz + PM[15]) + 0.0000001f;
    const float m_w = 1.0f / hw;
    const float mul1 = (PM[0] * mx + PM[4] * my + PM[8] * mz + PM[12]) * m_w * m_w;
    const float mul2 = (PM[1] * mx + PM[5] * my + PM[9] * mz + PM[13]) * m_w * m_w;
    gmx += (PM[0] * m_w - PM[3] * mul1) * d2x + (PM[1] * m_w - PM[3] * mul2) * d2y;
    gmy += (PM[4] * m_w - PM[7] * mul1) * d2x + (PM[5] * m_w - PM[7] * mul2) * d2y;
    gmz += (PM[8] * m_w - PM[11] * mul1) * d2x + (PM[9] * m_w - PM[11] * mul2) * d2y;
  }

    // ---- scale / rotation backward (reference computeCov3D backward) ------------------------------
    if (a.scales && !a.cov3D_precomp) {
      // dL/dSigma as a symmetric matrix; dL/dM = 2 M dL/dSigma, M rows = (A,B,C)
      const float S00 = dcov[0], S01 = 0.5f * dcov[1], S02 = 0.5f * dcov[2], S11 = dcov[3], S12 = 0.5f * dcov[4],
                  S22 = dcov[5];
      const float dA0 = 2.f * (A0 * S00 + A1 * S01 + A2 * S02), dA1 = 2.f * (A0 * S01 + A1 * S11 + A2 * S12),
                  dA2 = 2.f * (A0 * S02 + A1 * S12 + A2 * S22);
      const float dB0 = 2.f * (B0 * S00 + B1 * S01 + B2 * S02), dB1 = 2.f * (B0 * S01 + B1 * S11 + B2 * S12),
                  dB2 = 2.f * (B0 * S02 + B1 * S12 + B2 * S22);
      const float dC0 = 2.f * (C0 * S00 + C1 * S01 + C2 * S02), dC1 = 2.f * (C0 * S01 + C1 * S11 + C2 * S12),
                  dC2 = 2.f * (C0 * S02 + C1 * S12 + C2 * S22);
      // (GSR_ACT_SCALE_EXP: d exp(x)/dx = exp(x), torch's exp backward: grad * result)
      const bool raw_s = a.act & GSR_ACT_SCALE_EXP;
      a.dL_dscale[3 * (size_t)idx] = (R00 * dA0 + R01 * dA1 + R02 * dA2) * (raw_s ? act_s0 : 1.f);
      a.dL_dscale[3 * (size_t)idx + 1] = (R10 * dB0 + R11 * dB1 + R12 * dB2) * (raw_s ? act_s1 : 1.f);
      a.dL_dscale[3 * (size_t)idx + 2] = (R20 * dC0 + R21 * dC1 + R22 * dC2) * (raw_s ? act_s2 : 1.f);
      // E[i][j] = s_i * dM[i][j]
      const float E00 = sx * dA0, E01 = sx * dA1, E02 = sx * dA2;
      const float E10 = sy * dB0, E11 = sy * dB1, E12 = sy * dB2;
      const float E20 = sz * dC0, E21 = sz * dC1, E22 = sz * dC2;
      float4 dq;
      dq.x = 2.f * qz * (E01 - E10) + 2.f * qy * (E20 - E02) + 2.f * qx * (E12 - E21);
      dq.y = 2.f * qy * (E10 + E01) + 2.f * qz * (E20 + E02) + 2.f * qr * (E12 - E21) - 4.f * qx * (E22 + E11);
      dq.z = 2.f * qx * (E10 + E01) + 2.f * qr * (E20 - E02) + 2.f * qz * (E12 + E21) - 4.f * qy * (E22 + E00);
      dq.w = 2.f * qr * (E01 - E10) + 2.f * qx * (E20 + E02) + 2.f * qy * (E12 + E21) - 4.f * qz * (E11 + E00);
      if (a.act & GSR_ACT_ROTATION_NORMALIZE) {  // through q / max(|q|, eps): (g - qn (qn . g)) / |q|
        const float d = dq.x * qr + dq.y * qx + dq.z * qy + dq.w * qz;
        const float inv = 1.0f / q_norm;
        dq = make_float4((dq.x - qr * d) * inv, (dq.y - qx * d) * inv, (dq.z - qy * d) * inv, (dq.w - qz * d) * inv);
      }
      reinterpret_cast<float4*>(a.dL_drot)[idx] = dq;
    }
    if (a.act & GSR_ACT_OPACITY_SIGMOID) {  // torch's sigmoid backward: grad * (1 - y) * y
      const float y = sigmoid_ref(__ldg(a.opacities + idx));
      a.dL_dopacity[idx] = a.dL_dopacity[idx] * ((1.0f - y) * y);
    }

  }  // if (active): covariance + projection part

  // ---- SH backward (reference computeColorFromSH backward) -------------------------------------
  if (kVec16) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // every thread's part of the staged rows has landed
  }
  if (active && a.shs) {
    float ox, oy, oz, x, y, z;
    sh_direction(mx, my, mz, s_cam[0], s_cam[1], s_cam[2], ox, oy, oz, x, y, z);
    const uint32_t cl = cl_early;
    float dRGB[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) dRGB[c] = ((cl >> c) & 1u) ? 0.f : dRGB_early[c];
    float b[16], bx[16], by[16], bz[16];
    sh_basis_values(a.D, x, y, z, b);
    sh_basis_grads(a.D, x, y, z, bx, by, bz);
    const int ncoef = (a.D + 1) * (a.D + 1);
    float ddx = 0.f, ddy = 0.f, ddz = 0.f;  // dL/d(dir)
    if (kVec16) {
      // four coefficients (three 16-byte words of the staged row) at a time: 12 values in, 12 out -- not the whole
      // 48-float row in registers twice
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float4 in4[3], out4[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) in4[i] = s_sh[tid * kShRow + 3 * g + i];
        const float* sh = reinterpret_cast<const float*>(in4);
        float* o = reinterpret_cast<float*>(out4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = 4 * g + kk;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const bool on = k < ncoef;
            if (!kFactored) o[kk * 3 + c] = on ? __fmul_rn(b[k], dRGB[c]) : 0.f;
            if (on && k > 0) {
              const float w = sh[kk * 3 + c] * dRGB[c];
              ddx += bx[k] * w, ddy += by[k] * w, ddz += bz[k] * w;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (!kFactored) s_sh[tid * kShRow + 3 * g + i] = out4[i];  // own row only: no hazard
      }
    } else {
      const float* sh = a.shs + (size_t)idx * a.M * 3;
      float* o = a.dL_dsh + (size_t)idx * a.M * 3;
      for (int k = 0; k < ncoef && k < a.M; ++k) {
        for (int c = 0; c < 3; ++c) {
          if (!kFactored) o[k * 3 + c] = __fmul_rn(b[k], dRGB[c]);
          if (k > 0) {
            const float w = __ldg(sh + k * 3 + c) * dRGB[c];
            ddx += bx[k] * w, ddy += by[k] * w, ddz += bz[k] * w;
          }
        }
      }
      if (a.assign_all && !kFactored)  // coefficients above the active degree
        for (int k = ncoef * 3; k < a.M * 3; ++k) o[k] = 0.f;
    }
    // through the normalisation dir = o/|o| (reference dnormvdv, auxiliary.h:93-103)
    const float sum2 = ox * ox + oy * oy + oz * oz;
    const float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    gmx += ((sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * inv32;
    gmy += (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * inv32;
    gmz += (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * inv32;
  }
  // 1: the Gaussian's row goes out; 2: culled and the outputs are not known to be zero already: zeros go out; 0: nothing
  const uint8_t todo = active ? 1 : ((a.assign_all && tid < n && may_be_nonzero) ? 2 : 0);
  if (a.row_state && tid < n && may_be_nonzero != (uint8_t)active) a.row_state[idx] = active;
  if (kVec16 && !kFactored) s_active[tid] = todo;  // (its readers below come after the barrier)
  if (kVec16 && !kFactored) {  // coalesced store of the visible Gaussians' dL_dsh rows (culled rows stay at the caller's zeros)
    __syncthreads();
    float4* d4 = reinterpret_cast<float4*>(a.dL_dsh) + (size_t)base * 12;
    for (int q = tid, row = tid / 12, rem = tid % 12; q < n * 12; q += kBlock) {
      if (s_active[row] == 1) d4[q] = s_sh[row * kShRow + rem];
      else if (s_active[row] == 2) d4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      { rem += kBlock % 12, row += kBlock / 12; if (rem >= 12) rem -= 12, ++row; }
    }
  }
  if (!active) {
    if (todo == 2) {  // the caller did not zero the outputs: culled Gaussians get their zeros here
      const size_t i = (size_t)idx;
      a.dL_dmean3D[3 * i] = 0.f, a.dL_dmean3D[3 * i + 1] = 0.f, a.dL_dmean3D[3 * i + 2] = 0.f;
#pragma unroll
      for (int k = 0; k < 6; ++k) a.dL_dcov3D[6 * i + k] = 0.f;
      if (a.scales && !a.cov3D_precomp) {
        a.dL_dscale[3 * i] = 0.f, a.dL_dscale[3 * i + 1] = 0.f, a.dL_dscale[3 * i + 2] = 0.f;
        reinterpret_cast<float4*>(a.dL_drot)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (!kVec16 && !kFactored && a.shs)
        for (int k = 0; k < a.M * 3; ++k) a.dL_dsh[i * a.M * 3 + k] = 0.f;
    }
    return;
  }
  a.dL_dmean3D[3 * (size_t)idx] = gmx;
  a.dL_dmean3D[3 * (size_t)idx + 1] = gmy;
  a.dL_dmean3D[3 * (size_t)idx + 2] = gmz;

}

}  // namespace

cudaError_t launch_geom_bwd(const gsr_bwd_args& f, const GeomState& g, const int* radii, float focal_x, float focal_y,
                            cudaStream_t s) {
  GeomBwdArgs a;
  a.P = f.P, a.D = f.D, a.M = f.M, a.first = 0;
  if ((f.flags & GSR_BWD_ONLY_GEOM) && f.range_count > 0) {
    if (f.range_first < 0 || f.range_first + (long long)f.range_count > f.P) return cudaErrorInvalidValue;
    a.first = f.range_first, a.P = f.range_first + f.range_count;
  }
  a.scale_modifier = f.scale_modifier, a.tan_fovx = f.tan_fovx, a.tan_fovy = f.tan_fovy;
  a.focal_x = focal_x, a.focal_y = focal_y;
  a.means3D = f.means3D, a.radii = radii, a.shs = f.colors_precomp ? nullptr : f.shs, a.clamped = g.clamped;
  a.scales = f.scales, a.rotations = f.rotations, a.cov3D_precomp = f.cov3D_precomp;
  a.viewmatrix = f.viewmatrix, a.projmatrix = f.projmatrix, a.cam_pos = f.cam_pos;
  a.dL_dmean2D = f.dL_dmean2D, a.dL_dconic = f.dL_dconic, a.dL_dcolor = f.dL_dcolor;
  a.dL_dmean3D = f.dL_dmean3D, a.dL_dcov3D = f.dL_dcov3D, a.dL_dsh = f.dL_dsh, a.dL_dscale = f.dL_dscale,
  a.dL_drot = f.dL_drot;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!al16(f.dL_dconic) || (!f.cov3D_precomp && (!al16(f.rotations) || !al16(f.dL_drot))))
    return cudaErrorMisalignedAddress;
  const bool factored = a.shs && (f.flags & GSR_BWD_SH_FACTORED);
  a.vec_sh = a.shs && f.M == 16 && al16(f.shs) && (factored || al16(f.dL_dsh));
  a.assign_all = (f.flags & GSR_BWD_OUTPUTS_UNINITIALIZED) ? 1 : 0;
  a.row_state = a.assign_all ? f.row_state : nullptr;
  a.act = f.activations, a.opacities = f.opacities, a.dL_dopacity = f.dL_dopacity;
  if ((a.act & GSR_ACT_OPACITY_SIGMOID) && !f.opacities) return cudaErrorInvalidValue;
  const int blocks = (a.P - a.first + kBlock - 1) / kBlock;
  if (a.vec_sh) {
    constexpr int kSmem = kBlock * kShRow * (int)sizeof(float4);
    if (factored) {
      if (cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(geom_bwd_kernel<true, true>), kSmem)) return e;
      geom_bwd_kernel<true, true><<<blocks, kBlock, kSmem, s>>>(a);
    } else {
      if (cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(geom_bwd_kernel<true, false>), kSmem)) return e;
      geom_bwd_kernel<true, false><<<blocks, kBlock, kSmem, s>>>(a);
    }
  } else if (factored) {
    geom_bwd_kernel<false, true><<<blocks, kBlock, 0, s>>>(a);
  } else {
    geom_bwd_kernel<false, false><<<blocks, kBlock, 0, s>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace gsr
