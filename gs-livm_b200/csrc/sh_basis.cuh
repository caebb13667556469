// Real spherical-harmonics basis up to degree 3 as the reference evaluates it (constants: include/gs/cuda_rasterizer/
// auxiliary.h:22-33; polynomials: src/cuda_rasterizer/forward.cu:29-76, backward.cu:56-135), shared by geom_bwd.cu
// (dL_dsh = basis (x) dRGB and the direction gradient) and window_factored.cu (which rebuilds the same dL_dsh rows from
// the gathered dRGB of every view). The VALUES are written with _rn intrinsics so that both translation units get the
// same bits whatever the compiler contracts around them.
#pragma once

namespace gsr {

__device__ constexpr float kShC0 = 0.28209479177387814f;
__device__ constexpr float kShC1 = 0.4886025119029199f;
__device__ constexpr float kShC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                       -1.0925484305920792f, 0.5462742152960396f};
__device__ constexpr float kShC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                       0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                       -0.5900435899266435f};

// o = mean - camera centre and its normalisation (three IEEE divides by the length, like the reference:
// forward.cu:36-37, backward.cu:30-34), the same bits in both translation units
__device__ __forceinline__ void sh_direction(float mx, float my, float mz, float cx, float cy, float cz, float& ox,
                                             float& oy, float& oz, float& x, float& y, float& z) {
  ox = __fsub_rn(mx, cx), oy = __fsub_rn(my, cy), oz = __fsub_rn(mz, cz);
  const float len = __fsqrt_rn(__fmaf_rn(oz, oz, __fmaf_rn(oy, oy, __fmul_rn(ox, ox))));
  x = __fdiv_rn(ox, len), y = __fdiv_rn(oy, len), z = __fdiv_rn(oz, len);
}

// b[k] = basis_k(x, y, z) for k < (deg + 1)^2, 0 above
__device__ __forceinline__ void sh_basis_values(int deg, float x, float y, float z, float b[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) b[i] = 0.f;
  b[0] = kShC0;
  if (deg < 1) return;
  b[1] = __fmul_rn(-kShC1, y), b[2] = __fmul_rn(kShC1, z), b[3] = __fmul_rn(-kShC1, x);
  if (deg < 2) return;
  const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
  const float xy = __fmul_rn(x, y), yz = __fmul_rn(y, z), xz = __fmul_rn(x, z);
  const float xx_m_yy = __fsub_rn(xx, yy);
  const float zz2_m_xx_m_yy = __fsub_rn(__fsub_rn(__fmul_rn(2.f, zz), xx), yy);
  b[4] = __fmul_rn(kShC2[0], xy), b[5] = __fmul_rn(kShC2[1], yz), b[6] = __fmul_rn(kShC2[2], zz2_m_xx_m_yy);
  b[7] = __fmul_rn(kShC2[3], xz), b[8] = __fmul_rn(kShC2[4], xx_m_yy);
  if (deg < 3) return;
  const float zz4_m_xx_m_yy = __fsub_rn(__fsub_rn(__fmul_rn(4.f, zz), xx), yy);
  b[9] = __fmul_rn(__fmul_rn(kShC3[0], y), __fsub_rn(__fmul_rn(3.f, xx), yy));
  b[10] = __fmul_rn(__fmul_rn(kShC3[1], xy), z);
  b[11] = __fmul_rn(__fmul_rn(kShC3[2], y), zz4_m_xx_m_yy);
  b[12] = __fmul_rn(__fmul_rn(kShC3[3], z), __fsub_rn(__fsub_rn(__fmul_rn(2.f, zz), __fmul_rn(3.f, xx)), __fmul_rn(3.f, yy)));
  b[13] = __fmul_rn(__fmul_rn(kShC3[4], x), zz4_m_xx_m_yy);
  b[14] = __fmul_rn(__fmul_rn(kShC3[5], z), xx_m_yy);
  b[15] = __fmul_rn(__fmul_rn(kShC3[6], x), __fsub_rn(xx, __fmul_rn(3.f, yy)));
}

// partial derivatives of the basis values w.r.t. the (normalised) direction components
__device__ __forceinline__ void sh_basis_grads(int deg, float x, float y, float z, float bx[16], float by[16],
                                               float bz[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) bx[i] = by[i] = bz[i] = 0.f;
  if (deg < 1) return;
  by[1] = -kShC1, bz[2] = kShC1, bx[3] = -kShC1;
  if (deg < 2) return;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  bx[4] = kShC2[0] * y, by[4] = kShC2[0] * x;
  by[5] = kShC2[1] * z, bz[5] = kShC2[1] * y;
  bx[6] = kShC2[2] * -2.f * x, by[6] = kShC2[2] * -2.f * y, bz[6] = kShC2[2] * 4.f * z;
  bx[7] = kShC2[3] * z, bz[7] = kShC2[3] * x;
  bx[8] = kShC2[4] * 2.f * x, by[8] = kShC2[4] * -2.f * y;
  if (deg < 3) return;
  bx[9] = kShC3[0] * 6.f * xy, by[9] = kShC3[0] * 3.f * (xx - yy);
  bx[10] = kShC3[1] * yz, by[10] = kShC3[1] * xz, bz[10] = kShC3[1] * xy;
  bx[11] = kShC3[2] * -2.f * xy, by[11] = kShC3[2] * (4.f * zz - xx - 3.f * yy), bz[11] = kShC3[2] * 8.f * yz;
  bx[12] = kShC3[3] * -6.f * xz, by[12] = kShC3[3] * -6.f * yz, bz[12] = kShC3[3] * 3.f * (2.f * zz - xx - yy);
  bx[13] = kShC3[4] * (4.f * zz - 3.f * xx - yy), by[13] = kShC3[4] * -2.f * xy, bz[13] = kShC3[4] * 8.f * xz;
  bx[14] = kShC3[5] * 2.f * xz, by[14] = kShC3[5] * -2.f * yz, bz[14] = kShC3[5] * (xx - yy);
  bx[15] = kShC3[6] * 3.f * (xx - yy), by[15] = kShC3[6] * -6.f * xy;
}

}  // namespace gsr
