/*
 * TEST INFRASTRUCTURE — "Oracle B": scalar CPU restatement of GS-LIVM's rasterizer hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may load this
 * library; the product (libgsr_b200.so) never links or calls it.
 *
 * Pinning status: the reference ships NO golden vector or test for this path, so this
 * restatement is pinned against outputs of the reference itself — "Oracle A", the unmodified reference
 * CUDA sources compiled by oracle/Makefile into oracle/_ref/ — run on the GPU box by
 * tests/test_oracle_pinning.py, and against the fixtures under tests/golden/ generated from that run.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 */
#ifndef GSR_ORACLE_H_INCLUDED
#define GSR_ORACLE_H_INCLUDED

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gsro_inputs {
  int P, D, M, W, H;
  const float* bg;             /* [3] */
  const float* means3D;        /* [P,3] */
  const float* shs;            /* [P,M,3] or NULL */
  const float* colors_precomp; /* [P,3] or NULL */
  const float* opacities;      /* [P] */
  const float* scales;         /* [P,3] or NULL */
  const float* rotations;      /* [P,4] or NULL */
  const float* cov3D_precomp;  /* [P,6] or NULL */
  const float* view;           /* [16] */
  const float* proj;           /* [16] */
  const float* campos;         /* [3] */
  float scale_modifier, tan_fovx, tan_fovy;
} gsro_inputs;

/* src/cuda_rasterizer/forward.cu:179-286 (preprocessCUDA) with :19-25, :29-76, :79-133, :138-176 and
 * include/gs/cuda_rasterizer/auxiliary.h:35-64. Float operation order follows the reference's sm_100a
 * SASS (FMA contraction), so depths / means2D / radii / tiles_touched are bit-identical to the GPU. */
int gsro_preprocess(const gsro_inputs* in, int32_t* radii, float* means2D, float* depths, float* cov3D,
                    float* conic_opacity, float* rgb, uint8_t* clamped, uint32_t* tiles_touched);

/* cub::DeviceScan::InclusiveSum, rasterizer_impl.cu:270-273; returns R = offsets[P-1]. */
int64_t gsro_scan(int P, const uint32_t* tiles_touched, uint32_t* offsets);

/* duplicateWithKeys (rasterizer_impl.cu:64-101), stable sort on bits [0, 32+bit) (rasterizer_impl.cu:
 * 295-309, getHigherMsb :35-48), identifyTileRanges (:106-125, after the memset :311). */
int gsro_bin(int P, int64_t R, int W, int H, const int32_t* radii, const float* means2D, const float* depths,
             const uint32_t* offsets, uint64_t* keys_unsorted, uint32_t* vals_unsorted, uint64_t* keys,
             uint32_t* vals, uint32_t* ranges);

/* renderCUDA forward, src/cuda_rasterizer/forward.cu:291-407. `features` = rgb or colors_precomp.
 * fragile[pix] != 0 marks pixels where a discrete decision (power > 0, alpha < 1/255, T < 1e-4) was
 * within rounding distance of its threshold, i.e. where another correct expf may legitimately differ.
 * counters = {X evaluated pairs, X_p power<=0, X_a alpha>=1/255, X_b blended}. */
int gsro_blend_fwd(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                   const float* conic_opacity, const float* features, const float* depths, const float* bg,
                   float* out_color, float* out_depth, float* out_acc, float* final_T, uint32_t* n_contrib,
                   uint8_t* fragile, uint64_t counters[4], int nthreads);

/* renderCUDA backward, src/cuda_rasterizer/backward.cu:438-603: decisions in float exactly like the
 * forward, gradient arithmetic and accumulation in double (order-independent reference value for the
 * atomics). No depth gradient (backward.cu:567-571).
 * counters = {X' pairs visited (contributor < n_contrib), X'_p with power <= 0, X_b blended}. */
int gsro_blend_bwd(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                   const float* conic_opacity, const float* features, const float* bg, const float* final_T,
                   const uint32_t* n_contrib, const float* dL_dpix, const float* dL_dacc, double* dL_dmean2D,
                   double* dL_dconic, double* dL_dopacity, double* dL_dcolor, uint64_t counters[3], int nthreads);

/* computeCov2DCUDA (backward.cu:140-275) + preprocessCUDA backward (:371-435) with computeColorFromSH
 * backward (:20-135) and computeCov3D backward (:279-366); double arithmetic on the float forward state. */
int gsro_geom_bwd(const gsro_inputs* in, const int32_t* radii, const float* cov3D, const uint8_t* clamped,
                  const double* dL_dmean2D, const double* dL_dconic, const double* dL_dcolor, double* dL_dmean3D,
                  double* dL_dcov3D, double* dL_dsh, double* dL_dscale, double* dL_drot);

/* checkFrustum / in_frustum, rasterizer_impl.cu:52-60, auxiliary.h:120-144. */
int gsro_mark_visible(int P, const float* means3D, const float* view, uint8_t* present);

#ifdef __cplusplus
}
#endif
#endif
