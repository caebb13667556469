// Drop-in replacement for GS-LIVM include/gs/cuda_rasterizer/rasterizer.h:19-89.
//
// Declares CudaRasterizer::Rasterizer with the reference's exact static-member signatures so that
// src/gs/rasterize_points.cu (call sites :99, :184, :232) compiles unchanged. The definitions live in
// gs-livm_b200/csrc/rasterizer_shim.cpp and forward to the C ABI in ../gsr.h; errors reported by the
// C ABI are rethrown as std::runtime_error, which is what the reference does (rasterizer_impl.cu:229-231,
// auxiliary.h:146-154).
#pragma once

#include <functional>
#include <vector>

namespace CudaRasterizer {

using ScratchAllocator = std::function<char*(size_t)>;

class __attribute__((visibility("default"))) Rasterizer {
 public:
  // present[i] = (viewmatrix * means3D[i]).z > 0.2   (reference rasterizer_impl.cu:52-60)
  static void markVisible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present);

  // Returns the number of rendered (Gaussian, tile) instances R.
  static int forward(ScratchAllocator geometryBuffer, ScratchAllocator binningBuffer, ScratchAllocator imageBuffer,
                     const int P, int D, int M,
                     const float* background,
                     const int width, int height,
                     const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
                     const float* scales, const float scale_modifier, const float* rotations,
                     const float* cov3D_precomp,
                     const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                     const float tan_fovx, float tan_fovy,
                     const bool prefiltered,
                     float* out_color, float* out_depth, float* out_acc,
                     int* radii = nullptr, bool debug = false);

  static void backward(const int P, int D, int M, int R,
                       const float* background,
                       const int width, int height,
                       const float* means3D, const float* shs, const float* colors_precomp,
                       const float* scales, const float scale_modifier, const float* rotations,
                       const float* cov3D_precomp,
                       const float* viewmatrix, const float* projmatrix, const float* campos,
                       const float tan_fovx, float tan_fovy,
                       const int* radii,
                       char* geom_buffer, char* binning_buffer, char* image_buffer,
                       const float* dL_dpix, const float* dL_dacc,
                       float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                       float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                       bool debug);
};

}  // namespace CudaRasterizer
