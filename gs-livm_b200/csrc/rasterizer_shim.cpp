// C++ operator-surface shim: CudaRasterizer::Rasterizer::{forward, backward, markVisible} with the
// reference's exact signatures (include/gs/cuda_rasterizer/rasterizer.h:19-89), implemented on top of
// the C ABI in include/gsr.h. This is what lets GS-LIVM's src/gs/rasterize_points.cu (call sites :99,
// :184, :232) and src/gs/rasterizer.cu compile and link unchanged against libgsr_b200.so.
//
// Behavioural parity with the reference:
//  * kernels are enqueued on the stream set with gsr_shim_set_stream(); the default (NULL) is the
//    legacy default stream, which is where the reference launches everything;
//  * a failing stage surfaces as std::runtime_error (reference: rasterizer_impl.cu:229-231 and
//    CHECK_CUDA in auxiliary.h:146-154) — the C ABI itself never throws;
//  * the std::function allocators are invoked once each, geometry -> image -> binning.
#include <stdexcept>
#include <string>

#include "cuda_rasterizer/rasterizer.h"
#include "gsr.h"

namespace {

char* call_std_function(void* ctx, size_t bytes) {
  auto* fn = static_cast<CudaRasterizer::ScratchAllocator*>(ctx);
  return (*fn)(bytes);
}

[[noreturn]] void raise(const char* where) {
  throw std::runtime_error(std::string(where) + ": " + gsr_last_error());
}

}  // namespace

namespace CudaRasterizer {

void Rasterizer::markVisible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
  static_assert(sizeof(bool) == 1, "present[] is written as bytes");
  if (gsr_mark_visible(P, means3D, viewmatrix, projmatrix, reinterpret_cast<unsigned char*>(present),
                       gsr_shim_get_stream()) != GSR_OK)
    raise("CudaRasterizer::Rasterizer::markVisible");
}

int Rasterizer::forward(ScratchAllocator geometryBuffer, ScratchAllocator binningBuffer, ScratchAllocator imageBuffer,
                        const int P, int D, int M, const float* background, const int width, int height,
                        const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
                        const float* scales, const float scale_modifier, const float* rotations,
                        const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                        const float* cam_pos, const float tan_fovx, float tan_fovy, const bool prefiltered,
                        float* out_color, float* out_depth, float* out_acc, int* radii, bool debug) {
  gsr_fwd_args a{};
  a.P = P, a.D = D, a.M = M, a.width = width, a.height = height;
  a.background = background, a.means3D = means3D, a.shs = shs, a.colors_precomp = colors_precomp;
  a.opacities = opacities, a.scales = scales, a.scale_modifier = scale_modifier, a.rotations = rotations;
  a.cov3D_precomp = cov3D_precomp, a.viewmatrix = viewmatrix, a.projmatrix = projmatrix, a.cam_pos = cam_pos;
  a.tan_fovx = tan_fovx, a.tan_fovy = tan_fovy, a.prefiltered = prefiltered ? 1 : 0;
  a.out_color = out_color, a.out_depth = out_depth, a.out_acc = out_acc, a.radii = radii, a.debug = debug ? 1 : 0;
  int rendered = 0;
  if (gsr_forward(&a, call_std_function, &geometryBuffer, call_std_function, &binningBuffer, call_std_function,
                  &imageBuffer, gsr_shim_get_stream(), &rendered) != GSR_OK)
    raise("CudaRasterizer::Rasterizer::forward");
  return rendered;
}

void Rasterizer::backward(const int P, int D, int M, int R, const float* background, const int width, int height,
                          const float* means3D, const float* shs, const float* colors_precomp, const float* scales,
                          const float scale_modifier, const float* rotations, const float* cov3D_precomp,
                          const float* viewmatrix, const float* projmatrix, const float* campos, const float tan_fovx,
                          float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer,
                          char* image_buffer, const float* dL_dpix, const float* dL_dacc, float* dL_dmean2D,
                          float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D,
                          float* dL_dsh, float* dL_dscale, float* dL_drot, bool debug) {
  gsr_bwd_args a{};
  a.P = P, a.D = D, a.M = M, a.R = R, a.width = width, a.height = height;
  a.background = background, a.means3D = means3D, a.shs = shs, a.colors_precomp = colors_precomp;
  a.scales = scales, a.scale_modifier = scale_modifier, a.rotations = rotations, a.cov3D_precomp = cov3D_precomp;
  a.viewmatrix = viewmatrix, a.projmatrix = projmatrix, a.cam_pos = campos;
  a.tan_fovx = tan_fovx, a.tan_fovy = tan_fovy, a.radii = radii;
  a.geom_buffer = geom_buffer, a.binning_buffer = binning_buffer, a.image_buffer = image_buffer;
  a.dL_dpix = dL_dpix, a.dL_dacc = dL_dacc, a.dL_dmean2D = dL_dmean2D, a.dL_dconic = dL_dconic;
  a.dL_dopacity = dL_dopacity, a.dL_dcolor = dL_dcolor, a.dL_dmean3D = dL_dmean3D, a.dL_dcov3D = dL_dcov3D;
  a.dL_dsh = dL_dsh, a.dL_dscale = dL_dscale, a.dL_drot = dL_drot, a.debug = debug ? 1 : 0;
  if (gsr_backward(&a, gsr_shim_get_stream()) != GSR_OK) raise("CudaRasterizer::Rasterizer::backward");
}

}  // namespace CudaRasterizer
