// TEST INFRASTRUCTURE — link-level proof of the drop-in boundary.
//
// This file is ours. oracle/Makefile (`make dropin`) compiles the reference's UNCHANGED operator files
//   /root/reference/src/gs/rasterize_points.cu   (torch::Tensor <-> raw pointers, scratch allocators)
//   /root/reference/src/gs/rasterizer.cu         (GaussianRasterizationSettings, _RasterizeGaussians, GaussianRasterizer)
// with -I<repo>/include FIRST, so "cuda_rasterizer/config.h" and "cuda_rasterizer/rasterizer.h" resolve to OUR
// headers, and links them against OUR libgsr_b200.so (CudaRasterizer::Rasterizer symbols from
// gs-livm_b200/csrc/rasterizer_shim.cpp). The function below drives the reference's own C++ classes exactly the
// way include/gs/gs/render_utils.cuh:13-56 does, then back-propagates through LibTorch autograd.
#include <torch/extension.h>

#include <vector>

#include "gs/rasterizer.cuh"
#include "operator_session.h"

namespace {

std::vector<torch::Tensor> render_and_backward(torch::Tensor means3D, torch::Tensor opacity, torch::Tensor shs,
                                               torch::Tensor scales, torch::Tensor rotations, torch::Tensor bg,
                                               torch::Tensor viewmatrix, torch::Tensor projmatrix,
                                               torch::Tensor campos, double tanfovx, double tanfovy, int64_t height,
                                               int64_t width, int64_t sh_degree, torch::Tensor dL_dcolor,
                                               torch::Tensor dL_dacc) {
  GaussianRasterizationSettings settings = {
      .image_height = static_cast<int>(height),
      .image_width = static_cast<int>(width),
      .tanfovx = static_cast<float>(tanfovx),
      .tanfovy = static_cast<float>(tanfovy),
      .bg = bg,
      .scale_modifier = 1.0f,
      .viewmatrix = viewmatrix,
      .projmatrix = projmatrix,
      .sh_degree = static_cast<int>(sh_degree),
      .camera_center = campos,
      .prefiltered = false};
  GaussianRasterizer rasterizer(settings);

  auto m = means3D.detach().clone().requires_grad_(true);
  auto o = opacity.detach().clone().requires_grad_(true);
  auto s = shs.detach().clone().requires_grad_(true);
  auto sc = scales.detach().clone().requires_grad_(true);
  auto r = rotations.detach().clone().requires_grad_(true);
  auto means2D = torch::zeros_like(m).requires_grad_(true);

  auto [color, radii, depth, acc] =
      rasterizer.forward(m, means2D, o, s, torch::Tensor(), sc, r, torch::Tensor());
  auto loss = (color * dL_dcolor).sum() + (acc * dL_dacc).sum();
  loss.backward();
  return {color.detach(), radii, depth.detach(), acc.detach(), m.grad(), s.grad(), o.grad(), sc.grad(), r.grad()};
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "GS-LIVM's unchanged operator files linked against libgsr_b200.so (test infrastructure)";
  // the LibTorch autograd engine must not be entered with the Python GIL held
  m.def("render_and_backward", &render_and_backward, pybind11::call_guard<pybind11::gil_scoped_release>());
  gsr_session::bind(m);  // persistent session for bench.py's e2e_dropin (oracle/operator_session.h)
}
