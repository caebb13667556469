// TEST / MEASUREMENT INFRASTRUCTURE — ours, no reference code.
//
// A persistent "mapping-thread" session over the reference's UNCHANGED operator classes
// (GaussianRasterizationSettings, GaussianRasterizer, _RasterizeGaussians: include/gs/gs/rasterizer.cuh,
// src/gs/rasterizer.cu, src/gs/rasterize_points.cu). One step does what one view of GS-LIVM's mapping iteration
// does around the hot path (include/gs/gs/render_utils.cuh:13-56, src/liw/lioOptimization.cpp:1693-1736,1822-1828):
// settings from the camera, zeros_like(means2D), torch::cuda::synchronize(), GaussianRasterizer::forward, an image
// loss, loss.backward(), loss.item(). The leaf tensors live across steps like GaussianModel's parameters do.
//
// Included by oracle/dropin_binding.cpp (operator files linked to OUR libgsr_b200.so) and oracle/refop_binding.cpp
// (the same operator files linked to the reference's own kernels), so that bench.py's `e2e_dropin` compares exactly
// what a GS-LIVM user gets by re-linking: same host code, same inputs, only the library behind
// CudaRasterizer::Rasterizer differs.
#pragma once

#include <torch/extension.h>

#include <vector>

#include "gs/rasterizer.cuh"

namespace gsr_session {

class OperatorSession {
 public:
  OperatorSession(torch::Tensor means3D, torch::Tensor opacity, torch::Tensor shs, torch::Tensor scales,
                  torch::Tensor rotations, torch::Tensor bg, double tanfovx, double tanfovy, int64_t height,
                  int64_t width, int64_t sh_degree)
      : bg_(bg), tanfovx_((float)tanfovx), tanfovy_((float)tanfovy), height_((int)height), width_((int)width),
        sh_degree_((int)sh_degree) {
    means3D_ = means3D.detach().clone().requires_grad_(true);
    opacity_ = opacity.detach().clone().requires_grad_(true);
    shs_ = shs.detach().clone().requires_grad_(true);
    scales_ = scales.detach().clone().requires_grad_(true);
    rotations_ = rotations.detach().clone().requires_grad_(true);
  }

  // camera = 35 floats {viewmatrix 16, projmatrix 16, camera centre 3} and target [3,H,W]: HOST (pinned) tensors,
  // copied to the device inside the step. Returns the loss (device-to-host read: the step's synchronisation).
  double step(torch::Tensor camera_host, torch::Tensor target_host) {
    auto camera = camera_host.to(torch::kCUDA, /*non_blocking=*/true);
    auto target = target_host.to(torch::kCUDA, /*non_blocking=*/true);
    GaussianRasterizationSettings settings = {.image_height = height_,
                                              .image_width = width_,
                                              .tanfovx = tanfovx_,
                                              .tanfovy = tanfovy_,
                                              .bg = bg_,
                                              .scale_modifier = 1.0f,
                                              .viewmatrix = camera.slice(0, 0, 16).view({4, 4}),
                                              .projmatrix = camera.slice(0, 16, 32).view({4, 4}),
                                              .sh_degree = sh_degree_,
                                              .camera_center = camera.slice(0, 32, 35),
                                              .prefiltered = false};
    GaussianRasterizer rasterizer(settings);
    for (auto* t : {&means3D_, &opacity_, &shs_, &scales_, &rotations_}) t->mutable_grad() = torch::Tensor();
    auto means2D = torch::zeros_like(means3D_).requires_grad_(true);  // render_utils.cuh:39-40
    torch::cuda::synchronize();                                       // render_utils.cuh:49
    auto [color, radii, depth, acc] =
        rasterizer.forward(means3D_, means2D, opacity_, shs_, torch::Tensor(), scales_, rotations_, torch::Tensor());
    auto loss = (color - target).abs().mean();  // the L1 term of the mapping loss (lioOptimization.cpp:1729-1734)
    loss.backward();
    last_color_ = color.detach();
    return loss.item<double>();
  }

  std::vector<torch::Tensor> grads() const {
    return {means3D_.grad(), shs_.grad(), opacity_.grad(), scales_.grad(), rotations_.grad()};
  }
  torch::Tensor last_color() const { return last_color_; }

 private:
  torch::Tensor means3D_, opacity_, shs_, scales_, rotations_, bg_, last_color_;
  float tanfovx_, tanfovy_;
  int height_, width_, sh_degree_;
};

inline void bind(pybind11::module_& m) {
  namespace py = pybind11;
  py::class_<OperatorSession>(m, "OperatorSession")
      .def(py::init<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, double,
                    double, int64_t, int64_t, int64_t>())
      // the LibTorch autograd engine must not be entered with the Python GIL held
      .def("step", &OperatorSession::step, py::call_guard<py::gil_scoped_release>())
      .def("grads", &OperatorSession::grads)
      .def("last_color", &OperatorSession::last_color);
}

}  // namespace gsr_session
