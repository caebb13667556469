// TEST INFRASTRUCTURE — "Oracle A": Python binding around the UNMODIFIED reference rasterizer.
//
// This file is ours; it contains no reference code. The recipe in oracle/Makefile compiles the
// reference translation units where they lie under /root/reference
//   src/cuda_rasterizer/{forward,backward,rasterizer_impl}.cu   (nvcc, sm_100a)
//   src/gs/rasterize_points.cu                                  (host only)
// and links them with this binding into oracle/_ref/gsr_ref_ext*.so. Only tests/, bench.py's
// reference arm and __graft_entry__.smoke() may load the result; the product never does.
//
// Exposed (all tensors CUDA unless noted):
//   forward(...)   -> reference RasterizeGaussiansCUDA         (rasterize_points.cu:46-130)
//   backward(...)  -> reference RasterizeGaussiansBackwardCUDA (rasterize_points.cu:132-224)
//   mark_visible   -> reference markVisible                    (rasterize_points.cu:226-241)
//   geom_state / binning_state / image_state: carve the three opaque byte buffers with the
//   reference's own GeometryState/BinningState/ImageState::fromChunk (rasterizer_impl.cu:137-177)
//   so parity tests can compare every intermediate array.
#include <torch/extension.h>

#include <cuda_runtime_api.h>

#include <map>
#include <string>

#include "cuda_rasterizer/rasterizer_impl.h"
#include "gs/rasterize_points.cuh"

namespace {

using TensorMap = std::map<std::string, torch::Tensor>;

template <typename T>
torch::Tensor copy_out(const T* dev_ptr, std::vector<int64_t> shape, torch::ScalarType dt, const torch::Tensor& like) {
  auto out = torch::empty(shape, like.options().dtype(dt));
  if (out.numel() > 0) {
    cudaMemcpy(out.data_ptr(), dev_ptr, out.numel() * out.element_size(), cudaMemcpyDeviceToDevice);
  }
  return out;
}

TensorMap geom_state(const torch::Tensor& geomBuffer, int64_t P) {
  char* chunk = reinterpret_cast<char*>(geomBuffer.data_ptr());
  auto g = CudaRasterizer::GeometryState::fromChunk(chunk, (size_t)P);
  TensorMap m;
  m["depths"] = copy_out(g.depths, {P}, torch::kFloat32, geomBuffer);
  m["clamped"] = copy_out(g.clamped, {P, 3}, torch::kBool, geomBuffer);
  m["internal_radii"] = copy_out(g.internal_radii, {P}, torch::kInt32, geomBuffer);
  m["means2D"] = copy_out(reinterpret_cast<float*>(g.means2D), {P, 2}, torch::kFloat32, geomBuffer);
  m["cov3D"] = copy_out(g.cov3D, {P, 6}, torch::kFloat32, geomBuffer);
  m["conic_opacity"] = copy_out(reinterpret_cast<float*>(g.conic_opacity), {P, 4}, torch::kFloat32, geomBuffer);
  m["rgb"] = copy_out(g.rgb, {P, 3}, torch::kFloat32, geomBuffer);
  m["tiles_touched"] = copy_out(reinterpret_cast<int32_t*>(g.tiles_touched), {P}, torch::kInt32, geomBuffer);
  m["point_offsets"] = copy_out(reinterpret_cast<int32_t*>(g.point_offsets), {P}, torch::kInt32, geomBuffer);
  return m;
}

TensorMap binning_state(const torch::Tensor& binningBuffer, int64_t R) {
  char* chunk = reinterpret_cast<char*>(binningBuffer.data_ptr());
  auto b = CudaRasterizer::BinningState::fromChunk(chunk, (size_t)R);
  TensorMap m;
  m["point_list"] = copy_out(reinterpret_cast<int32_t*>(b.point_list), {R}, torch::kInt32, binningBuffer);
  m["point_list_unsorted"] =
      copy_out(reinterpret_cast<int32_t*>(b.point_list_unsorted), {R}, torch::kInt32, binningBuffer);
  m["point_list_keys"] = copy_out(reinterpret_cast<int64_t*>(b.point_list_keys), {R}, torch::kInt64, binningBuffer);
  m["point_list_keys_unsorted"] =
      copy_out(reinterpret_cast<int64_t*>(b.point_list_keys_unsorted), {R}, torch::kInt64, binningBuffer);
  return m;
}

TensorMap image_state(const torch::Tensor& imgBuffer, int64_t N, int64_t tiles) {
  char* chunk = reinterpret_cast<char*>(imgBuffer.data_ptr());
  auto s = CudaRasterizer::ImageState::fromChunk(chunk, (size_t)N);
  TensorMap m;
  m["accum_alpha"] = copy_out(s.accum_alpha, {N}, torch::kFloat32, imgBuffer);
  m["n_contrib"] = copy_out(reinterpret_cast<int32_t*>(s.n_contrib), {N}, torch::kInt32, imgBuffer);
  m["ranges"] = copy_out(reinterpret_cast<int32_t*>(s.ranges), {tiles, 2}, torch::kInt32, imgBuffer);
  return m;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "Oracle A: unmodified GS-LIVM reference rasterizer compiled for sm_100a (test infrastructure)";
  m.def("forward", &RasterizeGaussiansCUDA);
  m.def("backward", &RasterizeGaussiansBackwardCUDA);
  m.def("mark_visible", &markVisible);
  m.def("geom_state", &geom_state);
  m.def("binning_state", &binning_state);
  m.def("image_state", &image_state);
}
