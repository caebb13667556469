// TEST / MEASUREMENT INFRASTRUCTURE — ours, no reference code.
//
// oracle/Makefile (`make refop`) compiles the reference's UNCHANGED operator files src/gs/rasterize_points.cu and
// src/gs/rasterizer.cu against the REFERENCE's own headers and links them to the reference's own kernels
// (src/cuda_rasterizer/*.cu, the objects of `make ref`): the stock GS-LIVM operator, end to end. Together with
// gsr_dropin_ext.so (the same operator files on OUR library) it gives bench.py's `e2e_dropin` its two sides.
#include <torch/extension.h>

#include "operator_session.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "GS-LIVM's unchanged operator on GS-LIVM's own rasterizer kernels (measurement infrastructure)";
  gsr_session::bind(m);
}
