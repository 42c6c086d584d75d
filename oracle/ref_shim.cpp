// TEST INFRASTRUCTURE — not product code.
// pybind11 shim that exposes the four hot-path operators of the UNMODIFIED
// reference (compiled from the sources where they lie under /root/reference,
// see oracle/build_ref.py) under the module name `kaolin_ref_C`.
// Mirrors the registration the reference does in kaolin/csrc/bindings.cpp:111-115.
#include <torch/extension.h>
#include "render/mesh/rasterization.h"
#include "render/mesh/dibr_soft_mask.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("packed_rasterize_forward_cuda", &kaolin::packed_rasterize_forward_cuda);
  m.def("rasterize_backward_cuda", &kaolin::rasterize_backward_cuda);
  m.def("dibr_soft_mask_forward_cuda", &kaolin::dibr_soft_mask_forward_cuda);
  m.def("dibr_soft_mask_backward_cuda", &kaolin::dibr_soft_mask_backward_cuda);
}
