// pybind shim around the reference's own CPU rotated-IoU source (compiled from where it lies under /root/reference by
// oracle/build_ref_iou.py; TEST INFRASTRUCTURE, never linked into the product).
#include <torch/extension.h>

int boxes_iou_bev_cpu(at::Tensor boxes_a_tensor, at::Tensor boxes_b_tensor, at::Tensor ans_iou_tensor);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("boxes_iou_bev_cpu", &boxes_iou_bev_cpu, "rotated BEV IoU, reference iou3d_cpu.cpp");
}
