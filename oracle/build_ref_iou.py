"""Compile the reference's own CPU rotated-IoU source (det3d/core/iou3d_nms/src/iou3d_cpu.cpp) into oracle/_ref/
(TEST INFRASTRUCTURE).  Sources are compiled where they lie under /root/reference (nothing is copied); the recipe is
torch.utils.cpp_extension (g++ + ninja), not the reference's setup.py.  Only possible in the build container."""
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.path.join(os.environ.get("PNX_REFERENCE_ROOT", "/root/reference"), "det3d/core/iou3d_nms/src/iou3d_cpu.cpp")
OUT = os.path.join(ROOT, "_ref")
NAME = "ref_iou3d_cpu"


def available():
    return os.path.isfile(REF_SRC)


def load(build=True):
    """Returns the extension module (boxes_iou_bev_cpu) or None when neither the source nor a prebuilt .so exists."""
    so = os.path.join(OUT, NAME + ".so")
    if os.path.isfile(so) and not (build and available() and os.path.getmtime(so) < os.path.getmtime(REF_SRC)):
        import importlib.util
        import torch  # noqa: F401  (libtorch must be loaded first)
        spec = importlib.util.spec_from_file_location(NAME, so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    if not (build and available()):
        return None
    from torch.utils.cpp_extension import load as cpp_load
    os.makedirs(OUT, exist_ok=True)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    return cpp_load(name=NAME, sources=[REF_SRC, os.path.join(ROOT, "ref_iou_binding.cpp")], build_directory=OUT,
                    extra_include_paths=[cuda_inc, os.path.dirname(REF_SRC)], extra_cflags=["-O2", "-w"], verbose=False)


if __name__ == "__main__":
    m = load()
    print("built" if m is not None else "reference source not present", file=sys.stderr)
