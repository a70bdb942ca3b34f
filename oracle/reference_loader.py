"""Import the reference's OWN hot-path files from /root/reference on CPU (TEST INFRASTRUCTURE).

Only usable in the build container (the GPU box has no /root/reference); everything that
must travel is frozen by ``oracle/make_golden.py`` into ``tests/golden/``.

Shims installed while importing:
  * ``torch_scatter``  -> oracle/torch_scatter_shim.py (package absent, source not in tree)
  * ``det3d.core.iou3d_nms.iou3d_nms_cuda`` -> stub object (native ext; only needed by the
    Waymo IouLoss branch and eval NMS, neither used by the pinned nuScenes path)
The reference's ``det3d`` package name collides with this repo's drop-in ``det3d`` package,
so ``sys.modules`` is swapped during the import and restored afterwards.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("PNX_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "det3d/models/readers/pillar_encoder.py"))


_cache = None


def load_reference():
    """Returns a namespace with the reference classes: PillarFeatureNet, ASPPNeck, CenterHead,
    FastFocalLoss, RegLoss, IouRegLoss, AssignLabel, collate."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    from oracle import torch_scatter_shim

    saved = {k: v for k, v in sys.modules.items() if k == "det3d" or k.startswith("det3d.")}
    for k in saved:
        del sys.modules[k]
    had_ts = sys.modules.get("torch_scatter")
    sys.modules["torch_scatter"] = torch_scatter_shim
    stub = types.ModuleType("det3d.core.iou3d_nms.iou3d_nms_cuda")
    sys.path.insert(0, REF_ROOT)
    try:
        # stub must be in place before det3d.core.iou3d_nms/__init__ imports it
        importlib.import_module("det3d")
        sys.modules["det3d.core.iou3d_nms.iou3d_nms_cuda"] = stub
        ns = types.SimpleNamespace()
        pe = importlib.import_module("det3d.models.readers.pillar_encoder")
        ns.PillarFeatureNet, ns.PFNLayer, ns.PillarNet = pe.PillarFeatureNet, pe.PFNLayer, pe.PillarNet
        ns.ASPPNeck = importlib.import_module("det3d.models.necks.aspp").ASPPNeck
        ch = importlib.import_module("det3d.models.heads.centerhead")
        ns.CenterHead, ns.SepHead = ch.CenterHead, ch.SepHead
        ns.box_torch_ops = ch.box_torch_ops      # module object whose rotate_nms_pcdet the predict pin replaces
        cl = importlib.import_module("det3d.models.loss.centerloss")
        ns.FastFocalLoss, ns.RegLoss, ns.IouRegLoss = cl.FastFocalLoss, cl.RegLoss, cl.IouRegLoss
        ns.bbox3d_overlaps_diou = cl.bbox3d_overlaps_diou
        ns.centerloss = cl                       # module object (its boxes_aligned_iou3d_gpu is CUDA-only: the pin swaps it)
        ns.AssignLabel = importlib.import_module("det3d.datasets.pipelines.assign").AssignLabel
        ns.collate = importlib.import_module("det3d.datasets.loader.collate").collate
    finally:
        sys.path.remove(REF_ROOT)
        for k in [k for k in sys.modules if k == "det3d" or k.startswith("det3d.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        if had_ts is None:
            sys.modules.pop("torch_scatter", None)
        else:
            sys.modules["torch_scatter"] = had_ts
    _cache = ns
    return ns
