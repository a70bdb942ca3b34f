"""Freeze outputs of the reference's OWN CenterHead.predict (run on CPU in the build container) into
tests/golden/ref_predict.npz (TEST INFRASTRUCTURE).

The reference's predict calls the CUDA-only ``nms_gpu``; for this fixture that single call is replaced by the oracle's
greedy rotated NMS (oracle/predict_oracle.py, whose IoU is separately pinned to the reference's iou3d_cpu.cpp), so the
fixture pins decode, score/range filtering, per-class selection, truncation and the task merge of centerhead.py:231-384.
Inputs are regenerated from seeds by oracle/predict_fixtures.py; only the outputs are stored.

    python -m oracle.make_golden_predict
"""
import os
import types

import numpy as np

from oracle import predict_oracle as P, reference_loader
from oracle.predict_fixtures import fake_preds, test_cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TASKS = [["car"], ["truck", "construction_vehicle"]]
SHAPE = (2, 24, 20)
SEEDS = (10, 11)


def main():
    ref = reference_loader.load_reference()
    head = ref.CenterHead(in_channels=256, tasks=TASKS, weight=0.25, code_weights=[1.0] * 10,
                          common_heads=dict(reg=[2, 2], height=[1, 2], dim=[3, 2], rot=[2, 2], vel=[2, 2]),
                          strides=[2, 2], rectifier=[[0.0], [0.0, 0.0]])
    cfg = test_cfg()
    preds = [fake_preds(*SHAPE, len(t), SEEDS[i]) for i, t in enumerate(TASKS)]
    ns = lambda d: types.SimpleNamespace(**{k: (ns(v) if isinstance(v, dict) else v) for k, v in d.items()})
    saved = ref.box_torch_ops.rotate_nms_pcdet
    ref.box_torch_ops.rotate_nms_pcdet = P.rotate_nms_pcdet
    try:
        out = head.predict(dict(token=["a", "b"]), [{k: v.clone() for k, v in p.items()} for p in preds], ns(cfg))
    finally:
        ref.box_torch_ops.rotate_nms_pcdet = saved
    arrays = {}
    for i, o in enumerate(out):
        arrays["box3d_lidar_%d" % i] = o["box3d_lidar"].numpy()
        arrays["scores_%d" % i] = o["scores"].numpy()
        arrays["label_preds_%d" % i] = o["label_preds"].numpy()
    path = os.path.join(ROOT, "tests", "golden", "ref_predict.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
