"""Oracle pinning for row F1 (decode + rotated NMS), CPU only.

* oracle.predict_oracle.iou_bev  vs  the reference's own iou3d_cpu.cpp compiled into oracle/_ref (build container only;
  on the GPU box the prebuilt oracle/_ref/*.so travels with the snapshot, otherwise the test is skipped);
* oracle.predict_oracle.predict  vs  the reference's own CenterHead.predict run on CPU with its CUDA-only nms_gpu call
  replaced by the oracle's greedy rotated NMS (needs /root/reference: build container only);
* closed-form IoU cases and NMS invariants that need neither.
"""
import math
import types

import numpy as np
import pytest
import torch

from oracle import build_ref_iou, predict_oracle as P, reference_loader


def rand_boxes(n, seed, spread=6.0):
    g = torch.Generator().manual_seed(seed)
    xy = (torch.rand(n, 2, generator=g) - 0.5) * spread
    z = torch.rand(n, 1, generator=g)
    dims = torch.rand(n, 3, generator=g) * 3 + 0.5
    yaw = (torch.rand(n, 1, generator=g) - 0.5) * 2 * math.pi
    return torch.cat([xy, z, dims, yaw], 1).float()


def test_iou_closed_form():
    a = np.array([0, 0, 0, 4, 2, 1, 0], dtype=np.float32)
    b = np.array([1, 0, 0, 4, 2, 1, 0], dtype=np.float32)
    assert abs(float(P.iou_bev(a, b)) - 0.6) < 1e-6                       # 6 / (8 + 8 - 6)
    assert abs(float(P.iou_bev(a, a)) - 1.0) < 1e-6
    c = np.array([10, 10, 0, 1, 1, 1, 0.3], dtype=np.float32)
    assert float(P.iou_bev(a, c)) == 0.0
    sq = np.array([0, 0, 0, 2, 2, 1, 0], dtype=np.float32)
    sq45 = np.array([0, 0, 0, 2, 2, 1, math.pi / 4], dtype=np.float32)
    inter = 8 * (math.sqrt(2) - 1)                                         # regular octagon
    assert abs(float(P.iou_bev(sq, sq45)) - inter / (8 - inter)) < 1e-4


def test_iou_matches_reference_extension():
    ext = build_ref_iou.load(build=True)
    if ext is None:
        pytest.skip("reference iou3d_cpu.cpp neither compiled (oracle/_ref) nor present")
    a, b = rand_boxes(24, 1), rand_boxes(24, 2)
    want = torch.zeros(24, 24)
    ext.boxes_iou_bev_cpu(a.contiguous(), b.contiguous(), want)
    got = torch.tensor([[float(P.iou_bev(x.numpy(), y.numpy())) for y in b] for x in a])
    assert (want > 0.05).sum() > 20                                        # the sample does overlap
    assert (got - want).abs().max().item() < 2e-5, (got - want).abs().max().item()


def test_nms_invariants():
    boxes = rand_boxes(60, 5, spread=5.0).numpy()
    keep = P.nms_rotated(boxes, 0.2)
    assert keep[0] == 0 and keep == sorted(keep)
    for i, a in enumerate(keep):                                            # kept boxes do not suppress each other
        for b in keep[i + 1:]:
            assert float(P.iou_bev(boxes[a], boxes[b])) <= 0.2
    for j in set(range(60)) - set(keep):                                    # every dropped box has a kept suppressor before it
        assert any(k < j and float(P.iou_bev(boxes[k], boxes[j])) > 0.2 for k in keep)
    assert P.nms_rotated(boxes, 1.1) == list(range(60))


def _test_cfg():
    return dict(post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], score_threshold=0.1,
                out_size_factor=[4, 4], voxel_size=[0.075, 0.075], pc_range=[-50.4, -50.4],
                nms=dict(nms_iou_threshold=[[0.2], [0.2, 0.2]], nms_pre_max_size=1000, nms_post_max_size=83))


def _fake_preds(B, H, W, classes, seed):
    """Head outputs with a sparse set of confident, separated peaks (so the NMS input is small and not degenerate)."""
    g = torch.Generator().manual_seed(seed)
    pd = dict(reg=torch.rand(B, 2, H, W, generator=g), height=torch.randn(B, 1, H, W, generator=g),
              dim=torch.randn(B, 3, H, W, generator=g) * 0.3 + 0.5, rot=torch.randn(B, 2, H, W, generator=g),
              vel=torch.randn(B, 2, H, W, generator=g), hm=torch.full((B, classes, H, W), -6.0))
    n = 60
    for b in range(B):
        ys = torch.randint(0, H, (n,), generator=g)
        xs = torch.randint(0, W, (n,), generator=g)
        cs = torch.randint(0, classes, (n,), generator=g)
        pd["hm"][b, cs, ys, xs] = torch.randn(n, generator=g) * 1.5 + 0.5
    return pd


def test_predict_matches_reference_predict():
    if not reference_loader.available():
        pytest.skip("reference tree not present (build container only)")
    ref = reference_loader.load_reference()
    tasks = [["car"], ["truck", "construction_vehicle"]]
    head = ref.CenterHead(in_channels=256, tasks=tasks, weight=0.25, code_weights=[1.0] * 10,
                          common_heads=dict(reg=[2, 2], height=[1, 2], dim=[3, 2], rot=[2, 2], vel=[2, 2]),
                          strides=[2, 2], rectifier=[[0.0], [0.0, 0.0]])
    cfg = _test_cfg()
    preds = [_fake_preds(2, 24, 20, len(t), 10 + i) for i, t in enumerate(tasks)]
    ns = lambda d: types.SimpleNamespace(**{k: (ns(v) if isinstance(v, dict) else v) for k, v in d.items()})
    saved = ref.box_torch_ops.rotate_nms_pcdet
    ref.box_torch_ops.rotate_nms_pcdet = P.rotate_nms_pcdet                # nms_gpu is CUDA-only: the oracle NMS stands in
    try:
        want = head.predict(dict(token=["a", "b"]), [{k: v.clone() for k, v in p.items()} for p in preds], ns(cfg))
    finally:
        ref.box_torch_ops.rotate_nms_pcdet = saved
    got = P.predict(preds, [len(t) for t in tasks], cfg, [[0.0], [0.0, 0.0]], tokens=["a", "b"])
    assert len(got) == len(want) == 2
    for g, w in zip(got, want):
        assert g["token"] == w["token"]
        assert g["box3d_lidar"].shape == w["box3d_lidar"].shape and w["box3d_lidar"].shape[0] > 5
        assert torch.equal(g["label_preds"], w["label_preds"])
        assert torch.allclose(g["scores"], w["scores"], atol=1e-6)
        assert torch.allclose(g["box3d_lidar"], w["box3d_lidar"], atol=1e-5)
