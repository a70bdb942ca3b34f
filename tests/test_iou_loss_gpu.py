"""Waymo `iou` head loss (row F2 helper) on the GPU: aligned rotated 3-D IoU kernel and the IouLoss assembled on the
fused head output, against the CPU oracle (pinned to the reference's CenterHead.loss by tests/test_predict_cpu.py)."""
import math

import numpy as np
import pytest
import torch

from oracle import pillarnext_oracle as O, predict_oracle as P
from pillarnext_b200 import loss as L, modules, ops, synth

pytestmark = pytest.mark.gpu
OFF = dict(reg=0, height=2, dim=3, rot=6, vel=8, iou=10, hm=11)


def test_aligned_iou3d_kernel_matches_oracle():
    rng = np.random.default_rng(4)
    n = 160
    a = np.concatenate([rng.uniform(-2, 2, (n, 2)), rng.uniform(-1, 1, (n, 1)), rng.uniform(0.5, 3.5, (n, 3)),
                        rng.uniform(-math.pi, math.pi, (n, 1))], 1).astype(np.float32)
    b = a + rng.normal(0, 0.4, a.shape).astype(np.float32)
    b[:, 3:6] = np.abs(b[:, 3:6]) + 0.2
    got = ops.aligned_iou3d(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu()
    want = torch.tensor([float(P.aligned_iou3d(x, y)) for x, y in zip(a, b)])
    assert (want > 0.1).sum() > 40
    assert (got - want).abs().max().item() < 3e-6


def test_iou_head_loss_value_and_gradient():
    B, H, W, M = 2, 12, 10, 24
    g = torch.Generator().manual_seed(2)
    pd = dict(reg=torch.rand(B, 2, H, W, generator=g), height=torch.randn(B, 1, H, W, generator=g),
              dim=torch.randn(B, 3, H, W, generator=g) * 0.3 + 0.6, rot=torch.randn(B, 2, H, W, generator=g),
              vel=torch.randn(B, 2, H, W, generator=g), iou=torch.randn(B, 1, H, W, generator=g) * 0.5,
              hm=torch.randn(B, 2, H, W, generator=g))
    osf, vs, pc = 4, [0.5, 0.5, 8.0], [-10.0, -12.0, -5.0, 10.0, 12.0, 3.0]
    ind = torch.randint(0, H * W, (B, M), generator=g)
    mask = (torch.rand(B, M, generator=g) > 0.4).to(torch.uint8)
    # reference box decode (centerhead.py:172-209) on CPU, gt = decoded box + noise so the IoUs are spread over (0, 1)
    bdim = torch.exp(torch.clamp(pd["dim"], -5, 5)).permute(0, 2, 3, 1)
    brot = torch.atan2(pd["rot"].permute(0, 2, 3, 1)[..., 0:1], pd["rot"].permute(0, 2, 3, 1)[..., 1:2])
    ys, xs = torch.meshgrid(torch.arange(0, H), torch.arange(0, W), indexing="ij")
    xs = (xs.view(1, H, W, 1).float() + pd["reg"].permute(0, 2, 3, 1)[..., 0:1]) * osf * vs[0] + pc[0]
    ys = (ys.view(1, H, W, 1).float() + pd["reg"].permute(0, 2, 3, 1)[..., 1:2]) * osf * vs[1] + pc[1]
    boxes = torch.cat([xs, ys, pd["height"].permute(0, 2, 3, 1), bdim, brot], 3).permute(0, 3, 1, 2).contiguous()
    rows = boxes.permute(0, 2, 3, 1).reshape(B, H * W, 7)
    gt = torch.gather(rows, 1, ind.unsqueeze(-1).expand(-1, -1, 7)) + torch.randn(B, M, 7, generator=g) * 0.25
    gt[..., 3:6] = gt[..., 3:6].abs() + 0.2
    iou_pred = pd["iou"].clone().requires_grad_(True)
    want = O.iou_loss(iou_pred, mask, ind, boxes, gt)
    want.backward()
    out = torch.zeros(B * H * W, 16)
    for k, o in OFF.items():
        v = pd[k].permute(0, 2, 3, 1).reshape(B * H * W, -1)
        out[:, o:o + v.shape[1]] = v
    out = out.cuda().requires_grad_(True)
    got = L.iou_head_loss(out, OFF, B, H, W, ind.cuda(), mask.cuda(), gt.cuda(), osf, vs, pc)
    got.backward()
    assert 0.05 < float(want) < 2.0
    assert abs(float(got) - float(want)) < 2e-6 * max(1.0, float(want))
    gcol = out.grad[:, OFF["iou"]].cpu().view(B, H, W)
    assert torch.allclose(gcol, iou_pred.grad[:, 0], atol=1e-6)
    other = out.grad.clone()
    other[:, OFF["iou"]] = 0
    assert other.abs().max().item() == 0.0                                  # the target carries no gradient


def test_waymo_style_model_trains_one_step():
    tasks = [["vehicle"], ["pedestrian", "cyclist"]]
    cfg = synth.tiny_config(64, tasks)
    cfg["common_heads"] = dict(reg=(2, 2), height=(1, 2), dim=(3, 2), rot=(2, 2), vel=(2, 2), iou=(1, 2))
    torch.manual_seed(0)
    model = modules.build_pillarnext_b(cfg).cuda().train()
    assert model.head.with_iou
    ex = synth.make_batch([5, 6], 2000, cfg, n_boxes=12, sweeps=1)
    exg = {k: ([e.cuda() for e in v] if isinstance(v, list) and torch.is_tensor(v[0]) else (v.cuda() if torch.is_tensor(v) else v))
           for k, v in ex.items()}
    loss, logs = model(exg)
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss).item()
    for t in range(2):
        assert "iou_loss" in logs[t] and 0.0 <= float(logs[t]["iou_loss"]) < 2.5
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    w = model.head.tasks[0].iou[3].weight.grad
    assert w.abs().sum().item() > 0                                          # the iou head receives a gradient
