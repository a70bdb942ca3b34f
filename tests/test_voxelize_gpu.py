"""V1-V3 parity: CUDA voxelizer (through the C-ABI) vs the oracle restatement of
pillar_encoder.py:78-125.  Index work must be BIT-EXACT."""
import numpy as np
import pytest
import torch

from oracle import pillarnext_oracle as O
from pillarnext_b200 import ops, synth

pytestmark = pytest.mark.gpu


def run_case(points, batch, cfg):
    v = ops.voxelize(points.cuda(), batch, cfg["voxel_size"], cfg["pc_range"])
    P, Nv = v.sync_counts()
    ref = O.voxelize(points, cfg["voxel_size"], cfg["pc_range"])
    assert P == ref["coords"].shape[0]
    assert Nv == int(ref["keep"].sum())
    coords = v.coords[:P].cpu()
    assert torch.equal(coords, ref["coords"]), "pillar coords differ (must be bit exact, same order)"
    pop = v.pillar_of_point[:points.shape[0]].cpu().long()
    if points.shape[0]:
        assert torch.equal(pop, ref["pillar_of_point"]), "unq_inv differs"
    # CSR buckets: ascending point ids inside each pillar, every kept point exactly once
    off = v.bucket_off[:P + 1].cpu().long()
    pts = v.bucket_pts[:Nv].cpu().long()
    assert off[0] == 0 and off[-1] == Nv
    if Nv:
        seg = torch.repeat_interleave(torch.arange(P), off[1:] - off[:-1])
        assert torch.equal(ref["pillar_of_point"][pts], seg)
        same = seg[1:] == seg[:-1]
        assert bool((pts[1:][same] > pts[:-1][same]).all())
        assert torch.equal(torch.sort(pts)[0], torch.nonzero(ref["keep"]).flatten())
    return v, ref


@pytest.mark.parametrize("n,batch,kind", [(30000, 1, "uniform"), (30000, 2, "lidar"), (1000, 3, "uniform"), (257, 1, "uniform")])
def test_voxelize_matches_oracle(n, batch, kind):
    cfg = synth.NUSC
    pts = synth.collate_points([synth.make_frame(s, n, cfg, kind, sweeps=10) for s in range(batch)])
    run_case(pts, batch, cfg)


def test_voxelize_edge_cases():
    cfg = synth.NUSC
    # empty
    run_case(torch.zeros(0, 6), 1, cfg)
    # all out of range / NaN / inf / exactly on the borders
    p = torch.zeros(8, 6)
    p[:, 1] = torch.tensor([-50.4, 50.4, 50.399998, -50.400002, float("nan"), float("inf"), 0.0, 1e9])
    p[:, 2] = torch.tensor([-50.4, 0.0, 50.399998, 0.0, 0.0, 0.0, float("nan"), 0.0])
    p[:, 3] = 100.0  # z is never range-checked (pillar_encoder.py:98-101)
    run_case(p, 1, cfg)
    # one hot pillar with many points + duplicates
    g = torch.Generator().manual_seed(0)
    q = torch.zeros(5000, 6)
    q[:, 1:3] = torch.rand(5000, 2, generator=g) * 0.07 + 1.0
    q[:, 3] = torch.rand(5000, generator=g)
    run_case(q, 1, cfg)


def test_voxelize_waymo_shape_and_tiny():
    cfg = synth.WAYMO_BENCH
    pts = synth.collate_points([synth.make_frame(7, 180000, cfg, "uniform")])
    run_case(pts, 1, cfg)
    cfg = synth.tiny_config(40)  # V not a multiple of 32
    pts = synth.collate_points([synth.make_frame(s, 500, cfg) for s in range(2)])
    run_case(pts, 2, cfg)


def test_pfn_forward_matches_oracle():
    cfg = synth.NUSC
    torch.manual_seed(0)
    sd = {
        "reader.pfn_layers.0.linear.weight": torch.randn(32, 10) * 0.3,
        "reader.pfn_layers.0.norm.weight": torch.rand(32) + 0.5, "reader.pfn_layers.0.norm.bias": torch.randn(32) * 0.1,
        "reader.pfn_layers.0.norm.running_mean": torch.zeros(32), "reader.pfn_layers.0.norm.running_var": torch.ones(32),
        "reader.pfn_layers.1.linear.weight": torch.randn(64, 64) * 0.1,
        "reader.pfn_layers.1.norm.weight": torch.rand(64) + 0.5, "reader.pfn_layers.1.norm.bias": torch.randn(64) * 0.1,
        "reader.pfn_layers.1.norm.running_mean": torch.zeros(64), "reader.pfn_layers.1.norm.running_var": torch.ones(64),
    }
    pts = synth.collate_points([synth.make_frame(s, 20000, cfg, "lidar", sweeps=10) for s in range(2)])
    for training in (True, False):
        st = {}
        feat_ref, coords_ref, _ = O.reader_forward(pts, sd, cfg["voxel_size"], cfg["pc_range"], train=training, stats=st)
        d = {k: v.clone().cuda() for k, v in sd.items()}
        v = ops.voxelize(pts.cuda(), 2, cfg["voxel_size"], cfg["pc_range"])
        bn = lambda i: tuple(d["reader.pfn_layers.%d.norm.%s" % (i, k)] for k in ("weight", "bias", "running_mean", "running_var"))
        out = ops.pfn_forward(v, d["reader.pfn_layers.0.linear.weight"], bn(0), d["reader.pfn_layers.1.linear.weight"], bn(1), training)
        P, _ = v.sync_counts()
        feat = out["feat"][:P].cpu()
        err = (feat - feat_ref).abs().max().item()
        assert err < 2e-4, "PFN feature max abs err %g (training=%s)" % (err, training)   # fp32, tolerance 2e-4 abs
        fb = out["feat_bf16"][:P].float().cpu()
        assert torch.equal(fb, feat.bfloat16().float())  # bf16 copy == round-to-nearest of the fp32 result
        if training:
            for i in (0, 1):
                for k in ("running_mean", "running_var"):
                    key = "reader.pfn_layers.%d.norm.%s" % (i, k)
                    e = (d[key].cpu() - st[key]).abs().max().item()
                    assert e < 1e-4, (key, e)
