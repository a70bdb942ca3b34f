"""End-to-end parity of the assembled CUDA model (det3d API, through the C-ABI) vs the fp32 oracle.

The product computes convolutions in bf16 (fp32 accumulate) with bf16 activations; the reference is fp32.
Each kernel is checked tightly on identical operands elsewhere (test_igemm_gpu / test_wgrad_gpu); here the
ASSEMBLY is checked, with tolerances that reflect bf16 rounding through ~35 layers (stated per check) but are
far below what any logic error (tap order, layout, mask, BN population) produces (O(1) relative)."""
import pytest
import torch

from oracle import pillarnext_oracle as O
from oracle.weights import randomize_state_dict
from pillarnext_b200 import modules, synth

pytestmark = pytest.mark.gpu

TASKS = [["car"], ["truck", "construction_vehicle"]]


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def build(cfg, seed=3):
    model = modules.build_pillarnext_b(cfg)
    sd = randomize_state_dict(model.state_dict(), seed)
    model.load_state_dict(sd, strict=True)
    return model.cuda(), sd


def test_reader_module_forward_backward():
    cfg = synth.NUSC
    model, sd = build(synth.tiny_config(64, TASKS))
    reader = modules.PillarFeatureNet(5, [64, 64], cfg["voxel_size"], cfg["pc_range"])
    rsd = {k[len("reader."):]: v for k, v in sd.items() if k.startswith("reader.")}
    reader.load_state_dict(rsd, strict=True)
    reader = reader.cuda().train()
    pts = synth.collate_points([synth.make_frame(s, 6000, cfg, "lidar", sweeps=10) for s in range(2)])
    reader.batch_size = 2
    feat, coords, grid = reader(pts.cuda())
    p = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items() if k.startswith("reader.")}
    fo, co, go = O.reader_forward(pts, p, cfg["voxel_size"], cfg["pc_range"], train=True)
    assert torch.equal(coords.cpu(), co) and list(grid) == list(go)
    assert (feat.cpu() - fo).abs().max().item() < 2e-4                      # fp32 path
    R = torch.randn(fo.shape, generator=torch.Generator().manual_seed(1))
    (fo * R).sum().backward()
    (feat * R.cuda()).sum().backward()
    got = {"reader." + k: v.grad for k, v in reader.named_parameters()}
    for k, v in got.items():
        e = rel(v, p[k].grad)
        assert e < 2e-3, (k, e)                                             # fp32 backward: summation order only
    # eval mode forward
    reader.eval()
    fe, _, _ = reader(pts.cuda())
    sd_e = {k: v.detach() for k, v in p.items()}
    sd_e.update({"reader." + k: v.cpu() for k, v in reader.state_dict().items() if "running" in k})
    foe, _, _ = O.reader_forward(pts, sd_e, cfg["voxel_size"], cfg["pc_range"], train=False)
    assert (fe.cpu() - foe).abs().max().item() < 2e-4


@pytest.mark.parametrize("grid,npts,kind", [(128, 3000, "uniform"), (256, 4000, "lidar")])
def test_detector_matches_oracle(grid, npts, kind):
    cfg = synth.tiny_config(grid, TASKS)
    model, sd = build(cfg)
    model.train()
    B = 2
    ex = synth.make_batch([0, 1], npts, cfg, kind=kind, n_boxes=25, sweeps=10)
    exg = {k: ([e.cuda() for e in v] if isinstance(v, list) and torch.is_tensor(v[0]) else (v.cuda() if torch.is_tensor(v) else v)) for k, v in ex.items()}
    # ---- oracle (fp32 CPU, autograd)
    p = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    st = {}
    feat_o, coords_o, grid_o = O.reader_forward(ex["points"], p, cfg["voxel_size"], cfg["pc_range"], train=True, stats=st)
    f4, c4, shp = O.sparse_resnet_gather(feat_o, coords_o, grid_o, B, p, cfg["strides"], stats=st)
    bb_o = O.densify(f4, c4, shp)
    neck_o = O.aspp_forward(bb_o, p, train=True, stats=st)
    preds_o = O.centerhead_forward(neck_o, p, cfg["tasks"], cfg["common_heads"], train=True, stats=st)
    loss_o, rets_o = O.center_loss(ex, preds_o, cfg["weight"], cfg["code_weights"], True, cfg["voxel_size"], cfg["pc_range"], cfg["out_size_factor"])
    loss_o.backward()
    # ---- product, stage by stage through the reference-shaped API
    model.reader.batch_size = B
    x = model.reader(exg["points"])
    assert torch.equal(x[1].cpu(), coords_o)
    assert (x[0].cpu() - feat_o).abs().max().item() < 2e-4
    bb = model.backbone(*x)
    assert tuple(bb.shape) == tuple(bb_o.shape)
    assert torch.equal((bb.float().abs().sum(1) > 0).cpu() | (bb_o.abs().sum(1) > 0), bb_o.abs().sum(1) > 0), "active set differs"
    e = rel(bb, bb_o)
    assert e < 3e-2, "backbone output rel-L2 %g" % e                        # 21 bf16 conv+BN layers
    nk = model.neck(bb)
    e = rel(nk, neck_o)
    assert e < 4e-2, "neck output rel-L2 %g" % e
    preds = model.head(nk)
    for t in range(len(cfg["tasks"])):
        assert list(preds[t].keys()) == list(preds_o[t].keys())
        for k in preds[t]:
            assert preds[t][k].dtype == torch.float32 and tuple(preds[t][k].shape) == tuple(preds_o[t][k].shape)
            e = rel(preds[t][k], preds_o[t][k])
            assert e < 6e-2, "head %d/%s rel-L2 %g" % (t, k, e)
    loss, rets = model.head.loss(exg, preds)
    assert abs(loss.item() - loss_o.item()) < 3e-2 * abs(loss_o.item()), (loss.item(), loss_o.item())
    loss.backward()
    worst = []
    for k, v in model.named_parameters():
        assert v.grad is not None, "no gradient for %s" % k
        assert torch.isfinite(v.grad).all(), k
        worst.append((rel(v.grad, p[k].grad), k))
    worst.sort(reverse=True)
    bad = [(e, k) for e, k in worst if e > 0.25]                            # bf16 gradients through ~35 layers
    assert not bad, "gradient rel-L2 errors too large: %s" % bad[:8]
    import statistics
    assert statistics.median(e for e, _ in worst) < 0.08, worst[:5]
    # running statistics were updated like the reference's BatchNorm (momentum / unbiased variance)
    msd = model.state_dict()
    for k, v in st.items():
        e = rel(msd[k], v)
        assert e < 5e-2, (k, e)
    # the same call through the detector entry point (single_stage.py:35-45)
    model.zero_grad()
    out = model(exg)
    assert isinstance(out, tuple) and out[0].dim() == 0 and len(out[1]) == len(cfg["tasks"])
