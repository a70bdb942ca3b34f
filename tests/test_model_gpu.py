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


def surrogate_weights(t, k, shape):
    g = torch.Generator().manual_seed(1000 * t + sum(ord(c) for c in k))
    return torch.randn(tuple(shape), generator=g)


def run_oracle(cfg, sd, ex, B, quant):
    O.QUANT = quant
    try:
        p = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
        st = {}
        feat, coords, grid = O.reader_forward(ex["points"], p, cfg["voxel_size"], cfg["pc_range"], train=True, stats=st)
        f4, c4, shp = O.sparse_resnet_gather(feat, coords, grid, B, p, cfg["strides"], stats=st)
        bb = O.densify(f4, c4, shp)
        neck = O.aspp_forward(bb, p, train=True, stats=st)
        preds = O.centerhead_forward(neck, p, cfg["tasks"], cfg["common_heads"], train=True, stats=st)
        # (a) linear surrogate loss: d/dpred is the same fixed tensor on both sides, so parameter gradients
        #     compare the BACKWARD machinery without the sign discontinuities of the L1 / clamp terms
        sur = sum((v * surrogate_weights(t, k, v.shape)).sum() for t, pd in enumerate(preds) for k, v in pd.items())
        sur.backward(retain_graph=True)
        gsur = {k: v.grad.clone() for k, v in p.items() if v.grad is not None}
        for v in p.values():
            v.grad = None
        # (b) the real CenterPoint loss
        loss, rets = O.center_loss(ex, [dict(pd) for pd in preds], cfg["weight"], cfg["code_weights"], True, cfg["voxel_size"], cfg["pc_range"], cfg["out_size_factor"])
        loss.backward()
        return dict(p=p, st=st, feat=feat, coords=coords, bb=bb, neck=neck, preds=preds, loss=loss, gsur=gsur)
    finally:
        O.QUANT = False


@pytest.fixture
def deterministic():
    from pillarnext_b200 import functional as Fn
    prev = Fn.set_deterministic(True)
    yield
    Fn.set_deterministic(prev)


@pytest.mark.parametrize("grid,npts,kind", [(128, 3000, "uniform"), (256, 4000, "lidar")])
def test_detector_matches_oracle(grid, npts, kind, deterministic):
    cfg = synth.tiny_config(grid, TASKS)
    model, sd = build(cfg)
    model.train()
    B = 2
    ex = synth.make_batch([0, 1], npts, cfg, kind=kind, n_boxes=25, sweeps=10)
    exg = {k: ([e.cuda() for e in v] if isinstance(v, list) and torch.is_tensor(v[0]) else (v.cuda() if torch.is_tensor(v) else v)) for k, v in ex.items()}
    oq = run_oracle(cfg, sd, ex, B, True)     # bf16-faithful restatement: tight tolerances
    of = run_oracle(cfg, sd, ex, B, False)    # the reference's fp32 arithmetic: bf16-noise tolerances
    report = []

    def check(name, got, key, tol_q, tol_f, sub=None):
        a = oq[key] if sub is None else sub(oq[key])
        b = of[key] if sub is None else sub(of[key])
        eq, ef = rel(got, a), rel(got, b)
        report.append("%s: vs bf16-faithful %.2e (tol %.0e), vs fp32 %.2e (tol %.0e)" % (name, eq, tol_q, ef, tol_f))
        assert eq < tol_q and ef < tol_f, "\n".join(report)

    model.reader.batch_size = B
    x = model.reader(exg["points"])
    assert torch.equal(x[1].cpu(), of["coords"])
    assert (x[0].cpu() - of["feat"]).abs().max().item() < 2e-4
    bb = model.backbone(*x)
    assert tuple(bb.shape) == tuple(of["bb"].shape)
    act_o = of["bb"].abs().sum(1) > 0
    assert torch.equal((bb.float().abs().sum(1) > 0).cpu() | act_o, act_o), "active set differs"
    check("backbone", bb, "bb", 8e-2, 1.2e-1)
    nk = model.neck(bb)
    check("neck", nk, "neck", 1e-1, 1.5e-1)
    preds = model.head(nk)
    for t in range(len(cfg["tasks"])):
        assert list(preds[t].keys()) == list(of["preds"][t].keys())
        for k in preds[t]:
            assert preds[t][k].dtype == torch.float32 and tuple(preds[t][k].shape) == tuple(of["preds"][t][k].shape)
            check("head %d/%s" % (t, k), preds[t][k], "preds", 1.5e-1, 2e-1, sub=lambda d, t=t, k=k: d[t][k])
    sur = sum((v * surrogate_weights(t, k, v.shape).cuda()).sum() for t, pd in enumerate(preds) for k, v in pd.items())
    sur.backward(retain_graph=True)
    worst = []
    for k, v in model.named_parameters():
        assert v.grad is not None and torch.isfinite(v.grad).all(), k
        if k.endswith(".0.bias") and ("shared_conv" in k or ".tasks." in k):
            continue     # conv bias in front of a BatchNorm: exact gradient is zero, only rounding noise remains
        worst.append((rel(v.grad, oq["gsur"][k]), rel(v.grad, of["gsur"][k]), k))
    worst.sort(reverse=True)
    report += ["surrogate grad %s: vs bf16-faithful %.3f vs fp32 %.3f" % (k, a, b) for a, b, k in worst[:10]]
    import statistics
    med = statistics.median(a for a, _, _ in worst)
    report.append("median surrogate-grad rel-L2 vs bf16-faithful: %.4f" % med)
    # Gradients of a 35-layer ReLU network are DISCONTINUOUS in the activations: the 3-10 % forward difference
    # between two bf16 evaluations flips the ReLU gate of ~1-4 % of the units per layer, each flip moving the
    # gradient by a finite amount (the two ORACLE variants differ from each other by the same amount).  The exact
    # backward arithmetic is pinned op by op in tests/test_functional_gpu.py on identical inputs; here only
    # direction and magnitude are checked.
    cs = []
    for k, v in model.named_parameters():
        go = oq["gsur"].get(k)
        if go is not None and go.norm() > 1e-3 and not (k.endswith(".0.bias") and ("shared_conv" in k or ".tasks." in k)):
            cs.append((torch.nn.functional.cosine_similarity(v.grad.flatten().cpu(), go.flatten(), dim=0).item(),
                       (v.grad.norm() / go.norm()).item(), k))
    cs.sort()
    report += ["surrogate cos %.3f norm-ratio %.3f %s" % c for c in cs[:6]]
    print("\n".join(report[-8:]))
    # Runs under functional.set_deterministic(True) (fixture): ordered split-K + fp64 statistics make the bf16 run
    # bit-reproducible, so these are fixed numbers, not draws -- the assertions are the tight ones again (an earlier
    # revision had to tolerate one stray parameter because a different set of ReLU gates flipped from run to run).
    assert statistics.median(c for c, _, _ in cs) > 0.85 and cs[0][0] > 0.5, "\n".join(report)
    assert all(0.7 < r < 1.4 for _, r, _ in cs), "\n".join(report)
    model.zero_grad()
    loss, rets = model.head.loss(exg, [dict(pd) for pd in preds])
    report.append("loss %.6f  bf16-faithful %.6f  fp32 %.6f" % (loss.item(), oq["loss"].item(), of["loss"].item()))
    print("\n".join(report))
    assert abs(loss.item() - oq["loss"].item()) < 5e-2 * abs(oq["loss"].item()), "\n".join(report)
    assert abs(loss.item() - of["loss"].item()) < 5e-2 * abs(of["loss"].item()), "\n".join(report)
    loss.backward()
    cos = []
    for k, v in model.named_parameters():
        assert v.grad is not None and torch.isfinite(v.grad).all(), k
        go = oq["p"][k].grad
        if go.norm() > 1e-4:
            cos.append((torch.nn.functional.cosine_similarity(v.grad.flatten().cpu(), go.flatten(), dim=0).item(), k))
    cos.sort()
    # the real loss has sign discontinuities (L1, clamp): direction agreement only
    assert statistics.median(c for c, _ in cos) > 0.6, cos[:10]
    msd = model.state_dict()
    for k, v in of["st"].items():
        e = rel(msd[k], v)
        assert e < 5e-2, (k, e)
    model.zero_grad()
    out = model(exg)
    assert isinstance(out, tuple) and out[0].dim() == 0 and len(out[1]) == len(cfg["tasks"])


def test_full_size_frame_independence_and_point_order():
    """BASELINE.json configs[1] size (1344^2 pillar grid, 30k points per frame), size-independent properties:
    (1) in eval mode a frame's head outputs do not depend on which other frames share the batch -- bit-exact, because
        sites/pixels are ordered frame-major and every kernel computes a row from that row's neighbours only;
    (2) shuffling the points of a frame leaves the pillar set and its order unchanged (bit-exact) and the pillar
        features equal up to the fp32 summation order of the mean."""
    cfg = synth.NUSC
    model, _ = build(cfg, seed=5)
    model.eval()
    frames = [synth.make_frame(100 + s, 30000, cfg, "lidar", sweeps=10) for s in range(2)]

    def run(fr):
        ex = {"points": synth.collate_points(fr).cuda(), "token": ["t%d" % i for i in range(len(fr))]}
        with torch.no_grad():
            return model._forward(ex)

    both, solo = run(frames), run(frames[:1])
    assert len(both) == len(solo) == len(cfg["tasks"])
    for pb, ps in zip(both, solo):
        assert list(pb.keys()) == list(ps.keys())
        for k in pb:
            a, b = pb[k][0:1], ps[k]
            assert a.shape == b.shape and torch.isfinite(a).all(), k
            assert torch.equal(a, b), (k, (a - b).abs().max().item())

    pts = synth.collate_points(frames[:1])
    perm = torch.randperm(pts.shape[0], generator=torch.Generator().manual_seed(9))
    model.reader.batch_size = 1
    with torch.no_grad():
        f0, c0, _ = model.reader(pts.cuda())
        f1, c1, _ = model.reader(pts[perm].cuda())
    assert torch.equal(c0, c1)
    assert (f0 - f1).abs().max().item() < 1e-4


def test_packed_weight_cache_follows_fused_optimizer_steps():
    """torch's fused AdamW updates parameters without bumping `_version`: the packed bf16 weight cache must still be
    refreshed after every optimizer step (it is keyed on an optimizer-step generation as well), otherwise the forward
    keeps using the initial weights for the whole run."""
    from pillarnext_b200 import functional as Fn
    w = torch.nn.Parameter(torch.randn(64, 64, 3, 3, device="cuda") * 0.05)
    lay = Fn.WLayout("dense")
    p0 = Fn.packed(w, lay, "fwd")
    assert Fn.packed(w, lay, "fwd") is p0                                   # cached while nothing changed
    p0 = p0.clone()                                                          # stale operands are rewritten IN PLACE (pack.cu)
    opt = torch.optim.AdamW([w], lr=1e-1, fused=True)
    w.grad = torch.ones_like(w)
    opt.step()
    p1 = Fn.packed(w, lay, "fwd")
    assert not torch.equal(p1, p0)
    assert torch.equal(p1, lay.pack_fwd(w.detach()))
    with torch.no_grad():
        w.mul_(2.0)                                                          # plain in-place edit: version counter
    assert torch.equal(Fn.packed(w, lay, "fwd"), lay.pack_fwd(w.detach()))
