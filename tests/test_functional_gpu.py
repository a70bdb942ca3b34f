"""Op-level autograd parity: each differentiable building block (forward AND backward, incl. the weight
packing / tap flipping / transposed-rulebook logic in pillarnext_b200/functional.py) vs torch autograd of the
equivalent dense fp32 op on the SAME bf16-rounded inputs.  Tolerance 1.5e-2 relative-L2 (bf16 rounding of the
stored outputs / gradient rows; fp32 accumulation-order noise)."""
import pytest
import torch
import torch.nn.functional as F

from pillarnext_b200 import functional as Fn
from pillarnext_b200 import modules, ops, synth

pytestmark = pytest.mark.gpu
TOL = 1.5e-2


def rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def rows_to_canvas(rows, coords_buv, B, H, W):
    """rows [n, C] at sites (b, u=x, v=y) -> NCHW canvas [B, C, H(y), W(x)] (differentiable)."""
    C = rows.shape[1]
    cv = torch.zeros(B, H, W, C, device=rows.device, dtype=rows.dtype)
    c = coords_buv.long()
    cv = cv.index_put((c[:, 0], c[:, 2], c[:, 1]), rows)
    return cv.permute(0, 3, 1, 2)


def canvas_to_rows(cv, coords_buv):
    c = coords_buv.long()
    return cv.permute(0, 2, 3, 1)[c[:, 0], c[:, 2], c[:, 1]]


@pytest.fixture(scope="module")
def pyramid():
    cfg = synth.tiny_config(96)
    pts = synth.collate_points([synth.make_frame(s, 1500, cfg, "uniform") for s in range(2)]).cuda()
    vox = ops.voxelize(pts, 2, cfg["voxel_size"], cfg["pc_range"])
    pyr = modules.build_pyramid(vox, [1, 2, 2, 2])
    return pyr


@pytest.mark.parametrize("stage,cin,cout", [(0, 64, 64), (1, 64, 128), (2, 128, 256), (3, 256, 256)])
def test_sparse_entry_conv_and_subm(pyramid, stage, cin, cout):
    pyr = pyramid
    src, dst = pyr.levels[stage], pyr.levels[stage + 1]
    stride = 1 if stage == 0 else 2
    torch.manual_seed(stage)
    B = src.batch
    lay = Fn.WLayout("sp")
    for kind in ("entry", "subm"):
        if kind == "entry":
            spec, lin, c_in, c_out, st = pyr.entry[stage], src, cin, cout, stride
        else:
            spec, lin, c_in, c_out, st = pyr.subm[stage], dst, cout, cout, 1
        x = torch.randn(lin.n, c_in, device="cuda").bfloat16().requires_grad_()
        w = (torch.randn(c_out, 3, 3, c_in, device="cuda") * 0.05).bfloat16().float().requires_grad_()
        out, stats = Fn.conv(x, w, None, spec, lay, want_stats=True)
        dy = torch.randn(dst.n, c_out, device="cuda").bfloat16()
        out.backward(dy)
        # torch reference: dense conv on the zero-filled canvas, outputs read at the output sites
        xr = x.detach().float().requires_grad_()
        wr = w.detach().clone().requires_grad_()
        cv = rows_to_canvas(xr, lin.coords[:lin.n], B, lin.V, lin.U)
        yc = F.conv2d(cv, wr.permute(0, 3, 1, 2), stride=st, padding=1)
        yr = canvas_to_rows(yc, dst.coords[:dst.n])
        yr.backward(dy.float())
        assert rel(out, yr) < TOL, (kind, "fwd", rel(out, yr))
        assert rel(x.grad, xr.grad) < TOL, (kind, "dgrad", rel(x.grad, xr.grad))
        assert rel(w.grad, wr.grad) < TOL, (kind, "wgrad", rel(w.grad, wr.grad))
        s = out.detach().double()
        assert torch.allclose(stats[:c_out], s.sum(0), rtol=1e-5, atol=1e-2)


@pytest.mark.parametrize("cin,cout,k,dil,bias,fp32", [(256, 256, 3, 1, False, False), (256, 64, 3, 1, True, False), (64, 384, 3, 1, True, False),
                                                      (384, 16, 3, 1, True, True), (1536, 256, 1, 1, False, False), (256, 256, 3, 6, False, False)])
def test_dense_conv_fn(cin, cout, k, dil, bias, fp32):
    torch.manual_seed(cin + cout)
    B, H, W = 2, 20, 24
    M = B * H * W
    x = torch.randn(M, cin, device="cuda").bfloat16().requires_grad_()
    w = (torch.randn(cout, cin, k, k, device="cuda") * 0.05).bfloat16().float().requires_grad_()
    b = torch.randn(cout, device="cuda").requires_grad_() if bias else None
    spec = Fn.dense_spec(B, H, W, k, dil)
    out, _ = Fn.conv(x, w, b, spec, Fn.WLayout("dense"), out_fp32=fp32)
    dy = torch.randn(M, cout, device="cuda")
    dy = dy if fp32 else dy.bfloat16()
    out.backward(dy)
    xr = x.detach().float().requires_grad_()
    wr = w.detach().clone().requires_grad_()
    br = b.detach().clone().requires_grad_() if bias else None
    yc = F.conv2d(xr.view(B, H, W, cin).permute(0, 3, 1, 2), wr, br, padding=dil * (k // 2), dilation=dil)
    yr = yc.permute(0, 2, 3, 1).reshape(M, cout)
    yr.backward(dy.float().bfloat16().float() if fp32 else dy.float())   # the product rounds gradient rows to bf16
    assert rel(out, yr) < TOL
    assert rel(x.grad, xr.grad) < TOL, rel(x.grad, xr.grad)
    assert rel(w.grad, wr.grad) < TOL, rel(w.grad, wr.grad)
    if bias:
        assert rel(b.grad, br.grad) < TOL


def test_conv_transpose_fn():
    torch.manual_seed(9)
    B, H, W, C = 2, 10, 14, 64
    M = B * H * W
    x = torch.randn(M, C, device="cuda").bfloat16().requires_grad_()
    w = (torch.randn(C, C, 2, 2, device="cuda") * 0.1).bfloat16().float().requires_grad_()
    out, stats = Fn.conv(x, w, None, Fn.convT_spec(B, H, W), Fn.WLayout("convT"), want_stats=True)
    dy = torch.randn(4 * M, C, device="cuda").bfloat16()
    out.backward(dy)
    xr = x.detach().float().requires_grad_()
    wr = w.detach().clone().requires_grad_()
    yc = F.conv_transpose2d(xr.view(B, H, W, C).permute(0, 3, 1, 2), wr, stride=2)
    yr = yc.permute(0, 2, 3, 1).reshape(4 * M, C)
    yr.backward(dy.float())
    assert rel(out, yr) < TOL and rel(x.grad, xr.grad) < TOL and rel(w.grad, wr.grad) < TOL, (rel(out, yr), rel(x.grad, xr.grad), rel(w.grad, wr.grad))


def test_bn_act_fn_and_residual():
    torch.manual_seed(4)
    M, C = 3000, 128
    bn = torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.01).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.1)
    ref_bn = torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.01).cuda().train()
    ref_bn.load_state_dict(bn.state_dict())
    x = (torch.randn(M, C, device="cuda") * 3 + 1).bfloat16().requires_grad_()
    r = torch.randn(M, C, device="cuda").bfloat16().requires_grad_()
    stats = torch.cat([x.detach().double().sum(0), (x.detach().double() ** 2).sum(0)])
    y = Fn.bn_act(x, stats, bn, relu=True, residual=r)
    dy = torch.randn(M, C, device="cuda").bfloat16()
    y.backward(dy)
    xr = x.detach().float().requires_grad_()
    rr = r.detach().float().requires_grad_()
    # same mask on both sides: evaluate the reference gradient with the product's ReLU pattern
    pre = ref_bn(xr) + rr
    yr = F.relu(pre)
    mask = (y.detach().float() > 0).float()
    (pre * mask * dy.float()).sum().backward()
    assert rel(y, yr) < TOL
    assert rel(x.grad, xr.grad) < TOL and rel(r.grad, rr.grad) < TOL
    assert rel(bn.weight.grad, ref_bn.weight.grad) < TOL and rel(bn.bias.grad, ref_bn.bias.grad) < TOL
    assert rel(bn.running_var, ref_bn.running_var) < 1e-3 and int(bn.num_batches_tracked) >= 1


def test_aspp_branches_fn():
    torch.manual_seed(12)
    B, H, W, C = 1, 24, 24, 256
    M = B * H * W
    o = torch.randn(M, C, device="cuda").bfloat16().requires_grad_()
    idt = torch.randn(M, C, device="cuda").bfloat16().requires_grad_()
    w1 = (torch.randn(C, C, 1, 1, device="cuda") * 0.05).bfloat16().float().requires_grad_()
    ws = (torch.randn(C, C, 3, 3, device="cuda") * 0.02).bfloat16().float().requires_grad_()
    cat = Fn.ASPPBranchesFn.apply(o, idt, w1, ws, B, H, W)
    dy = torch.randn(M, 6 * C, device="cuda").bfloat16()
    cat.backward(dy)
    orf, ir = o.detach().float().requires_grad_(), idt.detach().float().requires_grad_()
    w1r, wsr = w1.detach().clone().requires_grad_(), ws.detach().clone().requires_grad_()
    x = F.relu(orf + ir).bfloat16().float()            # slot 0 is stored in bf16 (the cast is straight-through)
    xc = x.view(B, H, W, C).permute(0, 3, 1, 2)
    br = [xc, F.conv2d(xc, w1r)] + [F.conv2d(xc, wsr, padding=d, dilation=d) for d in (1, 6, 12, 18)]
    ref = torch.cat(br, 1).permute(0, 2, 3, 1).reshape(M, 6 * C)
    ref.backward(dy.float())
    assert rel(cat, ref) < TOL
    assert rel(o.grad, orf.grad) < 2 * TOL and rel(idt.grad, ir.grad) < 2 * TOL, (rel(o.grad, orf.grad))
    assert rel(w1.grad, w1r.grad) < TOL and rel(ws.grad, wsr.grad) < TOL, (rel(w1.grad, w1r.grad), rel(ws.grad, wsr.grad))


def test_densify_fn(pyramid):
    lv = pyramid.levels[4]
    x = torch.randn(lv.n, 256, device="cuda").bfloat16().requires_grad_()
    rows = Fn.DensifyFn.apply(x, lv)
    dy = torch.randn_like(rows)
    rows.backward(dy)
    cv = rows_to_canvas(x.detach().float(), lv.coords[:lv.n], lv.batch, lv.V, lv.U).permute(0, 2, 3, 1).reshape(-1, 256)
    assert torch.equal(rows.float(), cv)
    c = lv.coords[:lv.n].long()
    assert torch.equal(x.grad, dy.view(lv.batch, lv.V, lv.U, 256)[c[:, 0], c[:, 2], c[:, 1]])


def test_sep_head_batched_equals_per_head():
    """The batched sibling-head evaluation (one 64->64*h GEMM, block-diagonal final GEMM) vs the reference's
    per-head Sequential (centerhead.py:53-59) in torch fp32 on bf16-rounded weights/inputs."""
    torch.manual_seed(21)
    heads = {"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2), "hm": (2, 2)}
    sh = modules.SepHead(64, heads, stride=2, bn=True, final_kernel=3).cuda().train()
    with torch.no_grad():
        for p in sh.parameters():
            p.copy_(p.bfloat16().float())
    ref = torch.nn.ModuleDict()
    import copy
    B, H, W = 2, 12, 12
    x = torch.randn(B, 64, H, W, device="cuda").bfloat16()
    # torch reference with the same parameters
    deb_w = sh.deblock.conv.conv.weight.detach().clone().requires_grad_()
    y = F.conv_transpose2d(x.float(), deb_w, stride=2)
    bnr = copy.deepcopy(sh.deblock.norm)
    y = F.relu(bnr(y))
    outs_ref = {}
    seqs = {n: copy.deepcopy(getattr(sh, n)) for n in heads}
    for n in heads:
        outs_ref[n] = seqs[n](y)
    outs = sh(x)
    tot_ref = sum((v * (i + 1)).sum() for i, v in enumerate(outs_ref.values()))
    tot = sum((v * (i + 1)).sum() for i, v in enumerate(outs.values()))
    for n in heads:
        assert tuple(outs[n].shape) == tuple(outs_ref[n].shape)
        assert rel(outs[n], outs_ref[n]) < 3e-2, (n, rel(outs[n], outs_ref[n]))
    tot_ref.backward()
    tot.backward()
    for n in heads:
        for idx in (0, 3):
            a, b = getattr(sh, n)[idx].weight.grad, seqs[n][idx].weight.grad
            assert rel(a, b) < 0.1, (n, idx, rel(a, b))
        a, b = getattr(sh, n)[3].bias.grad, seqs[n][3].bias.grad
        assert rel(a, b) < 1e-2, (n, rel(a, b))
    assert rel(sh.deblock.conv.conv.weight.grad, deb_w.grad) < 0.1


def test_fused_center_loss_matches_torch_loss():
    """pnx_center_loss_* (fused CUDA) vs pillarnext_b200.loss.center_loss (torch restatement pinned to the reference's
    numbers in tests/test_oracle_cpu.py): loss values 1e-4 relative, gradient w.r.t. the head output 1e-3 of its scale."""
    from pillarnext_b200 import loss as PL
    cfg = synth.tiny_config(128, [["car"], ["truck", "construction_vehicle"]])
    torch.manual_seed(0)
    head = modules.CenterHead(256, cfg["tasks"], cfg["weight"], cfg["code_weights"], cfg["common_heads"], cfg["head_strides"],
                              with_reg_iou=True, voxel_size=cfg["voxel_size"], pc_range=cfg["pc_range"],
                              out_size_factor=cfg["out_size_factor"]).cuda().train()
    ex = synth.make_batch([0, 1, 2], 1000, cfg, n_boxes=30)
    exg = {k: [e.cuda() for e in v] for k, v in ex.items() if isinstance(v, list) and torch.is_tensor(v[0])}
    x = torch.randn(3, 256, 16, 16, device="cuda")
    preds = head(x)
    raws = [p["hm"]._pnx_raw["out"] for p in preds]
    total, rets = head.loss(exg, preds)                       # fused path
    gf = torch.autograd.grad(total, raws, retain_graph=True)
    preds2 = [{k: v for k, v in p.items()} for p in preds]    # plain dict views without the fast-path marker
    for p in preds2:
        p["hm"] = p["hm"] + 0
    total2, rets2 = PL.center_loss(exg, preds2, cfg["tasks"], cfg["weight"], cfg["code_weights"], True, cfg["voxel_size"],
                                   cfg["pc_range"], cfg["out_size_factor"])
    gt = torch.autograd.grad(total2, raws)
    assert abs(total.item() - total2.item()) < 1e-4 * abs(total2.item()), (total.item(), total2.item())
    for t in range(len(rets)):
        for k in ("hm_loss", "loc_loss", "iou_reg_loss", "num_positive"):
            assert abs(float(rets[t][k]) - float(rets2[t][k])) < 1e-4 * max(1.0, abs(float(rets2[t][k]))), (t, k)
        assert torch.allclose(rets[t]["loc_loss_elem"], rets2[t]["loc_loss_elem"], rtol=1e-4, atol=1e-6)
        e = (gf[t] - gt[t]).abs().max().item()
        assert e < 1e-3 * gt[t].abs().max().item(), (t, e, gt[t].abs().max().item())


def test_fused_center_loss_matches_oracle_directly():
    """pnx_center_loss_task/_finalize (the fused CUDA loss + gradient) against oracle.center_loss -- the restatement that
    is pinned to the reference's own centerhead.py:142-229 / centerloss.py by tests/golden/ref_tiny.npz -- on the same
    head outputs and labels (NaN velocities, empty tasks included): values 1e-5 relative, d(loss)/d(head output) 1e-4."""
    from oracle import pillarnext_oracle as O
    cfg = synth.tiny_config(128, [["car"], ["truck", "construction_vehicle"], ["bus", "trailer"]])
    torch.manual_seed(1)
    head = modules.CenterHead(256, cfg["tasks"], cfg["weight"], cfg["code_weights"], cfg["common_heads"], cfg["head_strides"],
                              with_reg_iou=True, voxel_size=cfg["voxel_size"], pc_range=cfg["pc_range"],
                              out_size_factor=cfg["out_size_factor"]).cuda().train()
    ex = synth.make_batch([3, 4], 1000, cfg, n_boxes=30)
    for key in ("mask",):
        ex[key][2].zero_()                                     # a task without any object: the "no positives" branch
    exg = {k: [e.cuda() for e in v] for k, v in ex.items() if isinstance(v, list) and torch.is_tensor(v[0])}
    preds = head(torch.randn(2, 256, 16, 16, device="cuda"))
    raws = [p["hm"]._pnx_raw for p in preds]
    total, rets = head.loss(exg, preds)
    gf = torch.autograd.grad(total, [r["out"] for r in raws])
    # the same numbers through the oracle: NCHW fp32 leaves built from the channels-last matrices
    leaves, preds_o = [], []
    for r in raws:
        out4 = r["out"].detach().cpu().view(r["B"], r["H"], r["W"], r["npad"]).clone().requires_grad_()
        leaves.append(out4)
        pd, names = {}, [n for n in cfg["common_heads"]] + ["hm"]
        for n in names:
            c = r["C"] if n == "hm" else cfg["common_heads"][n][0]
            pd[n] = out4[..., r["off"][n]:r["off"][n] + c].permute(0, 3, 1, 2)
        preds_o.append(pd)
    total_o, rets_o = O.center_loss(ex, preds_o, cfg["weight"], cfg["code_weights"], True, cfg["voxel_size"], cfg["pc_range"],
                                    cfg["out_size_factor"])
    go = torch.autograd.grad(total_o, leaves)
    assert abs(total.item() - total_o.item()) < 1e-5 * abs(total_o.item()), (total.item(), total_o.item())
    for t in range(len(rets)):
        for k in ("hm_loss", "loc_loss", "iou_reg_loss", "num_positive"):
            assert abs(float(rets[t][k]) - float(rets_o[t][k])) < 1e-5 * max(1.0, abs(float(rets_o[t][k]))), (t, k)
        assert torch.allclose(rets[t]["loc_loss_elem"].cpu(), rets_o[t]["loc_loss_elem"], rtol=1e-4, atol=1e-6)
        g, o = gf[t].cpu().view_as(go[t]), go[t]
        assert (g - o).abs().max().item() < 1e-4 * o.abs().max().item(), (t, (g - o).abs().max().item(), o.abs().max().item())


@pytest.mark.parametrize("kind", ["igemm_dense", "win", "convT_fanout"])
def test_fused_bn_backward_reduce_matches_separate_pass(kind):
    """The reduce pass of a BatchNorm backward folded into the epilogue of the GEMM that produces dy (pnx_igemm /
    pnx_conv3x3_win `bnr_*`) against the stand-alone pnx_bn_bwd_reduce: same parameter / input gradients up to the
    summation order of the two per-channel sums (fp32 partials, fp64 totals)."""
    torch.manual_seed(3)
    if kind == "win":
        B, H, W, c0, c1, c2 = 2, 40, 256, 64, 64, 128        # 3x3 convs 64 -> 64 -> 128 at a width that takes the window kernel
    else:
        B, H, W, c0, c1, c2 = 2, 24, 40, 64, 128, 64
    M = B * H * W
    x0 = torch.randn(M, c0, device="cuda").bfloat16()
    w1 = (torch.randn(c1, c0, 3, 3, device="cuda") * 0.05)
    R = torch.randn(M * (4 if kind == "convT_fanout" else 1), c2, device="cuda").bfloat16()
    if kind == "convT_fanout":
        w2s = [torch.randn(c1, c2, 2, 2, device="cuda") * 0.1 for _ in range(3)]
    else:
        w2s = [torch.randn(c2, c1, 3, 3, device="cuda") * 0.05]

    base_launches = [0]

    def run(fused):
        l0 = ops.LAUNCHES
        x = x0.clone().requires_grad_()
        wa = w1.clone().requires_grad_()
        wbs = [w.clone().requires_grad_() for w in w2s]
        bn = torch.nn.BatchNorm2d(c1).cuda().train()
        info = Fn.bn_info() if fused else None
        raw, stats = Fn.conv(x, wa, None, Fn.dense_spec(B, H, W, 3), Fn.WLayout("dense"), want_stats=True)
        y = Fn.bn_act(raw, stats, bn, relu=True, info=info)
        tot = 0
        for wb in wbs:
            if kind == "convT_fanout":
                o, _ = Fn.conv(y, wb, None, Fn.convT_spec(B, H, W), Fn.WLayout("convT"), bn_src=info)
            else:
                o, _ = Fn.conv(y, wb, None, Fn.dense_spec(B, H, W, 3), Fn.WLayout("dense"), bn_src=info)
            tot = tot + (o.float() * R.float()).sum()
        tot.backward()
        assert (info is None) or (info.red is None and not info.fused)     # consumed by the BatchNorm backward
        if fused and Fn.FUSE_BN_REDUCE_MIN_K == 0:
            assert ops.LAUNCHES - l0 < base_launches[0] or base_launches[0] == 0
        return x.grad.float(), wa.grad, bn.weight.grad, bn.bias.grad, [w.grad for w in wbs]

    prev, Fn.FUSE_BN_REDUCE_MIN_K = Fn.FUSE_BN_REDUCE_MIN_K, 0       # exercise the kernel path for every shape
    try:
        a = run(True)
    finally:
        Fn.FUSE_BN_REDUCE_MIN_K = prev
    b = run(False)
    # fan-out: the separate pass reduces the bf16-rounded SUM of the consumers' gradients, the fused one each part
    tol = 5e-3 if kind == "convT_fanout" else 1e-4
    assert rel(a[0], b[0]) < 5e-3 and rel(a[1], b[1]) < 5e-3, (rel(a[0], b[0]), rel(a[1], b[1]))
    assert rel(a[2], b[2]) < tol and rel(a[3], b[3]) < tol, (rel(a[2], b[2]), rel(a[3], b[3]))
    for ga, gb in zip(a[4], b[4]):
        assert rel(ga, gb) < 1e-5                                            # untouched by the fusion (fp32 atomics: order only)
