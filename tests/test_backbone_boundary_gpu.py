"""Backbone boundary and the spconv weight / kernel-offset convention.

(1) `SparseResNet.forward(features, coors, input_shape)` with PLAIN tensors -- the reference's own calling convention
    (det3d/models/backbones/sparse_resnet.py:61-64), any row order -- gives the same dense map as the fused
    reader -> backbone hand-off.
(2) The convention pin.  spconv is absent, so the kernel-offset <-> weight-index mapping is pinned to
    spconv 2.x's documented definition instead of to a run of the library:
      * `SubMConv2d.weight` has layout [Cout, kH, kW, Cin] (spconv/pytorch/conv.py, `weight_shape = [out_channels,
        *kernel_size, in_channels]` for the default KRSC layout) and is state-dict compatible with the reference;
      * spatial_shape / indices are (batch, y, x) as the reference builds them (sparse_resnet.py:63-64: coors (b, y, x),
        input_shape (H, W)), so kernel axis kH runs along y and kW along x;
      * the operation is a cross-correlation, like nn.Conv2d: out[p] = sum_{kh,kw} W[:, kh, kw, :] . in[p + (kh-1, kw-1)].
    A hand-computed, deliberately asymmetric 3x3 kernel on three active sites distinguishes all eight transposed /
    flipped readings; the oracle (CPU) and the product (GPU) must both reproduce the hand numbers, so a real checkpoint
    cannot load transposed without this test failing."""
import pytest
import torch

from oracle import pillarnext_oracle as O
from pillarnext_b200 import functional as Fn
from pillarnext_b200 import modules, ops, synth


def hand_case():
    """Three active sites on an 8x8 grid (frame 0): A = (y 5, x 5) value 1, B = (y 5, x 6) value 2, C = (y 6, x 5)
    value 3, carried by input channel 0; W[co=0, kh, kw, ci=0] = 10*kh + kw + 1.  SubM output, channel 0:
      out[A] = W[1,1]*1 + W[1,2]*2 (B is at offset (0,+1)) + W[2,1]*3 (C at (+1,0)) = 12 + 26 + 66 = 104
      out[B] = W[1,1]*2 + W[1,0]*1 (A at (0,-1)) + W[2,0]*3 (C at (+1,-1))         = 24 + 11 + 63 = 98
      out[C] = W[1,1]*3 + W[0,1]*1 (A at (-1,0)) + W[0,2]*2 (B at (-1,+1))         = 36 +  2 +  6 = 44"""
    coords = torch.tensor([[0, 5, 5], [0, 5, 6], [0, 6, 5]], dtype=torch.int32)           # (b, y, x)
    feat = torch.zeros(3, 64)
    feat[:, 0] = torch.tensor([1.0, 2.0, 3.0])
    w = torch.zeros(64, 3, 3, 64)                                                          # [Cout, kH, kW, Cin]
    for kh in range(3):
        for kw in range(3):
            w[0, kh, kw, 0] = 10 * kh + kw + 1
    want = torch.tensor([104.0, 98.0, 44.0])
    return coords, feat, w, want


def test_oracle_follows_the_spconv_convention():
    coords, feat, w, want = hand_case()
    out = O._gather_conv(feat, coords.long(), (1, 8, 8), coords.long(), w, 1)
    assert torch.equal(out[:, 0], want)
    # and the dense restatement (F.conv2d on the canvas through _spw) agrees
    canvas = torch.zeros(1, 64, 8, 8)
    canvas[0, :, coords[:, 1].long(), coords[:, 2].long()] = feat.t()
    dense = torch.nn.functional.conv2d(canvas, O._spw(w), padding=1)
    assert torch.equal(dense[0, 0, coords[:, 1].long(), coords[:, 2].long()], want)


@pytest.mark.gpu
def test_product_follows_the_spconv_convention():
    coords, feat, w, want = hand_case()
    bb = modules.SparseResNet([2, 2, 2, 2], [1, 2, 2, 2], [64, 128, 256, 256], 64).cuda()
    vox, rows, order = bb._pyramid_from_plain(feat.cuda(), coords.cuda(), (8, 8))
    lv = ops.level_from_bitmap(vox.bitmap, vox.blockpref, vox.counts[0:1], vox.batch, vox.gx, vox.gy, inblk=vox.inblk)
    ops.level_coords(lv, 3)
    sub = ops.nbr_table(lv, lv, 1, False)
    spec = Fn.ConvSpec(3, 3, 9, nbr=sub, d_nbr=sub, d_flip=True)
    out, _ = Fn.conv(rows.to(torch.bfloat16), w.cuda(), None, spec, Fn.WLayout("sp"))
    got = torch.empty(3)
    got[order.cpu()] = out[:, 0].float().cpu()                  # back to the caller's row order
    assert torch.equal(got, want), got                           # small integers: exact in bf16 x fp32 accumulation
    # the regular (dilating) SparseConv2d of the stage entries uses the same mapping: output site D = (y 4, x 5), which is
    # not active in the input, receives W[2,1]*A (A is at offset (+1, 0)) + W[2,2]*B = 22 + 46 = 68
    nxt = ops.level_dilate(lv, 1)
    n = int(nxt.count.item())
    ops.level_coords(nxt, n)
    spec_e = Fn.ConvSpec(n, 3, 9, nbr=ops.nbr_table(nxt, lv, 1, False), d_nbr=ops.nbr_table(lv, nxt, 1, True), d_flip=False)
    out_e, _ = Fn.conv(rows.to(torch.bfloat16), w.cuda(), None, spec_e, Fn.WLayout("sp"))
    c = nxt.coords[:n].cpu()                                      # (b, u = x, v = y)
    row = int(((c[:, 1] == 5) & (c[:, 2] == 4)).nonzero()[0])
    assert float(out_e[row, 0]) == 68.0


@pytest.mark.gpu
def test_backbone_accepts_plain_tensors_in_any_row_order():
    cfg = synth.tiny_config(128, [["car"]])
    torch.manual_seed(0)
    model = modules.build_pillarnext_b(cfg).cuda().eval()
    pts = synth.collate_points([synth.make_frame(s, 3000, cfg, "uniform") for s in range(2)]).cuda()
    model.reader.batch_size = 2
    with torch.no_grad():
        feat, coords, grid = model.reader(pts)
        ref = model.backbone(feat, coords, grid)                              # fused hand-off (feat carries the rulebook)
        plain = model.backbone(feat.clone(), coords.clone(), grid)            # plain tensors, reader order
        perm = torch.randperm(feat.shape[0], device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        shuf = model.backbone(feat[perm].clone(), coords[perm].clone(), grid)  # arbitrary row order
    assert torch.equal(plain, ref) and torch.equal(shuf, ref)
    # training mode with gradients through the plain path
    model.train()
    f2 = feat.clone().requires_grad_()
    out = model.backbone(f2, coords.clone(), grid)
    out.float().sum().backward()
    assert f2.grad is not None and torch.isfinite(f2.grad).all() and float(f2.grad.abs().sum()) > 0
