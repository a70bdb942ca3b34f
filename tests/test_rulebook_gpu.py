"""B1-B4 site sets / neighbour tables vs brute force (spconv semantics: SURVEY.md 8c)."""
import pytest
import torch
import torch.nn.functional as F

from pillarnext_b200 import ops

pytestmark = pytest.mark.gpu


def make_level(mask):
    """mask [B, U, V] bool (cpu) -> Level with bitmap."""
    B, U, V = mask.shape
    vw = (V + 31) // 32
    bits = torch.zeros(B, U, vw * 32, dtype=torch.int64)
    bits[:, :, :V] = mask.long()
    w = (bits.view(B, U, vw, 32) << torch.arange(32)).sum(-1)
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32).reshape(-1)
    pad = (-w.numel()) % 32
    bm = torch.cat([w, torch.zeros(pad, dtype=torch.int32)]).cuda()
    return ops.level_from_mask_words(bm, B, U, V)


def unpack(lv):
    vw = (lv.V + 31) // 32
    w = lv.bm.cpu().long()[:lv.batch * lv.U * vw] & 0xFFFFFFFF
    bits = (w.view(lv.batch, lv.U, vw, 1) >> torch.arange(32)) & 1
    return bits.reshape(lv.batch, lv.U, vw * 32)[:, :, :lv.V].bool()


@pytest.mark.parametrize("B,U,V,stride,density", [(2, 40, 40, 1, 0.05), (1, 64, 96, 2, 0.03), (2, 37, 45, 2, 0.1), (1, 1344, 1344, 1, 0.01), (1, 336, 336, 2, 0.2)])
def test_dilate_and_tables(B, U, V, stride, density):
    g = torch.Generator().manual_seed(U * 7 + V)
    mask = torch.rand(B, U, V, generator=g) < density
    src = make_level(mask)
    n_src = int(src.count.item())
    assert n_src == int(mask.sum())
    ops.level_coords(src, n_src)
    assert torch.equal(src.coords[:n_src].cpu().long(), torch.nonzero(mask))
    dst = ops.level_dilate(src, stride)
    ref = F.max_pool2d(mask.float().unsqueeze(1), 3, stride, 1)[:, 0] > 0
    assert torch.equal(unpack(dst), ref)
    n_dst = int(dst.count.item())
    assert n_dst == int(ref.sum())
    ops.level_coords(dst, n_dst)
    dc = torch.nonzero(ref)
    assert torch.equal(dst.coords[:n_dst].cpu().long(), dc)
    # forward table
    idx_src = torch.full((B, U, V), -1, dtype=torch.long)
    idx_src[mask] = torch.arange(n_src)
    nbr = ops.nbr_table(dst, src, stride, False)[:n_dst].cpu().long()
    exp = torch.full((n_dst, 9), -1, dtype=torch.long)
    for ku in range(3):
        for kv in range(3):
            u = dc[:, 1] * stride + ku - 1
            v = dc[:, 2] * stride + kv - 1
            ok = (u >= 0) & (u < U) & (v >= 0) & (v < V)
            exp[ok, ku * 3 + kv] = idx_src[dc[ok, 0], u[ok], v[ok]]
    assert torch.equal(nbr, exp)
    # transposed table: nbrT[j][t] == i  <=>  nbr[i][t] == j
    nbrT = ops.nbr_table(src, dst, stride, True)[:n_src].cpu().long()
    expT = torch.full((n_src, 9), -1, dtype=torch.long)
    ii, tt = torch.nonzero(exp >= 0, as_tuple=True)
    expT[exp[ii, tt], tt] = ii
    assert torch.equal(nbrT, expT)
    # submanifold table on dst
    sub = ops.nbr_table(dst, dst, 1, False)[:n_dst].cpu().long()
    idx_dst = torch.full(ref.shape, -1, dtype=torch.long)
    idx_dst[ref] = torch.arange(n_dst)
    Uo, Vo = ref.shape[1], ref.shape[2]
    for ku in range(3):
        for kv in range(3):
            u = dc[:, 1] + ku - 1
            v = dc[:, 2] + kv - 1
            ok = (u >= 0) & (u < Uo) & (v >= 0) & (v < Vo)
            e = torch.full((n_dst,), -1, dtype=torch.long)
            e[ok] = idx_dst[dc[ok, 0], u[ok], v[ok]]
            assert torch.equal(sub[:, ku * 3 + kv], e)
    # dense scatter / gather round trip
    feat = torch.randn(n_dst, 64, device="cuda").bfloat16()
    canvas = ops.scatter_dense(feat, dst, 64)
    exp_c = torch.zeros(B, Vo, Uo, 64, dtype=torch.bfloat16)
    exp_c[dc[:, 0], dc[:, 2], dc[:, 1]] = feat.cpu()
    assert torch.equal(canvas.cpu(), exp_c)
    back = torch.empty_like(feat)
    ops.scatter_dense(back, dst, 64, canvas=canvas, gather=True)
    assert torch.equal(back, feat)
