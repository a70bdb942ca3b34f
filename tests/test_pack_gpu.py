"""pnx_pack_weights (csrc/pack.cu): the one-launch repack must reproduce the per-weight torch packing bit for bit, for every
layout / flip combination, and must be what a training step uses after an optimizer step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def test_batched_repack_matches_torch_packing(dev):
    from pillarnext_b200 import functional as F
    g = torch.Generator(device="cpu").manual_seed(3)
    params = [("dense", (64, 32, 3, 3)), ("dense", (10, 64, 1, 1)), ("sp", (64, 3, 3, 32)), ("sp", (128, 3, 3, 64)),
              ("convT", (128, 64, 2, 2)), ("dense", (33, 64, 3, 3))]
    ws, reqs = [], []
    for kind, shape in params:
        w = torch.nn.Parameter(torch.randn(*shape, generator=g).to(dev))
        ws.append(w)
        lay = F.WLayout(kind)
        for which, flip in (("fwd", False), ("dgrad", False), ("dgrad", True)):
            if kind == "convT" and flip:
                continue
            reqs.append((w, lay, which, flip))
    first = [F.packed(*r) for r in reqs]
    ptrs = [t.data_ptr() for t in first]
    with torch.no_grad():
        for w in ws:
            w.mul_(1.5).add_(0.25)              # bumps the version -> every entry stale
    again = [F.packed(*r) for r in reqs]
    assert [t.data_ptr() for t in again] == ptrs, "the batched path rewrites the cached operands in place"
    for (w, lay, which, flip), got in zip(reqs, again):
        want = lay.pack_fwd(w.detach()) if which == "fwd" else lay.pack_dgrad(w.detach(), flip)
        assert got.shape == want.shape and torch.equal(got, want), (lay.kind, which, flip)
    # optimizer-generation path (fused AdamW does not bump the version)
    opt = torch.optim.AdamW(ws, lr=0.1, fused=True)
    for w in ws:
        w.grad = torch.ones_like(w)
    opt.step()
    third = [F.packed(*r) for r in reqs]
    for (w, lay, which, flip), got in zip(reqs, third):
        want = lay.pack_fwd(w.detach()) if which == "fwd" else lay.pack_dgrad(w.detach(), flip)
        assert torch.equal(got, want)
