"""Two-function stand-in for the third-party ``torch_scatter`` package (TEST INFRASTRUCTURE).

The reference calls exactly ``torch_scatter.scatter_max(src, index, dim=0)[0]`` and
``torch_scatter.scatter_mean(src, index, dim=0)`` (/root/reference/det3d/models/readers/
pillar_encoder.py:43,113-114,180).  torch_scatter is not installed and its source is not in
the reference tree (pip, unpinned: docker/Dockerfile:16); its published semantics are
restated here on plain torch CPU ops.
"""
import torch


def scatter_max(src, index, dim=0, dim_size=None):
    assert dim == 0
    n = int(index.max()) + 1 if dim_size is None else dim_size
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    out = torch.full((n,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype)
    out = out.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    # arg: smallest source row attaining the max (torch_scatter returns one attaining row)
    hit = src == out[index]
    rows = torch.arange(src.shape[0]).view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    big = torch.full_like(rows, src.shape[0])
    arg = torch.full((n,) + tuple(src.shape[1:]), src.shape[0], dtype=torch.long)
    arg = arg.scatter_reduce(0, idx, torch.where(hit, rows, big), reduce="amin", include_self=True)
    return out, arg


def scatter_mean(src, index, dim=0, dim_size=None):
    assert dim == 0
    n = int(index.max()) + 1 if dim_size is None else dim_size
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype)
    out.index_add_(0, index, src)                     # sequential, ascending row order on CPU
    cnt = torch.zeros(n, dtype=src.dtype)
    cnt.index_add_(0, index, torch.ones(index.shape[0], dtype=src.dtype))
    cnt = cnt.clamp(min=1)
    return out / cnt.view(-1, *([1] * (src.dim() - 1)))
