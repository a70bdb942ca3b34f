"""Scratch: launch one wgrad / igemm shape from the PillarNeXt-B step (for ncu --set full captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pillarnext_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else "wgrad_head"
torch.manual_seed(0)
B, H, W = 6, 336, 336
M = B * H * W
if which == "wgrad_head":      # head conv A weight gradient: X = dy [M,384] direct, Y = In [M,64] gathered 3x3
    X = torch.randn(M, 384, device="cuda").bfloat16(); Y = torch.randn(M, 64, device="cuda").bfloat16()
    dW = torch.zeros(9, 384, 64, device="cuda")
    for _ in range(2):
        ops.wgrad(X, 384, Y, 64, M, 9, dW, dense=(H, W, H, W, 3, 1, 1, 1))
elif which == "wgrad_256":
    M = 6 * 168 * 168
    X = torch.randn(M, 256, device="cuda").bfloat16(); Y = torch.randn(M, 256, device="cuda").bfloat16()
    dW = torch.zeros(9, 256, 256, device="cuda")
    for _ in range(2):
        ops.wgrad(X, 256, Y, 256, M, 9, dW, dense=(168, 168, 168, 168, 3, 1, 1, 1))
elif which == "igemm_head":    # head conv A forward 64 -> 384
    A = torch.randn(M, 64, device="cuda").bfloat16(); Wp = torch.randn(9, 384, 64, device="cuda").bfloat16()
    out = torch.empty(M, 384, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.igemm(A, M, Wp, 9, 64, 384, out, dense=(H, W, H, W, 3, 1, 1, 1))
elif which == "igemm_256":
    M = 6 * 168 * 168
    A = torch.randn(M, 256, device="cuda").bfloat16(); Wp = torch.randn(9, 256, 256, device="cuda").bfloat16()
    out = torch.empty(M, 256, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.igemm(A, M, Wp, 9, 256, 256, out, dense=(168, 168, 168, 168, 3, 1, 1, 1))
torch.cuda.synchronize()
if which == "igemm_1x1":
    M = 6 * 168 * 168
    A = torch.randn(M, 256, device="cuda").bfloat16(); Wp = torch.randn(1, 1536, 256, device="cuda").bfloat16()
    out = torch.empty(M, 1536, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.igemm(A, M, Wp, 1, 256, 1536, out, block_n=256)
    torch.cuda.synchronize()
if which == "win_dgrad":       # head conv A data gradient 384 -> 64 at 336^2 through the TMA-window kernel
    A = torch.randn(M, 384, device="cuda").bfloat16(); Wp = torch.randn(9, 64, 384, device="cuda").bfloat16()
    out = torch.empty(M, 64, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.conv3x3_win(A, B, H, W, Wp, 384, 64, out)
    torch.cuda.synchronize()
if which == "win_fwd":
    A = torch.randn(M, 64, device="cuda").bfloat16(); Wp = torch.randn(9, 384, 64, device="cuda").bfloat16()
    out = torch.empty(M, 384, device="cuda", dtype=torch.bfloat16)
    st = torch.zeros(768, dtype=torch.float64, device="cuda"); bs = torch.randn(384, device="cuda")
    for _ in range(2):
        ops.conv3x3_win(A, B, H, W, Wp, 64, 384, out, bias=bs, stats=st)
    torch.cuda.synchronize()
