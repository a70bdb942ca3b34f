"""Scratch micro-benchmark: CUDA-event timing of individual igemm / wgrad shapes from the PillarNeXt-B step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pillarnext_b200 import ops

def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

B, H, W = 6, 336, 336
M = B * H * W
geo = (H, W, H, W, 3, 1, 1, 1)
res = []
def run_igemm(name, K, N, bn, stats, bias, fp32=False, m=M, g=geo, taps=9):
    A = torch.randn(m, K, device="cuda").bfloat16(); Wp = (torch.randn(taps, N, K, device="cuda") * 0.05).bfloat16()
    out = torch.empty(m, N, device="cuda", dtype=torch.float32 if fp32 else torch.bfloat16)
    st = torch.zeros(2 * N, dtype=torch.float64, device="cuda") if stats else None
    bs = torch.randn(N, device="cuda") if bias else None
    ms = timeit(lambda: ops.igemm(A, m, Wp, taps, K, N, out, dense=g if taps > 1 else None, stats=st, bias=bs, block_n=bn))
    fl = 2.0 * m * taps * K * N
    print("%-42s %7.3f ms %7.1f TF/s" % (name, ms, fl / ms / 1e9), flush=True)

run_igemm("convA 64->384 bn192 plain", 64, 384, 192, False, False)
run_igemm("convA 64->384 bn192 +stats", 64, 384, 192, True, False)
run_igemm("convA 64->384 bn192 +stats+bias", 64, 384, 192, True, True)
run_igemm("convA 64->384 bn128 +stats+bias", 64, 384, 128, True, True)
run_igemm("convA-dgrad 384->64 bn64", 384, 64, 64, False, False)
run_igemm("convB 384->16 fp32 +bias", 384, 16, 16, False, True, fp32=True)
run_igemm("convB-dgrad 64->384 bn192", 64, 384, 192, False, False)
m2 = 6 * 168 * 168; g2 = (168, 168, 168, 168, 3, 1, 1, 1)
run_igemm("neck 256->256 bn256 +stats", 256, 256, 256, True, False, m=m2, g=g2)
run_igemm("neck 256->256 bn256 plain", 256, 256, 256, False, False, m=m2, g=g2)
run_igemm("neck 256->256 bn128 plain", 256, 256, 128, False, False, m=m2, g=g2)
run_igemm("1x1 1536->256", 1536, 256, 256, True, False, m=m2, taps=1)
run_igemm("1x1 256->1536", 256, 1536, 256, False, False, m=m2, taps=1)
def run_wgrad(name, X, Y, m=M, g=geo, taps=9):
    Xt = torch.randn(m, X, device="cuda").bfloat16(); Yt = torch.randn(m, Y, device="cuda").bfloat16()
    dW = torch.zeros(taps, X, Y, device="cuda")
    ms = timeit(lambda: ops.wgrad(Xt, X, Yt, Y, m, taps, dW, dense=g))
    print("%-42s %7.3f ms %7.1f TF/s" % (name, ms, 2.0 * m * taps * X * Y / ms / 1e9), flush=True)
run_wgrad("wgrad X384 Y64", 384, 64)
run_wgrad("wgrad X64 Y384", 64, 384)
run_wgrad("wgrad X256 Y256", 256, 256, m=m2, g=g2)
run_wgrad("wgrad X64 Y64", 64, 64)
def run_win(name, K, N, bn, m=M):
    A = torch.randn(m, K, device="cuda").bfloat16(); Wp = (torch.randn(9, N, K, device="cuda") * 0.05).bfloat16()
    out = torch.empty(m, N, device="cuda", dtype=torch.bfloat16)
    st = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: ops.conv3x3_win(A, B, H, W, Wp, K, N, out, stats=st, block_n=bn))
    print("%-42s %7.3f ms %7.1f TF/s" % (name, ms, 2.0 * m * 9 * K * N / ms / 1e9), flush=True)
run_win("WIN convA 64->384 bn192 +stats", 64, 384, 192)
run_win("WIN convA 64->384 auto(ws128) +stats", 64, 384, None)
run_win("WIN convA-dgrad 384->64 bn64", 384, 64, 64)
run_win("WIN 256->64 bn64", 256, 64, 64)
