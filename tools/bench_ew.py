"""GB/s of the BatchNorm row kernels at the shapes of the training step (tuning aid; the rows-in-flight factors it
was used to choose are now compile-time constants in csrc/elementwise.cu)."""
import sys, torch
sys.path.insert(0, '.')
from pillarnext_b200 import ops

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

dev = 'cuda'
for (M, C) in [(777600, 384), (194400, 256), (777600, 64), (140000, 64), (60000, 128)]:
    x = torch.randn(M, C, device=dev).bfloat16()
    dy = torch.randn(M, C, device=dev).bfloat16()
    y = torch.empty_like(x); dx = torch.empty_like(x)
    sc = torch.rand(C, device=dev) + 0.5; sh = torch.randn(C, device=dev)
    mean = torch.randn(C, device=dev) * 0.1; invstd = torch.rand(C, device=dev) + 0.5
    gamma = torch.ones(C, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    t_apply = timeit(lambda: ops.bn_apply(x, M, C, sc, sh, y, relu=True))
    def bwd():
        ops.bn_bwd(dy, None, x, M, C, mean, invstd, gamma, M, True, dx, affine=(sc, sh))
    t_bwd = timeit(bwd)
    b = M * C * 2
    print(f"M={M} C={C}: bn_apply {t_apply*1e3:.0f} us {2*b/t_apply/1e6:.0f} GB/s | bn_bwd(reduce+apply) {t_bwd*1e3:.0f} us {5*b/t_bwd/1e6:.0f} GB/s")
