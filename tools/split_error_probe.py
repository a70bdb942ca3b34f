"""GPU probe: where does the error of the split-mode GEMM come from?  (run under gpurun)
 (1) pure accumulation error of the tensor core: bf16-exact inputs, plain bf16 kernel vs fp64
 (2) split GEMM nseg 3/4 vs fp64 on fp32 inputs
 (3) torch fp32 matmul (no TF32) vs fp64 for scale"""
import torch
from pillarnext_b200 import functional as Fn, ops

torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return ((a.double() - b).norm() / b.norm()).item()


for (M, taps, cin, cout) in [(8192, 1, 64, 64), (8192, 9, 64, 64), (8192, 9, 256, 256), (8192, 1, 1536, 256)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(M, cin, device="cuda", generator=g)
    w = torch.randn(taps, cout, cin, device="cuda", generator=g) * 0.1
    nbr = torch.randint(0, M, (M, taps), device="cuda", generator=g, dtype=torch.int32) if taps > 1 else None

    def ref_of(a_, w_):
        r = torch.zeros(M, cout, dtype=torch.float64, device="cuda")
        for t in range(taps):
            src = a_.double() if nbr is None else a_.double()[nbr[:, t].long()]
            r += src @ w_[t].double().t()
        return r

    # (1) accumulation only
    ab, wb = a.bfloat16(), w.bfloat16().contiguous()
    out = torch.empty(M, cout, dtype=torch.float32, device="cuda")
    ops.igemm(ab, M, wb, taps, cin, cout, out, nbr=nbr)
    r_b = ref_of(ab.float(), wb.float())
    e_acc = rel(out, r_b)
    bias = ((out.double() - r_b) * r_b.sign()).mean().item() / r_b.abs().mean().item()
    # (2) split
    ref = ref_of(a, w)
    es = []
    for nseg in (3, 4):
        ops.igemm(ops.rows_split(a), M, Fn._to_hilo(w), taps, cin, cout, out, nbr=nbr, nseg=nseg, a_lo_off=cin)
        es.append(rel(out, ref))
    # (3) torch fp32
    o32 = torch.zeros(M, cout, device="cuda")
    for t in range(taps):
        src = a if nbr is None else a[nbr[:, t].long()]
        o32 += src @ w[t].t()
    print("K=%5d  accumulation-only rel %.2e (signed bias %.2e) | split nseg3 %.2e nseg4 %.2e | torch fp32 %.2e" %
          (taps * cin, e_acc, bias, es[0], es[1], rel(o32, ref)))
