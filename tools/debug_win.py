import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from pillarnext_b200 import ops
B, H, W, Cin, Cout = 1, 4, 200, 64, 64
torch.manual_seed(0)
x = torch.randn(B, Cin, H, W, device="cuda").bfloat16()
rows = x.permute(0, 2, 3, 1).contiguous().view(-1, Cin)
for mode in (0, 1):
    for t in range(9):
        w = torch.zeros(Cout, Cin, 3, 3, device="cuda")
        w[:, :, t // 3, t % 3] = torch.randn(Cout, Cin, device="cuda") * 0.1
        w = w.bfloat16()
        ref = F.conv2d(x.float(), w.float(), padding=1)
        wp = w.permute(2, 3, 0, 1).contiguous().view(9, Cout, Cin)
        out = torch.zeros(B * H * W, Cout, dtype=torch.bfloat16, device="cuda")
        ops.conv3x3_win(rows, B, H, W, wp, Cin, Cout, out, block_n=64, base_off=mode)
        got = out.view(B, H, W, Cout).permute(0, 3, 1, 2).float()
        d = (got - ref).abs()
        # per x-position error profile for row y=1
        e = d[0, :, 1, :].max(0)[0]
        bad = (e > 0.05).nonzero().flatten().tolist()
        print("mode", mode, "tap", t, "(r,s)=", (t // 3, t % 3), "rel err %.3g" % (d.max().item() / ref.abs().max().item()), "bad x (y=1):", bad[:10], "n_bad", len(bad))
