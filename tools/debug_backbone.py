"""Scratch diagnostic: per-layer error of the CUDA sparse backbone vs the bf16-faithful oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import pillarnext_oracle as O
from oracle.weights import randomize_state_dict
from pillarnext_b200 import modules, synth, ops, functional as Fn

TASKS = [["car"], ["truck", "construction_vehicle"]]
cfg = synth.tiny_config(128, TASKS)
model = modules.build_pillarnext_b(cfg)
sd = randomize_state_dict(model.state_dict(), 3)
model.load_state_dict(sd); model = model.cuda().train()
ex = synth.make_batch([0, 1], 3000, cfg, kind="uniform", n_boxes=25, sweeps=10)
B = 2
O.QUANT = True
feat_o, coords_o, grid_o = O.reader_forward(ex["points"], sd, cfg["voxel_size"], cfg["pc_range"], train=True)
model.reader.batch_size = B
feat, coords, grid = model.reader(ex["points"].cuda())
vox = feat._pnx_vox; pyr = vox.pyramid

def dens_mine(x, lv):
    C = x.shape[1]
    cv = torch.zeros(lv.batch, lv.V, lv.U, C)
    c = lv.coords[:lv.n].cpu().long()
    cv[c[:, 0], c[:, 2], c[:, 1]] = x[:lv.n].float().cpu()
    return cv
def dens_o(x, c, shape):
    cv = torch.zeros(shape[0], shape[1], shape[2], x.shape[1])
    cv[c[:, 0], c[:, 1], c[:, 2]] = x.detach()
    return cv
def rel(a, b): return ((a - b).norm() / (b.norm() + 1e-12)).item()

xo = O.q(feat_o.detach()); co = coords_o.long(); shape = (B, int(grid_o[0]), int(grid_o[1]))
xm = vox.feat_bf16[:vox.P]
print("input", rel(dens_mine(xm, pyr.levels[0]), dens_o(xo, co, shape)))
bb = model.backbone
with torch.no_grad():
  for s in range(4):
    p = "backbone.blocks.%d." % s
    oc, oshape = O._dilate_sites(co, shape, cfg["strides"][s])
    raw_o = O._gather_conv(xo, co, shape, oc, sd[p + "0.conv.weight"], cfg["strides"][s])
    co, shape = oc, oshape
    blk = bb.blocks[s][0]
    raw_m, stats = Fn.conv(xm, blk.conv.weight, None, pyr.entry[s], Fn.WLayout("sp"), want_stats=True)
    lv = pyr.levels[s + 1]
    print("stage", s, "entry conv raw", rel(dens_mine(raw_m, lv), dens_o(raw_o, co, shape)), "n", lv.n, raw_o.shape[0])
    n = raw_o.shape[0]
    st = stats.cpu()
    C = raw_o.shape[1]
    print("   stats sum err", (st[:C] - raw_o.double().sum(0)).abs().max().item(), "sq", ((st[C:] - (raw_o.double()**2).sum(0)).abs() / (raw_o.double()**2).sum(0)).max().item())
    xo = O.q(F.relu(O._bn(raw_o, sd, p + "0.norm.", 1e-3, [0], 1, True, None, 0.01)))
    xm = Fn.bn_act(raw_m, stats, blk.norm, relu=True)
    print("   after bn", rel(dens_mine(xm, lv), dens_o(xo, co, shape)))
    for j in (1, 2):
        bq = "%s%d." % (p, j)
        mb = bb.blocks[s][j]
        idt_o, idt_m = xo, xm
        o = O._gather_conv(xo, co, shape, co, sd[bq + "block1.conv.weight"], 1)
        m_raw, m_st = Fn.conv(xm, mb.block1.conv.weight, None, pyr.subm[s], Fn.WLayout("sp"), want_stats=True)
        print("   blk", j, "conv1 raw", rel(dens_mine(m_raw, lv), dens_o(o, co, shape)))
        o = O.q(F.relu(O._bn(o, sd, bq + "block1.norm.", 1e-3, [0], 1, True, None, 0.01)))
        m = Fn.bn_act(m_raw, m_st, mb.block1.norm, relu=True)
        print("   blk", j, "bn1", rel(dens_mine(m, lv), dens_o(o, co, shape)))
        o = O._gather_conv(o, co, shape, co, sd[bq + "conv2.weight"], 1)
        m_raw, m_st = Fn.conv(m, mb.conv2.weight, None, pyr.subm[s], Fn.WLayout("sp"), want_stats=True)
        print("   blk", j, "conv2 raw", rel(dens_mine(m_raw, lv), dens_o(o, co, shape)))
        o = O._bn(o, sd, bq + "norm2.", 1e-3, [0], 1, True, None, 0.01)
        xo = O.q(F.relu(o + idt_o))
        xm = Fn.bn_act(m_raw, m_st, mb.norm2, relu=True, residual=idt_m)
        print("   blk", j, "out", rel(dens_mine(xm, lv), dens_o(xo, co, shape)))
