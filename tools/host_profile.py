"""cProfile of the host-side enqueue of one training step (no synchronisation inside the profiled region)."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pillarnext_b200 import modules, synth

def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    cfg = synth.NUSC
    torch.manual_seed(0)
    model = modules.build_pillarnext_b(cfg).to(dev).train()
    params = list(model.parameters())
    opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.99), weight_decay=0.01, fused=True)
    host = [bench.pin(synth.make_batch([b * 6 + f for f in range(6)], 30000, cfg, kind="lidar", n_boxes=40, sweeps=10)) for b in range(2)]
    resident = [bench.to_device(h, dev) for h in host]

    def step(ex):
        loss, _ = model(ex)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for i in range(4):
        step(resident[i % 2])
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for i in range(3):
        step(resident[i % 2])
    pr.disable()
    torch.cuda.synchronize()
    out = io.StringIO()
    st = pstats.Stats(pr, stream=out)
    st.sort_stats("cumulative").print_stats(60)
    st.sort_stats("tottime").print_stats(40)
    print(out.getvalue())

main()
