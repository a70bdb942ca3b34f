"""One steady-state training step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pillarnext_b200 import modules, synth

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
cfg = synth.NUSC
torch.manual_seed(0)
model = modules.build_pillarnext_b(cfg).to(dev).train()
opt = torch.optim.AdamW(list(model.parameters()), lr=1e-4, betas=(0.9, 0.99), weight_decay=0.01, fused=True)
ex = bench.to_device(synth.make_batch(list(range(6)), 30000, cfg, kind="lidar", n_boxes=40, sweeps=10), dev)

def step():
    loss, _ = model(ex)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)

for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
