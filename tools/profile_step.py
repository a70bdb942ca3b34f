"""Scratch: torch.profiler (kineto) kernel table of one steady-state training step of bench.py's workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, collections, re
from torch.profiler import profile, ProfilerActivity
from pillarnext_b200 import modules, synth
import bench
cfg = synth.NUSC
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 6
torch.manual_seed(0)
model = modules.build_pillarnext_b(cfg).cuda().train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
ex = bench.to_device(synth.make_batch(list(range(frames)), 30000, cfg, kind="lidar", n_boxes=40, sweeps=10), "cuda")
def step():
    loss, _ = model(ex); loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name
        if 'igemm_kernel' in n: n = 'igemm_kernel' + re.search(r'<[^>]*>', n).group(0)
        elif 'wgrad_kernel' in n: n = 'wgrad_kernel' + re.search(r'<[^>]*>', n).group(0)
        elif 'at::' in n:
            f = re.findall(r'(\w+Functor\w*|\w+_kernel_cuda|\w+_kernel_impl\w*|\w+Op\b)', n)
            n = 'torch:' + (f[0] if f else n[:40])
        else:
            m = re.search(r'(\w+_kernel)', n); n = m.group(1) if m else n[:50]
        agg[n][0] += 1; agg[n][1] += e.device_time
tot = sum(v[1] for v in agg.values())
mine = sum(v[1] for k, v in agg.items() if not k.startswith('torch:') and 'Memset' not in k and 'Memcpy' not in k)
out = ["# torch.profiler (kineto) device-time table of ONE steady-state training step (%d frames x 30k pts), round 2" % frames,
       "# %d device activities, %.3f ms summed device time; libpnx kernels %.1f%%" % (sum(v[0] for v in agg.values()), tot / 1e3, 100 * mine / tot)]
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append("%9.3f ms %5.1f%% %5d  %s" % (t / 1e3, 100 * t / tot, c, k[:70]))
open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "step_kernels.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:45]))
