"""Scratch: a few pnx_voxelize / pnx_voxelize_frames calls for ncu.  argv: frames [config] [frames|global]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pillarnext_b200 import ops, synth
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sel = synth.BENCH_CONFIGS[sys.argv[2] if len(sys.argv) > 2 else "nusc"]
tiled = (sys.argv[3] if len(sys.argv) > 3 else "frames") == "frames" and "force"
cfg, n = sel["cfg"], sel["points"]
nbase = 16 if n <= 60000 else 2
base = [synth.make_frame(5000 + i, n, cfg, "lidar", sweeps=sel["sweeps"], rings=sel["rings"]) for i in range(nbase)]
pts = synth.collate_points([base[i % nbase] for i in range(frames)]).cuda()
for _ in range(3):
    v = ops.voxelize(pts, frames, cfg["voxel_size"], cfg["pc_range"], buckets=False, frame_sorted=tiled)
torch.cuda.synchronize()
v.check_order()
