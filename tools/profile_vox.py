import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pillarnext_b200 import ops, synth
cfg = synth.NUSC
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
base = [synth.make_frame(5000 + i, 30000, cfg, "lidar", sweeps=10) for i in range(16)]
pts = synth.collate_points([base[i % 16] for i in range(frames)]).cuda()
for _ in range(3):
    v = ops.voxelize(pts, frames, cfg["voxel_size"], cfg["pc_range"], buckets=True)
torch.cuda.synchronize()
