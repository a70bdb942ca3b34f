"""Voxelizer roofline sweep for every bench config (run under gpurun): prints bench.voxelize_roofline()."""
import json, sys, torch
sys.path.insert(0, ".")
import bench
from pillarnext_b200 import synth
pk = bench.peaks()
dev = torch.device("cuda", 0)
for name in (sys.argv[1:] or ["nusc", "waymo180k", "waymo540k"]):
    r = bench.voxelize_roofline(dev, synth.BENCH_CONFIGS[name], pk)
    print(name, json.dumps({"frac": round(r["frac"], 3), "sweep": [(s["frames_per_launch"], s["points"], round(s["algorithmic_bytes"] / 1e6, 1), round(s["ms"] * 1e3, 1), round(s["frac"], 3)) for s in r["sweep"]]}))

