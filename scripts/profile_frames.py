"""Minimal frame loop for ncu: N frames of a bench workload through the C ABI (no timing, no oracle).
Usage: python scripts/profile_frames.py [workload] [frames] [first_frame_count_for_warmup]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench
from supereight_b200 import Map

name = sys.argv[1] if len(sys.argv) > 1 else "planar_sweep_sdf512"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cfg = bench.WORKLOADS[name]
depth, poses, k = bench.make_frames(cfg, frames, 0)
m = Map(cfg["field"], cfg["size"], cfg["dim"], cfg["W"], cfg["H"], max_blocks=cfg.get("max_blocks", 0))
out = np.empty((cfg["H"], cfg["W"], 4), np.uint8)
mu = cfg["mu"]
for f in range(frames):
    m.preprocess(depth[f])
    m.integrate(poses[f], k, mu, f)
    m.raycast(poses[f], k, mu)
    m.render_volume(poses[f], k, mu, 0.75 * mu, False, out=out)
print("frames", frames, m.counters(), "launches", m.launch_count())
