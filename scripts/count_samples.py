"""Per-pixel sample statistics of the raycast on a bench workload (measurement helper)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from supereight_b200 import Map
name = sys.argv[1] if len(sys.argv) > 1 else "planar_sweep_sdf512"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cfg = bench.WORKLOADS[name]
depth, poses, k = bench.make_frames(cfg, frames, 0)
m = Map(cfg["field"], cfg["size"], cfg["dim"], cfg["W"], cfg["H"], max_blocks=cfg.get("max_blocks", 0))
for f in range(frames):
    m.preprocess(depth[f]); m.integrate(poses[f], k, cfg["mu"], f)
s = m.raycast_count_samples(poses[frames - 1], k, cfg["mu"])
px = cfg["W"] * cfg["H"]
print(name, {kk: round(v / px, 2) for kk, v in s.items()}, "per pixel;", m.counters())
