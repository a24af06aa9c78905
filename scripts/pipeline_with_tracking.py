"""Closed-loop run of the whole pipeline including N1 (bilateral filter + ICP tracking) on the corner-view scene:
per-stage wall times and the drift of the tracked pose.  Usage: python scripts/pipeline_with_tracking.py [frames]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from supereight_b200 import Map, synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 60
W, H, size, dim, mu = 640, 480, 512, 4.8, 0.1
k = synth.DEFAULT_K
it = [10, 5, 4]
data = [synth.corner_view(f, dim, W, H, k, noise_mm=1.0, dropout=0.005) for f in range(frames)]
g = Map(0, size, dim, W, H)
g.set_stage_timing(False)
out = np.empty((H, W, 4), np.uint8)
est = None
t = {"preprocess+filter": 0.0, "track": 0.0, "integrate": 0.0, "raycast": 0.0, "render": 0.0}
errs, n_tracked, counted = [], 0, 0
for f, (d, gt) in enumerate(data):
    t0 = time.perf_counter(); g.preprocess(d); g.filter_depth(True, 3); g.sync(); t1 = time.perf_counter()
    if f < 4:
        est = gt.copy()
    else:
        est, ok = g.track(est, rp, k, 1e-5, it); n_tracked += ok
    t2 = time.perf_counter(); g.integrate(est, k, mu, f); g.sync(); t3 = time.perf_counter()
    g.raycast(est, k, mu); g.sync(); t4 = time.perf_counter()
    g.render_volume(est, k, mu, 0.75 * mu, False, out=out); t5 = time.perf_counter()
    rp = est.copy()
    if f >= 10:
        counted += 1
        for name, dt in zip(t, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            t[name] += dt
        errs.append(float(np.abs(est[:3, 3] - gt[:3, 3]).max()))
tot = sum(t.values())
print(f"{frames} frames 640x480 SDF 512^3 (corner view), tracked {n_tracked}/{frames - 4}; position error max {max(errs)*1e3:.1f} mm, final {errs[-1]*1e3:.1f} mm")
print("wall ms/frame:", {n: round(1e3 * v / counted, 3) for n, v in t.items()}, "total", round(1e3 * tot / counted, 3), "->", round(counted / tot, 1), "frames/s (with H2D, D2H and ICP host solves)")
