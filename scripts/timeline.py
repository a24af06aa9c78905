"""Phase time stamps of the integrate kernel, per CTA (SE_TIMELINE build of the library: supereight_b200/variants/libse_b200_timeline.so,
`nvcc ... -DSE_TIMELINE`).  Usage: SE_B200_LIB=<that build> python scripts/timeline.py [workload] [frames]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench
from supereight_b200 import Map, capi

name = sys.argv[1] if len(sys.argv) > 1 else "box_room_sdf2048"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 22
cfg = bench.WORKLOADS[name]
depth, poses, k = bench.make_frames(cfg, frames, 0)
m = Map(cfg["field"], cfg["size"], cfg["dim"], cfg["W"], cfg["H"], max_blocks=cfg.get("max_blocks", 0))
lib = capi.load_library()
lib.se_b200_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
mu = cfg["mu"]
for f in range(frames):
    m.preprocess(depth[f])
    m.integrate(poses[f], k, mu, f)
    if f == frames - 1:
        n = 1200
        buf = np.zeros(16 * n, np.uint64)
        assert lib.se_b200_debug_timeline(m.h, buf.ctypes.data, 16 * n) == 0
        t = buf.reshape(n, 16)[:592, :5].astype(np.int64)
        na = 1200
        ta = buf.reshape(-1, 16)[:na, 8:10].astype(np.int64)
        ta = (ta - ta[:, 0].min()) / 1000.0
        da = ta[:, 1] - ta[:, 0]
        print(f"alloc: {na} CTAs, kernel {ta[:, 1].max():.1f} us; CTA duration min {da.min():.1f} median {np.median(da):.1f} p90 {np.percentile(da, 90):.1f} max {da.max():.1f}; starts p50 {np.median(ta[:, 0]):.1f} max {ta[:, 0].max():.1f}")
        print("  mean CTA duration by eighth of the grid:", " ".join(f"{da[b].mean():.1f}" for b in np.array_split(np.arange(na), 8)))
        for back in (2, 5, 10, 20):
            print(f"  CTAs running {back:2d} us before the end: {int(((ta[:, 0] <= ta[:, 1].max() - back) & (ta[:, 1] > ta[:, 1].max() - back)).sum())}")
        t0 = t[:, 0].min()
        t = (t - t0) / 1000.0
        names = ["start", "list produced", "blocks done (warp 0)", "list complete", "nodes done"]
        for i, nm in enumerate(names):
            print(f"{nm:24s} min {t[:, i].min():8.1f}  median {np.median(t[:, i]):8.1f}  max {t[:, i].max():8.1f} us")
    m.raycast(poses[f], k, mu)
    if f == frames - 1:
        n = int(os.environ.get('SE_RAY_CTAS', '2400'))
        buf = np.zeros(16 * n, np.uint64)
        assert lib.se_b200_debug_timeline(m.h, buf.ctypes.data, 16 * n) == 0
        t = buf.reshape(n, 16)[:, 5:7].astype(np.int64)
        t = (t - t[:, 0].min()) / 1000.0
        dur = t[:, 1] - t[:, 0]
        end = t[:, 1].max()
        print(f"raycast: {n} CTAs, kernel {end:.1f} us; CTA duration (thread 0) min {dur.min():.1f} median {np.median(dur):.1f} p90 {np.percentile(dur, 90):.1f} max {dur.max():.1f} us")
        print("  CTA starts: p50 %.1f p90 %.1f max %.1f us" % (np.median(t[:, 0]), np.percentile(t[:, 0], 90), t[:, 0].max()))
        bands = np.array_split(np.arange(n), 8)
        print("  mean CTA duration by eighth of the grid (top of the image first):", " ".join(f"{dur[b].mean():.1f}" for b in bands))
        print("  longest 5 CTAs: ", [(int(i), round(float(dur[i]), 1), round(float(t[i, 0]), 1)) for i in np.argsort(dur)[-5:]], "(index, duration, start)")
        for back in (2, 4, 6, 8, 10, 15):
            running = int(((t[:, 0] <= end - back) & (t[:, 1] > end - back)).sum())
            print(f"  CTAs running {back:2d} us before the end: {running}")
print(m.counters())
