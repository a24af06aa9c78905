"""N4 timing: marching cubes on the device (se_b200_extract_mesh + download) against the oracle port on the host.
Usage: python scripts/mesh_timing.py [frames]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C

import numpy as np

from supereight_b200 import Map, synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 30
k = synth.DEFAULT_K
for name, size, scene in (("planar_sweep 512^3", 512, synth.planar_sweep), ("box_room 1024^3", 1024, synth.box_room)):
    W, H, dim, mu = 640, 480, 4.8, 0.1
    g = Map(0, size, dim, W, H)
    data = [scene(f * (1 if size == 512 else 10), dim, W, H, k, noise_mm=2.0, dropout=0.01) for f in range(frames)]
    for f, (d, pose) in enumerate(data):
        g.preprocess(d); g.integrate(pose, k, mu, f)
    g.sync()
    n = C.c_int64()
    g.lib.se_b200_extract_mesh(g.h, C.byref(n))            # warm-up: table upload, buffer allocation
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        g.lib.se_b200_extract_mesh(g.h, C.byref(n))
    t1 = time.perf_counter()
    tri = g.mesh()
    t2 = time.perf_counter()
    print(f"{name}: {g.block_count()} blocks -> {n.value} triangles; extract {1e3 * (t1 - t0) / reps:.3f} ms (sort + count + scan + write, with host syncs); "
          f"extract + download of {tri.nbytes / 1e6:.1f} MB {1e3 * (t2 - t1):.2f} ms")
    if size == 512:
        import mc_table_ref
        from oracle_lib import Oracle
        o = Oracle(0, size, dim, W, H)
        for f, (d, pose) in enumerate(data):
            o.preprocess(d); o.integrate(pose, k, mu, f)
        t3 = time.perf_counter()
        want = o.marching_cube(mc_table_ref.table())
        t4 = time.perf_counter()
        same = want.shape == tri.shape and np.array_equal(want.view(np.uint32), tri.view(np.uint32))
        print(f"  oracle port (serial, host): {1e3 * (t4 - t3):.1f} ms for {len(want)} triangles; identical to the device mesh: {same}")
