"""Per-step host times of the synchronous end-to-end loop for the first frames of a run (bench.py's `e2e` leg keeps only the
mean and the median).  Usage: python scripts/e2e_trace.py [steps] [warmup]"""
import os
import sys
import time
import ctypes as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
warmup = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = bench.WORKLOADS["planar_sweep_sdf512"]
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
legs = bench.GpuLegs("planar_sweep_sdf512", cfg, 0, 0, steps, warmup, stream, flush, torch.cuda.synchronize)
for rep in range(2):
    m = legs.new_map()
    lib, h, W, H = legs.lib, m.h, legs.W, legs.H
    h_depth = torch.from_numpy(legs.depth.view(np.int16)).pin_memory()
    h_rgba = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
    dptr = [C.c_void_p(h_depth[f].data_ptr()) for f in range(legs.n_frames)]
    optr = C.c_void_p(h_rgba.data_ptr())
    per, parts = [], []
    for f in range(warmup + steps):
        flush.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        lib.se_b200_preprocess_depth_host(h, dptr[f], W, H); t1 = time.perf_counter()
        lib.se_b200_integrate(h, legs.pose_ptr[f], legs.k_ptr, legs.c_mu, f); t2 = time.perf_counter()
        lib.se_b200_raycast(h, legs.pose_ptr[f], legs.k_ptr, legs.c_mu); t3 = time.perf_counter()
        lib.se_b200_render_volume_host(h, optr, legs.pose_ptr[f], legs.k_ptr, legs.c_mu, legs.c_ls, 0); t4 = time.perf_counter()
        per.append(1e6 * (t4 - t0)); parts.append([1e6 * (t1 - t0), 1e6 * (t2 - t1), 1e6 * (t3 - t2), 1e6 * (t4 - t3)])
    print(f"rep {rep}: per-step us:", " ".join(f"{p:.0f}" for p in per))
    p = np.array(parts)
    for lo, hi in ((warmup, warmup + 20), (warmup + 20, warmup + steps)):
        print(f"  frames {lo}..{hi - 1}: mean {np.mean(per[lo:hi]):.1f} us; host time in preprocess / integrate / raycast / render calls: "
              + " / ".join(f"{v:.1f}" for v in p[lo:hi].mean(axis=0)), "counters", m.counters()["active"] if hi == warmup + steps else "")
    m.close()
