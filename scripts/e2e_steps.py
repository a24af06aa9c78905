"""Per-step host times of the synchronous end-to-end loop (bench.py's `e2e` leg), with and without the render target.
Usage: python scripts/e2e_steps.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
cfg = bench.WORKLOADS["planar_sweep_sdf512"]
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
legs = bench.GpuLegs("planar_sweep_sdf512", cfg, 0, 0, steps, 10, stream, flush, torch.cuda.synchronize)
import time


def loop(buffers, rt):
    # bench.GpuLegs.host_loop, keeping the per-step times
    r = legs.host_loop(buffers, render_target=rt)
    return r


for buffers, rt in (("pinned", True), ("pinned", False), ("pinned", True), ("pinned", False), ("registered", True)):
    r = legs.host_loop(buffers, render_target=rt)
    print(f"{buffers:10s} render_target={rt}: mean {r['mean_ms']:.4f} ms  median {r['median_ms']:.4f} ms  -> {1e3 / r['mean_ms']:.0f} / {1e3 / r['median_ms']:.0f} frames/s")
