"""bench.py's synchronous end-to-end leg (20 steps after 5 warm-up frames, as the driver runs it) for a sequence of buffer kinds in ONE
process: which host buffers the loop is fast and steady on, and whether the order of the legs matters.
Usage: python scripts/e2e_order.py pinned,registered,pinned,resident,sleep,...   (pinned = torch's pinned allocator, registered =
numpy + se_b200_register_host_buffer, pageable; resident / stages = the device-timed legs; sleep = half a second of nothing)"""
import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
cfg = bench.WORKLOADS["planar_sweep_sdf512"]
torch.cuda.set_device(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
legs = bench.GpuLegs("planar_sweep_sdf512", cfg, 0, 0, 20, 5, stream, flush, torch.cuda.synchronize)
order = sys.argv[1].split(",")
for leg in order:
    if leg == "resident":
        r = legs.resident(render_target=True); print("resident", round(np.mean(r["step_ms"]), 4)); continue
    if leg == "stages":
        legs.stages(); print("stages"); continue
    if leg == "sleep":
        time.sleep(0.5); print("sleep"); continue
    r = legs.host_loop(leg, render_target=False)
    print(f"{leg:10s} mean {r['mean_ms']:.4f} median {r['median_ms']:.4f}")
