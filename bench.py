#!/usr/bin/env python
"""bench.py -- frames/sec of the per-frame dense-SLAM hot path (BASELINE.json `metric`).

A "step" is one frame of a synthetic 640x480 depth stream through the whole hot path:
mm2meters -> block allocation -> TSDF integration (+ node update) -> raycast -> renderVolume (reuse
path), i.e. DenseSLAMSystem::{preprocessing, integration, raycasting, renderVolume} with poses
supplied (tracking excluded on both sides).

  python bench.py [--gpus N --steps K --warmup W]        our CUDA path (one process per GPU)
  python bench.py --impl reference ...                   the reference's own CPU code on the host cores
                                                         (oracle/_ref, OpenMP)

The line's headline workload is BASELINE.json configs[1] (planar_sweep_sdf512).  At N = 1 the line also carries
`extra_workloads`: short legs of configs[2] (OFusion 1024^3) and configs[3] (SDF 2048^3, the HBM-bound one) with
their per-kernel times and roofline fractions, so that those numbers are driver-run too.

N>1 (launched by torch.distributed.run): one independent sequence + map per GPU ("replicas only",
SURVEY.md 8(e)); NCCL is used for the barrier and to gather timings, nothing else.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec (integrate+raycast) 640x480 -> 512^3 TSDF octree"
UNIT = "frames/s"
WORKLOADS = {
    # BASELINE.json configs[1]: synthetic planar sweep, SDF 512^3 -- the configuration the metric is quoted on
    "planar_sweep_sdf512": dict(field=0, size=512, dim=4.8, mu=0.1, scene="plane", W=640, H=480),
    # configs[2] / configs[3]: reported as `extra_workloads` of the default line, selectable as the main workload for profiling runs
    "box_room_ofusion1024": dict(field=1, size=1024, dim=4.8, mu=0.008, scene="room", W=640, H=480),
    "box_room_sdf2048": dict(field=0, size=2048, dim=4.096, mu=0.1, scene="room", W=640, H=480, max_blocks=1 << 22),   # a full turn of the room at 2048^3 allocates ~3 M blocks (12 GB)
    "planar_sweep_sdf256_small": dict(field=0, size=256, dim=4.8, mu=0.1, scene="plane", W=160, H=120),
}
EXTRA = ("box_room_ofusion1024", "box_room_sdf2048")
K_CAM = (481.2, 480.0, 320.0, 240.0)
NOISE_MM, DROPOUT = 2.0, 0.01
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback
STAGES = ("alloc", "fuse", "raycast", "render")


def camera_for(cfg):
    s = cfg["W"] / 640.0
    return tuple(v * s for v in K_CAM)


def make_frames(cfg, n, seq):
    """frames 0 .. n-1 of sequence `seq` (BASELINE.json config 5: noise seed 1234 + seq, supereight_b200/synth.py)"""
    from supereight_b200 import synth
    gen = synth.planar_sweep if cfg["scene"] == "plane" else synth.box_room
    k = camera_for(cfg)
    depth = np.empty((n, cfg["H"], cfg["W"]), np.uint16)
    poses = np.empty((n, 4, 4), np.float32)
    for f in range(n):
        depth[f], poses[f] = gen(f, cfg["dim"], cfg["W"], cfg["H"], k, noise_mm=NOISE_MM, dropout=DROPOUT, seed=seq)
    return depth, poses, k


def config_of(name, cfg, world=1):
    """what both arms echo (the driver compares them)"""
    return {"workload": name, "field": "SDF" if cfg["field"] == 0 else "OFusion", "volume": f"{cfg['size']}^3 @ {cfg['dim']} m",
            "image": f"{cfg['W']}x{cfg['H']}", "mu": cfg["mu"], "scene": cfg["scene"], "noise_mm": NOISE_MM, "dropout": DROPOUT,
            "noise_seed": "1234 + g for sequence g (one sequence per GPU)",
            "stages": "mm2meters+alloc+integrate+raycast+renderVolume(reuse)", "poses": "supplied (no tracking)",
            "parallelism": f"replicas x{world} (one map per GPU)", "l2": "flushed between timed steps (256 MiB write)"}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(cfg, c, s):
    """SURVEY.md 8(d) / BASELINE.md section 4, per frame, per stage, from the unit counts of that frame."""
    vb = 8 if cfg["field"] == 0 else 16
    bb = 512 * vb
    px = cfg["W"] * cfg["H"]
    new_blocks = c["blocks"] - c["blocks_before"]
    new_nodes = c["nodes"] - c["nodes_before"]
    unique_keys = new_blocks if cfg["field"] == 0 else c["requests"]
    return {
        "alloc": px * 4 + new_blocks * bb + new_nodes * (8 * 4 + 8 + 8 * vb) + unique_keys * 8,
        "fuse": c["active"] * (2 * bb + 16) + c["nodes"] * 8 * vb * 2 + px * 4,
        "raycast": s["n_get"] * vb + s["n_interp"] * 8 * vb + s["n_grad"] * 32 * vb + px * 24,
        "render": px * (24 + 4),
    }


def measured_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch and stage, from the committed ncu capture of this workload
    (profiles/traffic.json, written by scripts/summarise_profiles.py from `ncu --set full`); None where there is none."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(tp)).get(workload)
        return t if isinstance(t, dict) else None
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 10 ms while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        if os.environ.get("SE_B200_BENCH_NO_SAMPLER"):      # (diagnostic: how much the 10 ms NVML polling disturbs the host-timed legs)
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "10"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU side: the reference's own code (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------------
def run_cpu(cfg, depth, poses, k, warmup, steps, budget_s=25.0):
    """Times the reference's CPU implementation of the path over the same frames: frames [0, warmup) warm the map up (and
    calibrate the thread count), frames warmup .. warmup+steps-1 are timed in order -- the frames the GPU arm times.
    kind "reference": oracle/_ref/libse_ref_<field>_fast.so -- the reference's own DenseSLAMSystem.cpp compiled where it lies
    (g++ -O3 -march=x86-64-v3 -fopenmp; its own build uses -O3 -march=native, but the library is built in the development
    container and has to run on the GPU box's CPU) against the stand-in Eigen / Sophus headers (oracle/Makefile); used
    whenever that build exists.  kind "port": the oracle (oracle/_build/liboracle_fast.so) otherwise.
    The thread count is calibrated (the reference's alloc pass writes block->active from every ray, which scales badly
    across sockets), the best one is used and reported.  Stops early when the time budget runs out and says how far it got."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    ref_kind = ("ref_sdf_fast", "ref_ofusion_fast")[cfg["field"]]
    kind = ref_kind if os.path.exists(oracle_lib.lib_file(ref_kind)) else "fast"
    lib = oracle_lib.load(kind)
    ncpu = os.cpu_count() or 1
    o = oracle_lib.Oracle(cfg["field"], cfg["size"], cfg["dim"], cfg["W"], cfg["H"], kind=kind)
    mu = cfg["mu"]

    def frame(f):
        o.preprocess(depth[f]); o.integrate(poses[f], k, mu, f); o.raycast(poses[f], k, mu)
        o.render_volume(poses[f], k, mu, 0.75 * mu, False)

    t_start = time.perf_counter()
    warmup = max(warmup, 1)
    w = 0                                             # warm-up frames are visited in order, again from 0 if calibration needs more
    for _ in range(min(warmup, 3)):                   # the first frames allocate most of the map
        frame(w % warmup); w += 1
    cands = sorted({c for c in (8, 12, 16, 24, 32, 48, 64, ncpu) if c <= ncpu})
    best, best_t = cands[0], float("inf")
    for c in cands:
        lib.seo_set_omp_threads(c)
        frame(w % warmup); w += 1                     # settle
        times = []
        for _ in range(3):
            t0 = time.perf_counter(); frame(w % warmup); w += 1
            times.append(time.perf_counter() - t0)
        dt = sorted(times)[1]                         # median of 3
        if dt < best_t:
            best, best_t = c, dt
    lib.seo_set_omp_threads(best)
    while w < warmup and time.perf_counter() - t_start < budget_s * 0.4:
        frame(w); w += 1
    done, per = 0, []
    t0 = time.perf_counter()
    while done < steps and (time.perf_counter() - t0) < budget_s * 0.6:
        t1 = time.perf_counter(); frame(warmup + done); per.append(time.perf_counter() - t1); done += 1
    dt = time.perf_counter() - t0
    fps = done / dt if dt > 0 else 0.0
    what = "the reference's own sources (oracle/_ref, stand-in Eigen/Sophus)" if kind != "fast" else "oracle port"
    return dict(value=fps, unit=UNIT, cores=best, kind="reference" if kind != "fast" else "port",
                sample=f"frames {warmup}..{warmup + done - 1} of the same stream ({done} frames) after {w} warm-up frames, {what} with OpenMP "
                       f"({best} of {ncpu} host threads, best of {cands})",
                ms_per_step=1e3 * dt / max(done, 1), ms_per_step_median=1e3 * float(np.median(per)) if per else 0.0, steps=done)


# ------------------------------------------------------------------------------------------------
# N > 1: replicas only -- the one "collective" is a MAX over ranks of the timed durations
# ------------------------------------------------------------------------------------------------
def max_over_ranks(values_ms, world, device=None):
    """all_reduce(MAX) of per-rank durations (works on nccl with a cuda device and on gloo with cpu)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values_ms), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def aggregate_value(world, steps, max_ms):
    """whole-job throughput: units all ranks processed / the slowest rank's time"""
    return world * steps / (max_ms * 1e-3)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
class GpuLegs:
    """One workload on one GPU: the frames, buffers and bound ABI calls of the timed legs."""

    def __init__(self, name, cfg, seq, local_rank, steps, warmup, stream, flush, barrier):
        import torch
        from supereight_b200 import Map
        self.torch, self.Map = torch, Map
        self.name, self.cfg, self.steps, self.warmup = name, cfg, steps, warmup
        self.dev = torch.device("cuda", local_rank)
        self.local_rank, self.stream, self.flush, self.barrier = local_rank, stream, flush, barrier
        self.W, self.H, self.mu = cfg["W"], cfg["H"], cfg["mu"]
        self.n_frames = warmup + steps
        self.depth, self.poses, self.k = make_frames(cfg, self.n_frames, seq)
        self.poses_c = np.ascontiguousarray(self.poses, np.float32)
        self.k_c = np.ascontiguousarray(self.k, np.float32)
        self.pose_ptr = [C.c_void_p(self.poses_c.ctypes.data + 64 * f) for f in range(self.n_frames)]
        self.k_ptr = C.c_void_p(self.k_c.ctypes.data)
        self.c_mu, self.c_ls = C.c_float(self.mu), C.c_float(0.75 * self.mu)
        self.d_depth = torch.from_numpy(self.depth.view(np.int16)).to(self.dev)         # the stream, resident in HBM
        self.ddepth_ptr = [C.c_void_p(self.d_depth[f].data_ptr()) for f in range(self.n_frames)]
        self.d_rgba = torch.empty((self.H, self.W, 4), dtype=torch.uint8, device=self.dev)
        self.drgba_ptr = C.c_void_p(self.d_rgba.data_ptr())
        self.lib = None

    def new_map(self):
        m = self.Map(self.cfg["field"], self.cfg["size"], self.cfg["dim"], self.W, self.H, max_blocks=self.cfg.get("max_blocks", 0), device=self.local_rank)
        assert self.stream.cuda_stream != 0
        m.set_stream(self.stream.cuda_stream)
        self.lib = m.lib
        return m

    def check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.se_b200_last_error().decode())

    # The four ABI calls of a frame, bound with raw pointers (no per-call numpy conversion: the Python glue would otherwise
    # cost as much as a kernel).
    def step_resident(self, m, f):
        lib, h = self.lib, m.h
        self.check(lib.se_b200_preprocess_depth_device(h, self.ddepth_ptr[f], self.W, self.H))
        self.check(lib.se_b200_integrate(h, self.pose_ptr[f], self.k_ptr, self.c_mu, f))
        self.check(lib.se_b200_raycast(h, self.pose_ptr[f], self.k_ptr, self.c_mu))
        self.check(lib.se_b200_render_volume_device(h, self.drgba_ptr, self.pose_ptr[f], self.k_ptr, self.c_mu, self.c_ls, 0))

    def resident(self, render_target=True):
        """`value`: inputs resident in HBM, the image left in HBM; K steps bracketed by CUDA-event pairs on the launching
        stream, L2 flushed before each.  render_target: se_b200_set_render_target(the device image) -- renderVolume's reuse
        path is then fused into the raycast kernel (3 launches per frame instead of 4)."""
        torch = self.torch
        m = self.new_map()
        if render_target:
            self.check(self.lib.se_b200_set_render_target(m.h, self.drgba_ptr))
        for f in range(self.warmup):
            self.flush.zero_(); self.step_resident(m, f)
        self.barrier()
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(self.steps)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(self.steps)]
        launches0 = m.launch_count()
        wall0 = time.perf_counter()
        for i in range(self.steps):
            self.flush.zero_()                      # L2 flush between timed steps (outside the event pair)
            ev0[i].record(); self.step_resident(m, self.warmup + i); ev1[i].record()
        self.barrier()
        wall = time.perf_counter() - wall0
        launches = m.launch_count() - launches0
        step_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
        checksum = int(self.d_rgba.sum(dtype=torch.int64).item())
        m.close()
        return dict(step_ms=step_ms, launches=launches, wall=wall, checksum=checksum)

    def stages(self):
        """per-kernel durations (per-stage CUDA event pairs on, which costs the PDL overlap and the cross-frame overlap of the
        allocation pass: the stage times add up to more than a step of `resident`) and, for the SAME frames, the unit counts
        the algorithmic bytes are made of"""
        m = self.new_map()
        m.set_stage_timing(True)
        for f in range(self.warmup):
            self.flush.zero_(); self.step_resident(m, f)
        self.barrier()
        ms = {s: 0.0 for s in STAGES}
        by = {s: 0 for s in STAGES}
        units = dict(active=0, blocks=0, nodes=0, new_blocks=0)
        for i in range(self.steps):
            f = self.warmup + i
            self.flush.zero_()
            self.step_resident(m, f)
            for s in STAGES:                        # per-kernel device times of this step (CUDA events on the same stream)
                ms[s] += m.elapsed_ms(s)
            c = m.counters()
            smp = m.raycast_count_samples(self.poses[f], self.k, self.mu)     # the same rays once more, counting (map unchanged)
            ab = algorithmic_bytes(self.cfg, c, smp)
            for s in STAGES:
                by[s] += ab[s]
            units["active"] += c["active"]; units["new_blocks"] += c["blocks"] - c["blocks_before"]
            units["blocks"], units["nodes"] = c["blocks"], c["nodes"]
        self.barrier()
        m.close()
        n = self.steps
        return ({s: ms[s] / n for s in STAGES}, {s: by[s] / n for s in STAGES},
                dict(blocks=units["blocks"], nodes=units["nodes"], active_blocks_mean=round(units["active"] / n, 1), new_blocks_mean=round(units["new_blocks"] / n, 2)))

    def host_loop(self, buffers="pinned", render_target=False):
        """`e2e`: the synchronous per-frame loop through the C ABI with HOST buffers -- H2D of the depth frame and D2H of the
        rendered image inside the timed region, one host clock pair per step.
        buffers: "pinned" (cudaHostAlloc), "pageable" (plain numpy), "registered" (plain numpy + se_b200_register_host_buffer)."""
        torch = self.torch
        m = self.new_map()
        lib, h, W, H = self.lib, m.h, self.W, self.H
        if buffers == "pinned":
            h_depth = torch.from_numpy(self.depth.view(np.int16)).pin_memory()
            h_rgba = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
            dptr = [C.c_void_p(h_depth[f].data_ptr()) for f in range(self.n_frames)]
            out_np, optr = h_rgba.numpy(), C.c_void_p(h_rgba.data_ptr())
        else:
            h_depth = np.ascontiguousarray(self.depth)
            out_np = np.zeros((H, W, 4), np.uint8)
            dptr = [C.c_void_p(h_depth[f].ctypes.data) for f in range(self.n_frames)]
            optr = C.c_void_p(out_np.ctypes.data)
            if buffers == "registered":
                self.check(lib.se_b200_register_host_buffer(C.c_void_p(h_depth.ctypes.data), h_depth.nbytes))
                self.check(lib.se_b200_register_host_buffer(optr, out_np.nbytes))
        if render_target:
            self.check(lib.se_b200_set_render_target(h, optr))

        def step(f):
            self.check(lib.se_b200_preprocess_depth_host(h, dptr[f], W, H))          # cudaMemcpyAsync H2D (+ mm2meters inside the allocation kernel)
            self.check(lib.se_b200_integrate(h, self.pose_ptr[f], self.k_ptr, self.c_mu, f))
            self.check(lib.se_b200_raycast(h, self.pose_ptr[f], self.k_ptr, self.c_mu))
            self.check(lib.se_b200_render_volume_host(h, optr, self.pose_ptr[f], self.k_ptr, self.c_mu, self.c_ls, 0))   # D2H + sync

        try:
            for f in range(self.warmup):
                self.flush.zero_(); step(f)
            self.barrier()
            per = []
            for i in range(self.steps):
                self.flush.zero_(); torch.cuda.synchronize()
                t0 = time.perf_counter(); step(self.warmup + i); per.append(time.perf_counter() - t0)
            self.barrier()
            checksum = int(out_np.astype(np.uint64).sum())
        finally:
            m.close()
            if buffers == "registered":
                lib.se_b200_unregister_host_buffer(C.c_void_p(h_depth.ctypes.data)); lib.se_b200_unregister_host_buffer(optr)
        return dict(total_s=float(np.sum(per)), median_ms=1e3 * float(np.median(per)), mean_ms=1e3 * float(np.mean(per)), checksum=checksum)

    def host_overlapped(self):
        """the same frames through the *_host_async calls (upload / download on the map's copy streams, double-buffered: the
        copies of neighbouring frames overlap the kernels); K frames issued back to back, one synchronisation at the end; no L2
        flush -- the 300-frame input stream is 184 MB"""
        torch = self.torch
        m = self.new_map()
        lib, h, W, H = self.lib, m.h, self.W, self.H
        h_depth = torch.from_numpy(self.depth.view(np.int16)).pin_memory()
        dptr = [C.c_void_p(h_depth[f].data_ptr()) for f in range(self.n_frames)]
        h_out = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
        optr = [C.c_void_p(x.data_ptr()) for x in h_out]

        def step(f, i):
            self.check(lib.se_b200_preprocess_depth_host_async(h, dptr[f], W, H))
            self.check(lib.se_b200_integrate(h, self.pose_ptr[f], self.k_ptr, self.c_mu, f))
            self.check(lib.se_b200_raycast(h, self.pose_ptr[f], self.k_ptr, self.c_mu))
            self.check(lib.se_b200_render_volume_host_async(h, optr[i & 1], self.pose_ptr[f], self.k_ptr, self.c_mu, self.c_ls, 0))

        try:
            for f in range(self.warmup):
                step(f, f)
            self.check(lib.se_b200_sync(h))
            self.barrier()
            t0 = time.perf_counter()
            for i in range(self.steps):
                step(self.warmup + i, i)
            self.check(lib.se_b200_sync(h))
            dt = time.perf_counter() - t0
            checksum = int(h_out[(self.steps - 1) & 1].numpy().astype(np.uint64).sum())
        finally:
            m.close()
        return dict(total_s=dt, checksum=checksum)


def kernel_table(name, stage_ms, stage_bytes, peak):
    """per stage: CUDA-event time, algorithmic bytes, `frac` = algorithmic bytes / time / peak, and -- where an
    ncu capture of this workload is committed -- `traffic` = DRAM bytes per launch and `dram_frac` = traffic / time / peak.
    At 512^3 the map is L2-resident, so dram_frac << frac there: `frac` is a rate of USEFUL bytes, `dram_frac` the HBM load."""
    traffic = measured_traffic(name) or {}
    out = {}
    for s in STAGES:
        ms = stage_ms[s]
        gbs = stage_bytes[s] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        e = {"ms": round(ms, 5), "algorithmic_bytes": int(stage_bytes[s]), "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
        t = traffic.get(s)
        if isinstance(t, (int, float)) and ms > 0:
            e["traffic"] = int(t)
            e["dram_frac"] = round(t / (ms * 1e-3) / 1e9 / peak, 4)
        out[s] = e
    return out


def run_gpu(args, cfg, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # all work (ours and torch's flush/copies) goes to one explicit, non-default stream, so the CUDA events
    # recorded through torch bracket exactly the kernels the library launches
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    steps, warmup = args.steps, max(args.warmup, 3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    legs = GpuLegs(args.workload, cfg, rank, local_rank, steps, warmup, stream, flush, barrier)      # sequence g = rank: noise seed 1234 + g
    sampler = ClockSampler(local_rank)
    sampler.start()
    res = legs.resident(render_target=True)
    res_plain = legs.resident(render_target=False)
    stage_ms, stage_bytes, units = legs.stages()
    # e2e: the four stock calls on the caller's own (malloc'd) buffers, page-locked once with se_b200_register_host_buffer -- what
    # se_b200_benchmark.cpp does with the buffers se_apps/src/benchmark.cpp:90-97 allocates.  (Round 2: the loop with the render
    # target -- the raycast kernel writing the image to host memory as the rays finish -- is measured beside it.  It does not win
    # any more: the 1.2 MB image needs ~22 us of PCIe time, which a 31 us raycast that starts its expensive rays first no longer
    # hides.  And buffers from torch's pinned allocator, the first headline, are measured beside it too: over a run's first frames
    # they are slower and jittery on these boxes -- 0.19 against 0.142 ms per frame over frames 5..24, the same over 300 frames;
    # scripts/e2e_order.py, profiles/r2b_e2e_order.log.)
    e2e = legs.host_loop("registered", render_target=False)
    e2e_rt = legs.host_loop("registered", render_target=True)
    clocks = sampler.stop()                 # sampled over the timed regions above (resident and end-to-end)
    notes = {}

    def optional(label, fn):
        try:
            return fn()
        except Exception as e:               # an extra measurement must never take the bench line down
            notes[label] = f"{type(e).__name__}: {e}"
            return None

    e2e_pageable = optional("e2e_pageable", lambda: legs.host_loop("pageable", render_target=False))
    e2e_torch_pinned = optional("e2e_torch_pinned", lambda: legs.host_loop("pinned", render_target=False))
    ov = optional("e2e_overlapped", legs.host_overlapped)
    barrier()

    # ---------------- aggregate: max over ranks ----------------
    total_ms = sum(res["step_ms"])
    ov_ms = ov["total_s"] * 1e3 if ov else 0.0
    total_ms_max, plain_ms_max, e2e_ms_max, e2e_rt_ms_max, ov_ms_max, ov_failed_any = max_over_ranks(
        [total_ms, sum(res_plain["step_ms"]), e2e["total_s"] * 1e3, e2e_rt["total_s"] * 1e3, ov_ms, 0.0 if ov else 1.0], world, dev)
    result = None
    if rank == 0:
        peak, peak_src = peak_hbm()
        kernels = kernel_table(args.workload, stage_ms, stage_bytes, peak)
        dom = max(STAGES, key=lambda s: kernels[s]["ms"])
        frame_bytes = sum(stage_bytes.values())
        frame_ms = sum(kernels[s]["ms"] for s in STAGES)
        images_agree = len({res["checksum"], res_plain["checksum"], e2e["checksum"], e2e_rt["checksum"]}) == 1
        conf = config_of(args.workload, cfg, world)
        W, H = cfg["W"], cfg["H"]

        def e2e_entry(r, api, ms_max=None, extra=None):
            if r is None:
                return None
            ms = (ms_max if ms_max is not None else r["total_s"] * 1e3) / steps
            d = {"value": round((world if ms_max is not None else 1) * 1e3 / ms, 2), "unit": UNIT, "ms_per_step": round(ms, 5),
                 "ms_per_step_median": round(r["median_ms"], 5) if "median_ms" in r else None, "result_checksum": r["checksum"], "api": api}
            if ms_max is None:
                d["scope"] = "rank 0"
            d.update(extra or {})
            return d

        result = {
            "metric": METRIC, "value": round(aggregate_value(world, steps, total_ms_max), 2), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": round(total_ms_max / steps, 5),
            "ms_per_step_median": round(float(np.median(res["step_ms"])), 5),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": conf,
            "workload_stats": units,
            "api": "se_b200_preprocess_depth_device + se_b200_integrate + se_b200_raycast + se_b200_render_volume_device per frame, "
                   "se_b200_set_render_target(device image) set once: renderVolume's reuse path runs inside the raycast kernel",
            "value_without_render_target": {"value": round(aggregate_value(world, steps, plain_ms_max), 2), "ms_per_step": round(plain_ms_max / steps, 5),
                                            "gpu_launches": int(res_plain["launches"]), "note": "the same loop with the separate shading kernel"},
            "e2e": e2e_entry(e2e, "synchronous se_b200_preprocess_depth_host, se_b200_integrate, se_b200_raycast, se_b200_render_volume_host per frame (the "
                                  "reference's stage semantics) on host buffers page-locked once with se_b200_register_host_buffer (pinned host "
                                  "memory): depth copied in by preprocess, the image written to the caller's buffer by renderVolume's shading "
                                  "kernel", e2e_ms_max, {"h2d_bytes_per_step": W * H * 2, "d2h_bytes_per_step": W * H * 4}),
            "e2e_render_target": e2e_entry(e2e_rt, "the same loop with se_b200_set_render_target(the output buffer) set once: the raycast kernel writes "
                                                   "the image to host memory as the rays finish, renderVolume waits for it", e2e_rt_ms_max),
            "e2e_pageable": e2e_entry(e2e_pageable, "the same loop with malloc'd buffers, as se_apps/src/benchmark.cpp:90-97 allocates them (staged copies)")
                            or {"unavailable": notes.get("e2e_pageable")},
            "e2e_torch_pinned": e2e_entry(e2e_torch_pinned, "the same loop on buffers from torch's pinned allocator (tensor.pin_memory())")
                                or {"unavailable": notes.get("e2e_torch_pinned")},
            "e2e_overlapped": ({"value": round(aggregate_value(world, steps, ov_ms_max), 2), "unit": UNIT, "ms_per_step": round(ov_ms_max / steps, 5),
                                "result_checksum": ov["checksum"],
                                "api": "se_b200_preprocess_depth_host_async + se_b200_render_volume_host_async (copy streams, double-buffered); "
                                       "same bytes per step as e2e, frames issued back to back, one synchronisation at the end, no L2 flush"}
                               if ov and ov_failed_any == 0 else {"unavailable": notes.get("e2e_overlapped", "failed on another rank")}),
            "images_agree": images_agree,
            "gpu_launches": int(res["launches"]),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                         "frac": kernels[dom]["frac"], "traffic": kernels[dom].get("traffic"), "dram_frac": kernels[dom].get("dram_frac"),
                         "peak_source": peak_src,
                         "what": "achieved / frac: ALGORITHMIC bytes (SURVEY.md 8d, from the unit counts of the timed frames) / CUDA-event time of the stage; "
                                 "traffic / dram_frac: DRAM bytes of the committed ncu capture (profiles/traffic.json) over the same time. "
                                 "The 512^3 map (~40 MB) lives in L2, so its kernels are bound by issue / latency, not by HBM: see extra_workloads.box_room_sdf2048",
                         "frame": {"algorithmic_bytes": int(frame_bytes), "achieved_gbs": round(frame_bytes / (frame_ms * 1e-3) / 1e9, 1),
                                   "frac": round(frame_bytes / (frame_ms * 1e-3) / 1e9 / peak, 4), "ms_sum_of_stages": round(frame_ms, 5)},
                         "kernels": kernels},
            "wall_s_timed_region": round(res["wall"], 3),
        }
        if not images_agree:
            result["images_agree_note"] = {"resident": res["checksum"], "resident_plain": res_plain["checksum"], "e2e": e2e["checksum"], "e2e_render_target": e2e_rt["checksum"]}
    del legs
    torch.cuda.empty_cache()

    # ---------------- extra workloads (N = 1 only): configs[2] and configs[3], short legs ----------------
    if rank == 0 and world == 1 and not args.no_extra and args.workload == "planar_sweep_sdf512":
        peak, _ = peak_hbm()
        extra = {}
        for name in EXTRA:
            try:
                xcfg = WORKLOADS[name]
                xl = GpuLegs(name, xcfg, 0, local_rank, args.extra_steps, 5, stream, flush, barrier)
                r = xl.resident(render_target=True)
                sms, sby, xunits = xl.stages()
                xe = xl.host_loop("registered", render_target=False)
                k = kernel_table(name, sms, sby, peak)
                xconf = config_of(name, xcfg, 1)
                ms = sum(r["step_ms"]) / len(r["step_ms"])
                extra[name] = {"value": round(1e3 / ms, 2), "unit": UNIT, "steps": args.extra_steps, "warmup": 5, "ms_per_step": round(ms, 5),
                               "e2e": {"value": round(args.extra_steps / xe["total_s"], 2), "ms_per_step": round(xe["mean_ms"], 5)},
                               "gpu_launches": int(r["launches"]), "config": xconf, "workload_stats": xunits,
                               "roofline": {"bound": "hbm", "kernel": max(STAGES, key=lambda s: k[s]["ms"]), "peak": peak, "kernels": k}}
                del xl
                torch.cuda.empty_cache()
            except Exception as e:
                extra[name] = {"unavailable": f"{type(e).__name__}: {e}"}
        result["extra_workloads"] = extra

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        depth, poses, k = make_frames(cfg, min(warmup, 5) + min(steps, 40), 0)
        cb = run_cpu(cfg, depth, poses, k, warmup=min(warmup, 5), steps=min(steps, 40), budget_s=25.0)
        cb.pop("ms_per_step", None); cb.pop("steps", None); cb.pop("ms_per_step_median", None)
        cb["value"] = round(cb["value"], 3)
        result["cpu_baseline"] = cb
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)    # S1 is a 300-frame sweep
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--workload", default="planar_sweep_sdf512", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_workloads legs (configs[2], configs[3])")
    ap.add_argument("--extra-steps", type=int, default=30)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = WORKLOADS[args.workload]

    if args.impl == "reference":
        # The reference's own CPU implementation of the path: oracle/_ref (the reference's sources compiled against
        # stand-in Eigen / Sophus headers) with OpenMP on the host cores; the oracle port only where that
        # build is absent.  Rank 0 alone works; the other ranks exit 0.  Same frames as the GPU arm: sequence 0, frames
        # [0, warmup) warm up, frames warmup .. warmup+steps-1 are timed (as many as fit the time budget).
        if rank != 0:
            return
        warmup = max(args.warmup, 3)
        depth, poses, k = make_frames(cfg, warmup + args.steps, 0)
        r = run_cpu(cfg, depth, poses, k, warmup=warmup, steps=args.steps, budget_s=150.0)
        line = {
            "impl": "reference", "metric": METRIC, "value": round(r["value"], 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": warmup, "ms_per_step": round(r["ms_per_step"], 3), "ms_per_step_median": round(r["ms_per_step_median"], 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args.workload, cfg, args.gpus),
            "cpu_baseline": {"value": round(r["value"], 3), "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": round(r["value"], 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return

    result = run_gpu(args, cfg, rank, world, local_rank)
    if rank == 0:
        print(json.dumps(result), flush=True)


if __name__ == "__main__":
    main()
