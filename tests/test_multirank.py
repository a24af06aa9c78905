"""The N>1 path on CPU (gloo, world_size 2): replicas only, so the only cross-rank step is the MAX over ranks
of the timed durations; and the reference arm under a multi-rank launch (rank 0 works, others exit 0)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def torchrun(args, port, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_max_over_ranks_gloo_world2():
    r = torchrun([os.path.join(ROOT, "tests", "_gloo_worker.py")], 29631)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2
    assert d["max_ms"] == [150.0, 300.0]                       # the slowest rank decides
    assert abs(d["value"] - 2 * 20 / 0.150) < 1e-9             # whole-job units / slowest time


def test_reference_arm_under_multirank_launch():
    r = torchrun([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3",
                  "--workload", "planar_sweep_sdf256_small"], 29632)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1                                      # rank 0 alone prints
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0      # "reference" where oracle/_ref is built
    for key in ("metric", "unit", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config", "gpu_launches"):
        assert key in d
