#!/bin/bash
# The driver's own sequence, as a rehearsal: GPU parity tier, smoke, the default bench line (with extra_workloads and the CPU
# baseline), the reference arm.  Wall times recorded.
mkdir -p gpurun_out
R=${1:-r2}
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
(time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4) > gpurun_out/${R}_pytest_gpu.log 2>&1
(time python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/${R}_smoke.log 2>&1
(time python bench.py > gpurun_out/bench_${R}_sdf512.json 2> gpurun_out/bench_${R}_sdf512.err) 2> gpurun_out/bench_${R}_time.log
(time python bench.py --impl reference > gpurun_out/bench_${R}_reference_arm.json 2> gpurun_out/bench_${R}_reference_arm.err) 2>> gpurun_out/bench_${R}_time.log
cat gpurun_out/${R}_pytest_gpu.log gpurun_out/${R}_smoke.log gpurun_out/bench_${R}_time.log
tail -c 1500 gpurun_out/bench_${R}_sdf512.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2_sdf512.json"))
print("value", d["value"], "ms", d["ms_per_step"], "launches/step", d["gpu_launches"] / d["steps"], "plain", d["value_without_render_target"])
for k in ("e2e", "e2e_without_render_target", "e2e_pageable", "e2e_registered", "e2e_overlapped"):
    print(k, {a: b for a, b in d[k].items() if a != "api"})
print("images_agree", d["images_agree"], "clocks", d["clocks"])
print("kernels", json.dumps(d["roofline"]["kernels"]))
for n, x in d.get("extra_workloads", {}).items():
    print(n, json.dumps({a: b for a, b in x.items() if a != "config"})[:1500])
print("cpu", d.get("cpu_baseline"))
r = json.load(open("gpurun_out/bench_r2_reference_arm.json"))
print("reference arm", r["value"], r["steps"], r["cpu_baseline"]["sample"])
print("same config", r["config"] == d["config"])
PY
