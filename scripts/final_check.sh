#!/bin/bash
# Last GPU call of a round, most important first: the GPU parity tests on the final library, the bench lines of the three
# workloads (no CPU arm: that one is timed by the driver's own run), then the ncu launch list and one full capture of the
# headline workload.  Everything lands in gpurun_out/ as it is produced.
R=${1:-r1c}
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1      # a cold box needs up to a minute for the first import
(timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -5) > gpurun_out/${R}_pytest_gpu.log 2>&1
python bench.py --no-cpu-baseline > gpurun_out/bench_${R}_sdf512.json 2> gpurun_out/bench_${R}_sdf512.err
python bench.py --workload box_room_ofusion1024 --no-cpu-baseline > gpurun_out/bench_${R}_ofusion1024.json 2> gpurun_out/bench_${R}_ofusion1024.err
python bench.py --workload box_room_sdf2048 --no-cpu-baseline > gpurun_out/bench_${R}_sdf2048.json 2> gpurun_out/bench_${R}_sdf2048.err
WL=planar_sweep_sdf512
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_${WL}.csv \
    python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_launches_${WL}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_raycast|k_alloc_sdf|k_integrate_sdf|k_active_list|k_render_shade|k_mm2meters" \
    -s 36 -c 6 -o gpurun_out/${R}_full_${WL} python scripts/profile_frames.py $WL 9 > gpurun_out/${R}_full_${WL}.log 2>&1
cat gpurun_out/${R}_pytest_gpu.log; tail -c 600 gpurun_out/bench_${R}_*.json
