#!/bin/bash
# Runs on the GPU box (under gpurun): refreshes the evidence for the headline workload only (launch list, one full
# capture) and the bench lines of every workload + the reference arm.  scripts/summarise_profiles.py turns the ncu files
# into profiles/ in the build container.
set -x
R=${1:-r1}
mkdir -p gpurun_out
WL=planar_sweep_sdf512
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_${WL}.csv \
    python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_launches_${WL}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_raycast|k_alloc_sdf|k_integrate_sdf|k_active_list|k_render_shade|k_mm2meters" \
    -s 36 -c 6 -o gpurun_out/${R}_full_${WL} python scripts/profile_frames.py $WL 9 > gpurun_out/${R}_full_${WL}.log 2>&1
python bench.py > gpurun_out/bench_${R}_sdf512.json 2> gpurun_out/bench_${R}_sdf512.err
python bench.py --impl reference > gpurun_out/bench_${R}_reference_arm.json 2> gpurun_out/bench_${R}_reference_arm.err
python bench.py --workload box_room_sdf2048 --no-cpu-baseline > gpurun_out/bench_${R}_sdf2048.json 2> gpurun_out/bench_${R}_sdf2048.err
python bench.py --workload box_room_ofusion1024 --no-cpu-baseline > gpurun_out/bench_${R}_ofusion1024.json 2> gpurun_out/bench_${R}_ofusion1024.err
tail -c 400 gpurun_out/bench_${R}_*.json
