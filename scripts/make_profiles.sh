#!/bin/bash
# Runs on the GPU box (under gpurun): ncu launch lists of the bench command and full captures of the hot kernels.
# Outputs go to gpurun_out/; scripts/summarise_profiles.py (run in the build container) turns them into profiles/.
set -x
R=${1:-r1}
mkdir -p gpurun_out
for WL in planar_sweep_sdf512 box_room_sdf2048; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_${WL}.csv \
      python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_launches_${WL}.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"k_raycast|k_alloc_sdf|k_integrate_sdf|k_active_list|k_render_shade|k_mm2meters" \
      -s 36 -c 6 -o gpurun_out/${R}_full_${WL} python scripts/profile_frames.py $WL 9 > gpurun_out/${R}_full_${WL}.log 2>&1
done
WL=box_room_ofusion1024
ncu --set full --clock-control none --import-source on -k regex:"k_raycast|k_alloc_ofusion|k_integrate_ofusion|k_active_list|k_alloc_first" \
    -s 35 -c 5 -o gpurun_out/${R}_full_${WL} python scripts/profile_frames.py $WL 9 > gpurun_out/${R}_full_${WL}.log 2>&1
ls -la gpurun_out
