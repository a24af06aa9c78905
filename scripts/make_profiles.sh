#!/bin/bash
# Runs on the GPU box (under gpurun): ncu launch lists of the bench command and full captures (with source) of one frame's
# kernels at the bench's operating point -- the middle frame of the frames bench.py times: frame 160 of the 300-frame sweep
# for the headline workload, frame 20 for the two 30-step extra_workloads legs.
# Outputs go to gpurun_out/; scripts/summarise_profiles.py <round> (run in the build container) turns them into profiles/
# (launch lists, key-metric tables, opcode / stall summaries and profiles/traffic.json, which bench.py reads).
set -x
R=${1:-r2b}
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
capture() {  # workload kernels-per-frame frame regex
  local WL=$1 PER=$2 FRAME=$3 REGEX=$4
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_${WL}.csv \
      python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${R}_launches_${WL}.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $((PER * FRAME)) -c $PER \
      -o gpurun_out/${R}_full_${WL} python scripts/profile_frames.py $WL $((FRAME + 2)) > gpurun_out/${R}_full_${WL}.log 2>&1
}
capture planar_sweep_sdf512 5 160 "k_alloc_sdf|k_filter_blocks|k_integrate_sdf|k_raycast|k_render_shade"
capture box_room_sdf2048 5 20 "k_alloc_sdf|k_filter_blocks|k_integrate_sdf|k_raycast|k_render_shade"
capture box_room_ofusion1024 5 20 "k_alloc_ofusion|k_alloc_first|k_integrate_ofusion|k_raycast|k_render_shade"
ls -la gpurun_out
