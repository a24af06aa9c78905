#!/bin/bash
# Round 2, second GPU call: the new kernels (alloc v2 + fused mm2meters, in-kernel active list, dense-mask ray walk, predicated
# gather, check-free normalisation) -- GPU parity tier, A/B against the round-1 library with all its experiments, ncu with source.
mkdir -p gpurun_out
LOG=gpurun_out/r2_call4.log
: > $LOG
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
(echo "== gpu tests (new lib)"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) >> $LOG 2>&1
run() {  # lib workload steps
  SE_B200_LIB=$1 timeout 300 python bench.py --workload $2 --steps $3 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        print('$2 lib=[$1] value', d['value'], 'ms', d['ms_per_step'], 'median', d.get('ms_per_step_median'), 'e2e', d['e2e']['value'], {k: v['ms'] for k, v in d['roofline']['kernels'].items()}, 'launches/step', d['gpu_launches'] / d['steps'], d['clocks'].get('sm_mhz'))
    elif line: print(line[:300])
" >> $LOG 2>&1
}
run "" planar_sweep_sdf512 150
run $PWD/ab_libs/all.so planar_sweep_sdf512 150
run "" planar_sweep_sdf512 150
run "" box_room_sdf2048 40
run $PWD/ab_libs/all.so box_room_sdf2048 40
run "" box_room_ofusion1024 40
run $PWD/ab_libs/all.so box_room_ofusion1024 40
SE_B200_BENCH_RENDER_TARGET=1 timeout 300 python bench.py --steps 150 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line); print('render target value', d['value'], 'e2e', d['e2e']['value'], 'e2e_render_target', d['e2e_render_target'], 'overlapped', d['e2e_overlapped'])
" >> $LOG 2>&1
WL=planar_sweep_sdf512
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_raycast|k_alloc_sdf|k_integrate_sdf|k_render_shade" \
    -s 24 -c 4 -o gpurun_out/r2d_full_${WL} python scripts/profile_frames.py $WL 9 > gpurun_out/r2d_full_${WL}.log 2>&1
cat $LOG
