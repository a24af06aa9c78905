#!/bin/bash
# Round 2, first GPU call: A/B of the round-1 compile-time experiments (scripts/build_variants.sh) on the three workloads,
# the render-target extension, the GPU parity tier, and one full ncu capture (with source) of the headline kernels.
mkdir -p gpurun_out
LOG=gpurun_out/r2_call1.log
: > $LOG
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
run() {  # lib workload steps
  SE_B200_LIB=$1 timeout 300 python bench.py --workload $2 --steps $3 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        print('$2 lib=[$1] value', d['value'], 'ms', d['ms_per_step'], 'median', d.get('ms_per_step_median'), 'e2e', d['e2e']['value'], {k: v['ms'] for k, v in d['roofline']['kernels'].items()}, d['clocks'].get('sm_mhz'))
    elif line: print(line[:300])
" >> $LOG 2>&1
}
run "" planar_sweep_sdf512 150
for v in stage4 nbhd uni ray all t32; do run $PWD/ab_libs/$v.so planar_sweep_sdf512 150; done
run "" planar_sweep_sdf512 150
run "" box_room_sdf2048 40
for v in stage4 all; do run $PWD/ab_libs/$v.so box_room_sdf2048 40; done
run "" box_room_ofusion1024 40
run $PWD/ab_libs/ray.so box_room_ofusion1024 40
for v in "" $PWD/ab_libs/t32.so; do
  SE_B200_BENCH_RENDER_TARGET=1 SE_B200_LIB=$v timeout 300 python bench.py --steps 150 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line); print('render target lib=[$v] value', d['value'], 'e2e', d['e2e'], 'e2e_render_target', d['e2e_render_target'])
" >> $LOG 2>&1
done
(echo "== gpu tests (default lib)"; timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3) >> $LOG 2>&1
WL=planar_sweep_sdf512
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_raycast|k_alloc_sdf|k_integrate_sdf|k_active_list|k_render_shade|k_mm2meters" \
    -s 36 -c 6 -o gpurun_out/r2a_full_${WL} python scripts/profile_frames.py $WL 9 > gpurun_out/r2a_full_${WL}.log 2>&1
cat $LOG
