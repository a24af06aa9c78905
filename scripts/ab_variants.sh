#!/bin/bash
# GPU box (under gpurun): A/B of the default library against the builds of scripts/build_variants.sh on the headline and
# the HBM-stress workloads, then the parity tests on each variant.  Results -> gpurun_out/ab_variants.log as they come.
#   gpurun --timeout 600 -- 'bash scripts/ab_variants.sh'
# (Warm the box first: on a cold one the first `import torch` alone can take more than a minute.)
mkdir -p gpurun_out
LOG=gpurun_out/ab_variants.log
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
for rep in 1 2; do
  run "" planar_sweep_sdf512 200
  for v in stage4 nbhd uni ray all; do run $PWD/ab_libs/$v.so planar_sweep_sdf512 200; done
done
run "" box_room_sdf2048 60
for v in stage4 all; do run $PWD/ab_libs/$v.so box_room_sdf2048 60; done
run "" box_room_ofusion1024 60
run $PWD/ab_libs/ray.so box_room_ofusion1024 60
# the render-target extension: e2e vs e2e_render_target, with 8x4 and with 32x1 pixel tiles
for v in "" $PWD/ab_libs/t32.so; do
  SE_B200_BENCH_RENDER_TARGET=1 SE_B200_LIB=$v timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line); print('render target lib=[$v] value', d['value'], 'e2e', d['e2e'], 'e2e_render_target', d['e2e_render_target'])
" >> $LOG 2>&1
done
(echo "== extension tests"; timeout 300 python -m pytest tests/test_zz_extensions.py -x -q -m gpu 2>&1 | tail -3) >> $LOG 2>&1
for v in stage4 ray all; do
  (echo "== parity on $v"; SE_B200_LIB=$PWD/ab_libs/$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sdf_512_full or sdf_2048 or ieee_division or point_queries or ofusion_plane" 2>&1 | tail -3) >> $LOG 2>&1
done
cat $LOG
