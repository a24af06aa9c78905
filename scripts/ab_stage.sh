#!/bin/bash
# A/B of the integrate pipeline's stage size (in-tree build = whole-block stages, ab_libs/half.so = the opt-in
# csrc/se_integrate_staged.cuh built with -DSE_INT_STAGE_SLICES=4) on the headline and the HBM-stress workloads, plus the
# parity tests on the variant.  Results -> gpurun_out/ab_stage.log as they come.
#   nvcc <flags of __graft_entry__.NVCC_FLAGS> -DSE_INT_STAGE_SLICES=4 -shared -o ab_libs/half.so supereight_b200/csrc/se_b200.cu
# Round 1: the one call that ran this landed on a cold box (56 s to acquire; the first `import torch` alone outlasted the
# first run's timeout) and produced nothing before the budget ended -- warm the box first, as below.
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
mkdir -p gpurun_out
LOG=gpurun_out/ab_stage.log
run() {  # lib workload steps
  SE_B200_LIB=$1 timeout 120 python bench.py --workload $2 --steps $3 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        print('$2 lib=[$1] value', d['value'], 'ms', d['ms_per_step'], 'median', d.get('ms_per_step_median'), 'e2e', d['e2e']['value'], {k: v['ms'] for k, v in d['roofline']['kernels'].items()}, d['clocks'].get('sm_mhz'))
    elif line: print(line[:300])
" >> $LOG 2>&1
}
H=$PWD/ab_libs/half.so
run "" planar_sweep_sdf512 200
run $H planar_sweep_sdf512 200
run $H planar_sweep_sdf512 200
run "" planar_sweep_sdf512 200
run "" box_room_sdf2048 60
run $H box_room_sdf2048 60
(SE_B200_LIB=$H timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sdf_512_full or sdf_2048 or ieee_division" 2>&1 | tail -3) >> $LOG 2>&1
cat $LOG
