#!/bin/bash
# one short GPU call: A/B of library builds (ab_libs/*.so against the in-tree one), then the GPU parity tests.
# Results go to gpurun_out/ as they come, so a call cut off by the budget still leaves what it finished.
mkdir -p gpurun_out
STEPS=${STEPS:-150}
run() {
  SE_B200_LIB=$1 timeout 90 python bench.py --steps $STEPS --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        print('lib=[$1] value', d['value'], 'ms', d['ms_per_step'], 'median', d.get('ms_per_step_median'), 'e2e', d['e2e']['value'], {k: v['ms'] for k, v in d['roofline']['kernels'].items()}, d['clocks'])
    elif line: print(line[:300])
" >> gpurun_out/ab.log 2>&1
}
run ""
for l in $ROOT_LIBS; do run $PWD/ab_libs/$l.so; done
(timeout ${PYTEST_TIMEOUT:-600} python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/pytest_gpu.log 2>&1
for f in 0 1; do
  SE_B200_OFUSION_FAST=$f timeout 90 python bench.py --workload box_room_ofusion1024 --steps 60 --warmup 5 --no-cpu-baseline 2>&1 | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('ofusion fast=$f value', d['value'], 'ms', d['ms_per_step'], {k: v['ms'] for k, v in d['roofline']['kernels'].items()})" >> gpurun_out/ab.log 2>&1
done
cat gpurun_out/ab.log gpurun_out/pytest_gpu.log
