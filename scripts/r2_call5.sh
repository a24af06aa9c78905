#!/bin/bash
# Round 2, GPU call 5: in-kernel streamed list (default) vs the list in a kernel of its own (SE_B200_LIST_KERNEL=1)
mkdir -p gpurun_out
LOG=gpurun_out/r2_call5.log
: > $LOG
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
run() {  # env workload steps
  env $1 timeout 300 python bench.py --workload $2 --steps $3 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        print('$2 [$1] value', d['value'], 'ms', d['ms_per_step'], 'median', d.get('ms_per_step_median'), 'e2e', d['e2e']['value'], {k: v['ms'] for k, v in d['roofline']['kernels'].items()}, 'launches/step', d['gpu_launches'] / d['steps'], d['clocks'].get('sm_mhz'))
    elif line: print(line[:300])
" >> $LOG 2>&1
}
for rep in 1 2; do
run SE_B200_LIST_KERNEL=0 planar_sweep_sdf512 200
run SE_B200_LIST_KERNEL=1 planar_sweep_sdf512 200
done
run SE_B200_LIST_KERNEL=0 box_room_sdf2048 40
run SE_B200_LIST_KERNEL=1 box_room_sdf2048 40
run SE_B200_LIST_KERNEL=0 box_room_ofusion1024 40
run SE_B200_LIST_KERNEL=1 box_room_ofusion1024 40
(echo "== gpu tests, list kernel mode"; SE_B200_LIST_KERNEL=1 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5) >> $LOG 2>&1
(echo "== gpu tests, streamed mode"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5) >> $LOG 2>&1
cat $LOG
