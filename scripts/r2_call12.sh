#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call12.log
: > $LOG
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
(echo "== gpu tests x2"; for i in 1; do timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6; done) >> $LOG 2>&1
(echo "== bench"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        print('value', d['value'], 'ms', d['ms_per_step'], 'launches/step', d['gpu_launches'] / d['steps'], {k: v['ms'] for k, v in d['roofline']['kernels'].items()})
        for k in ('e2e', 'e2e_without_render_target', 'e2e_registered', 'e2e_overlapped'):
            print(k, {a: b for a, b in d[k].items() if a != 'api'})
        for n, x in d.get('extra_workloads', {}).items():
            print(n, x.get('value'), x.get('ms_per_step'), {k: (v['ms'], v['frac']) for k, v in x['roofline']['kernels'].items()} if 'roofline' in x else x)
    elif line: print(line[:300])
") >> $LOG 2>&1
cat $LOG
