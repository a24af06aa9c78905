#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_call8.log
: > $LOG
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
(echo "== gpu tests"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) >> $LOG 2>&1
(echo "== e2e steps"; timeout 300 python scripts/e2e_steps.py 300 2>&1 | tail -8) >> $LOG 2>&1
(echo "== bench"; timeout 600 python bench.py --no-extra --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        print('value', d['value'], 'ms', d['ms_per_step'], 'launches/step', d['gpu_launches'] / d['steps'])
        for k in ('e2e', 'e2e_without_render_target', 'e2e_pageable', 'e2e_registered', 'e2e_overlapped'):
            print(k, {a: b for a, b in d[k].items() if a != 'api'})
    elif line: print(line[:300])
") >> $LOG 2>&1
cat $LOG
