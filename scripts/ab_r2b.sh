#!/bin/bash
# A/B on the GPU box: the in-tree library against supereight_b200/variants/libse_b200_<name>.so builds (SE_B200_LIB), same box, same call.
# usage: scripts/ab_r2b.sh <log-name> <workload:steps> ...   (variants: every .so under supereight_b200/variants)
mkdir -p gpurun_out
LOG=gpurun_out/$1.log; shift
: > $LOG
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
summ() { python -c "
import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        k = d['roofline']['kernels']
        print('value', d['value'], 'ms', d['ms_per_step'], 'median', d.get('ms_per_step_median'), 'e2e', d['e2e']['value'], 'stages us', {s: round(1000 * v['ms'], 2) for s, v in k.items()}, 'frac', {s: v['frac'] for s, v in k.items()})
    elif line: print(line[:300])
"; }
for spec in "$@"; do
  WL=${spec%%:*}; STEPS=${spec##*:}
  for rep in 1 2; do
    for lib in in-tree $(ls supereight_b200/variants/*.so 2>/dev/null); do
      echo "== $WL steps $STEPS lib $lib rep $rep" >> $LOG
      if [ "$lib" = in-tree ]; then unset SE_B200_LIB; else export SE_B200_LIB=$PWD/$lib; fi
      timeout 600 python bench.py --workload $WL --steps $STEPS --warmup ${WARMUP:-5} --no-extra --no-cpu-baseline 2>&1 | summ >> $LOG 2>&1
    done
  done
done
unset SE_B200_LIB
cat $LOG
