#!/bin/bash
# Last GPU call of a round, most important first: the GPU parity tier on the final library, the driver's own sequence (smoke, the
# default bench line, the reference arm) with wall times, then the ncu evidence (scripts/make_profiles.sh).
R=${1:-r2b}
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
(time timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6) > gpurun_out/${R}_pytest_gpu.log 2>&1
(time python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/${R}_smoke.log 2>&1
(time python bench.py > gpurun_out/bench_${R}_sdf512.json 2> gpurun_out/bench_${R}_sdf512.err) 2> gpurun_out/bench_${R}_time.log
(time python bench.py --impl reference > gpurun_out/bench_${R}_reference_arm.json 2> gpurun_out/bench_${R}_reference_arm.err) 2>> gpurun_out/bench_${R}_time.log
# ... and with the arguments the driver passes (the sweep's first frames: longer lists, a growing map)
(time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${R}_sdf512_driver_args.json 2> /dev/null) 2>> gpurun_out/bench_${R}_time.log
(time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${R}_reference_arm_driver_args.json 2> /dev/null) 2>> gpurun_out/bench_${R}_time.log
cat gpurun_out/${R}_pytest_gpu.log gpurun_out/${R}_smoke.log gpurun_out/bench_${R}_time.log
tail -c 600 gpurun_out/bench_${R}_sdf512.err
bash scripts/make_profiles.sh $R > gpurun_out/${R}_make_profiles.log 2>&1
tail -3 gpurun_out/${R}_make_profiles.log
