#!/bin/bash
# round 2, call 15: automatic panel schedule for NLL-only, potf2 phase stamps, per-launch timelines at small N
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -k "lookahead or automatic" > gpurun_out/r02_c15_tests.log 2>&1
timeout 120 python tools/potf2_bench.py > gpurun_out/r02_c15_potf2.log 2>&1
for n in 498 1180 2640; do
  PIGP_PROF_DUMP=gpurun_out/r02_c15_timeline_$n.csv timeout 120 python tools/one_step.py $n >> gpurun_out/r02_c15_onestep.log 2>&1
done
timeout 600 python tools/sweep.py --sizes 2640,5018,10570,20000 --reps 3 --no-library > gpurun_out/r02_c15_sweep.jsonl 2> gpurun_out/r02_c15_sweep.err
