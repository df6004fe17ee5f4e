#!/bin/bash
# round 2, call 32: inverse tiles on their own stream (off the L^-T chain)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c32_tests.log 2>&1
timeout 600 python tools/sweep.py --sizes 498,1180,2640,5018,10570,20000 --reps 5 --no-library > gpurun_out/r02_c32_sweep.jsonl 2> gpurun_out/r02_c32_sweep.err
PIGP_PROF_DUMP=gpurun_out/r02_c32_timeline_2640.csv timeout 120 python tools/one_step.py 2640 >> gpurun_out/r02_c32_onestep.log 2>&1
