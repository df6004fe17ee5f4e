#!/bin/bash
# round 2, call 34: K^-1 accumulated from finished column ranges of Y underneath the chains (single GPU, 8 <= tiles <= 100)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c34_tests.log 2>&1
timeout 600 python tools/sweep.py --sizes 1180,2640,5018,8000,10570,12700 --reps 5 --no-library > gpurun_out/r02_c34_sweep.jsonl 2> gpurun_out/r02_c34_sweep.err
PIGP_EARLY_KINV=0 timeout 600 python tools/sweep.py --sizes 1180,2640,5018,8000,10570,12700 --reps 5 --no-library > gpurun_out/r02_c34_sweep_off.jsonl 2>> gpurun_out/r02_c34_sweep.err
