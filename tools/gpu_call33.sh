#!/bin/bash
# round 2, call 33: 16-row assembly / gradient tiles for small problems
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c33_tests.log 2>&1
timeout 600 python tools/sweep.py --sizes 498,1180,2640,5018 --reps 5 --no-library --golden > gpurun_out/r02_c33_sweep.jsonl 2> gpurun_out/r02_c33_sweep.err
