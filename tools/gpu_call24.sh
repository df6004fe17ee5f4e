#!/bin/bash
# round 2, call 24: shuffled pivot chain, cp.async loads restored, stand-alone potrf with the panel schedule
mkdir -p gpurun_out
timeout 120 python tools/potf2_bench.py > gpurun_out/r02_c24_potf2.log 2>&1
timeout 300 python tools/chol_accuracy.py > gpurun_out/r02_c24_chol_accuracy.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c24_tests.log 2>&1
timeout 600 python tools/sweep.py --sizes 498,1180,2640,5018,10570,20000 --reps 5 > gpurun_out/r02_c24_sweep.jsonl 2> gpurun_out/r02_c24_sweep.err
