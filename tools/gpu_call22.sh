#!/bin/bash
# round 2, call 22: inverse tiles completed off the critical path (k_tile_inv), potf2 leaves the diagonal blocks only
mkdir -p gpurun_out
timeout 120 python tools/potf2_bench.py > gpurun_out/r02_c22_potf2.log 2>&1
timeout 300 python tools/chol_accuracy.py > gpurun_out/r02_c22_chol_accuracy.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c22_tests.log 2>&1
timeout 600 python tools/sweep.py --sizes 498,1180,2640,5018,10570,20000 --reps 5 > gpurun_out/r02_c22_sweep.jsonl 2> gpurun_out/r02_c22_sweep.err
PIGP_PROF_DUMP=gpurun_out/r02_c22_timeline_1180.csv timeout 120 python tools/one_step.py 1180 >> gpurun_out/r02_c22_onestep.log 2>&1
