#!/bin/bash
# round 2, call 18: block-substitution TRSM; potf2 load/store/inverse trims; diagonal-block update over four warps
mkdir -p gpurun_out
timeout 120 python tools/potf2_bench.py > gpurun_out/r02_c18_potf2.log 2>&1
timeout 300 python tools/chol_accuracy.py > gpurun_out/r02_c18_chol_accuracy.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_pin.py tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02_c18_tests.log 2>&1
timeout 600 python tools/sweep.py --sizes 498,1180,2640,5018,10570 --reps 5 > gpurun_out/r02_c18_sweep.jsonl 2> gpurun_out/r02_c18_sweep.err
PIGP_TRSM=refine timeout 600 python tools/sweep.py --sizes 1180,5018 --reps 5 --no-library > gpurun_out/r02_c18_sweep_refine.jsonl 2>> gpurun_out/r02_c18_sweep.err
PIGP_PROF_DUMP=gpurun_out/r02_c18_timeline_1180.csv timeout 120 python tools/one_step.py 1180 >> gpurun_out/r02_c18_onestep.log 2>&1
