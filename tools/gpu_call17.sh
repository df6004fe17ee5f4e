#!/bin/bash
# round 2, call 17: two-warp factor / inverse of the 32x32 diagonal blocks
mkdir -p gpurun_out
timeout 120 python tools/potf2_bench.py > gpurun_out/r02_c17_potf2.log 2>&1
timeout 300 python tools/chol_accuracy.py > gpurun_out/r02_c17_chol_accuracy.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_pin.py -m gpu -x -q > gpurun_out/r02_c17_tests.log 2>&1
timeout 600 python tools/sweep.py --sizes 498,1180,2640,5018 --reps 5 > gpurun_out/r02_c17_sweep.jsonl 2> gpurun_out/r02_c17_sweep.err
