#!/bin/bash
# round 2, call 26: potf2 with an uncontended critical path (phase A alone on the tensor pipe, background warps, early stores)
mkdir -p gpurun_out
timeout 120 python tools/potf2_bench.py > gpurun_out/r02_c26_potf2.log 2>&1
timeout 300 python tools/chol_accuracy.py > gpurun_out/r02_c26_chol_accuracy.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_pin.py tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02_c26_tests.log 2>&1
timeout 600 python tools/sweep.py --sizes 498,1180,2640,5018,10570,20000 --reps 5 > gpurun_out/r02_c26_sweep.jsonl 2> gpurun_out/r02_c26_sweep.err
