#!/bin/bash
# round 2, call 16: pair-pivot potrf32 with fused inverse, 8/16-row TRSM slabs, 32x64 update tiles
mkdir -p gpurun_out
timeout 120 python tools/potf2_bench.py > gpurun_out/r02_c16_potf2.log 2>&1
timeout 300 python tools/chol_accuracy.py > gpurun_out/r02_c16_chol_accuracy.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_pin.py tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02_c16_tests.log 2>&1
timeout 600 python tools/sweep.py --sizes 498,1180,2640,5018,10570 --reps 5 > gpurun_out/r02_c16_sweep.jsonl 2> gpurun_out/r02_c16_sweep.err
for n in 1180; do
  PIGP_PROF_DUMP=gpurun_out/r02_c16_timeline_$n.csv timeout 120 python tools/one_step.py $n >> gpurun_out/r02_c16_onestep.log 2>&1
done
