#!/bin/bash
# round 2, call 25: potf2 with the panel off its critical path; panel width re-swept with the faster chain
mkdir -p gpurun_out
timeout 120 python tools/potf2_bench.py > gpurun_out/r02_c25_potf2.log 2>&1
timeout 300 python tools/chol_accuracy.py > gpurun_out/r02_c25_chol_accuracy.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_pin.py tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02_c25_tests.log 2>&1
timeout 600 python tools/sweep.py --sizes 498,1180,2640,5018,10570 --reps 5 > gpurun_out/r02_c25_sweep.jsonl 2> gpurun_out/r02_c25_sweep.err
for cfg in "2640 0,2,3,4" "5018 2,4,6,8" "10570 4,6,8,12,16" "20000 8,12,16,24"; do set -- $cfg; timeout 300 python tools/lookahead_sweep.py $1 $2 nll >> gpurun_out/r02_c25_lookahead_nll.log 2>&1; done
for cfg in "10570 0,4,8,12" "20000 0,8,16"; do set -- $cfg; timeout 300 python tools/lookahead_sweep.py $1 $2 >> gpurun_out/r02_c25_lookahead_grad.log 2>&1; done
