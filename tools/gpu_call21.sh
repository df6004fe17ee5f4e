#!/bin/bash
# round 2, call 21: potf2 store / task tweaks; look-ahead with the gradient at small N; ncu of k_potf2 and k_trsm_blk
mkdir -p gpurun_out
timeout 120 python tools/potf2_bench.py > gpurun_out/r02_c21_potf2.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_reference_pin.py -m gpu -x -q > gpurun_out/r02_c21_tests.log 2>&1
timeout 600 python tools/sweep.py --sizes 498,1180,2640,5018 --reps 5 --no-library > gpurun_out/r02_c21_sweep.jsonl 2> gpurun_out/r02_c21_sweep.err
for cfg in "1180 0,2,3" "2640 0,2,4" "5018 0,2,4,8"; do set -- $cfg; timeout 300 python tools/lookahead_sweep.py $1 $2 >> gpurun_out/r02_c21_lookahead_grad.log 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_potf2 -s 20 -c 1 -o gpurun_out/r02_c21_potf2 python tools/potf2_bench.py > gpurun_out/r02_c21_ncu_potf2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_trsm_blk -s 6 -c 1 -o gpurun_out/r02_c21_trsm python tools/one_step.py 1180 > gpurun_out/r02_c21_ncu_trsm.log 2>&1
