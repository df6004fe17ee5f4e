#!/bin/bash
# round 2, call 37: per-launch times and flops of one evaluation at N = 20000 (serial pass) -- where the GEMM class loses its 10 %
mkdir -p gpurun_out
PIGP_SERIAL=1 PIGP_PROF_DUMP=gpurun_out/r02_c37_serial_20000.csv timeout 300 python tools/one_step.py 20000 > gpurun_out/r02_c37_onestep.log 2>&1
