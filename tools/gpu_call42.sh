#!/bin/bash
# round 2, call 42: N = 80000 on one GPU with eps = 1 (well conditioned): rules out 64-bit indexing problems behind the NaN of the
# eps = 1e-6 workload at that size; and N = 40000 on one GPU for the strong-scaling reference
mkdir -p gpurun_out
timeout 300 python tools/one_step.py 40000 > gpurun_out/r02_c42_onestep_40000.log 2>&1
timeout 600 python tools/one_step.py 80000 1.0 > gpurun_out/r02_c42_onestep_80000_eps1.log 2>&1
