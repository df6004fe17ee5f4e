#!/bin/bash
# round 2, call 35: does a one-CTA-per-SM GEMM (PIGP_GEMM_BN=128) let the chain's whole-SM kernels in sooner?
mkdir -p gpurun_out
PIGP_GEMM_BN=128 timeout 600 python tools/sweep.py --sizes 2640,5018,10570,20000 --reps 5 > gpurun_out/r02_c35_sweep_bn128.jsonl 2> gpurun_out/r02_c35_sweep.err
timeout 600 python tools/sweep.py --sizes 2640,5018,10570,20000 --reps 5 --no-library > gpurun_out/r02_c35_sweep_bn64.jsonl 2>> gpurun_out/r02_c35_sweep.err
