#!/bin/bash
# round 2, call 48: the bench-workload parity test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s -k bench_workload > gpurun_out/r02_c48_bench_parity_test.log 2>&1
