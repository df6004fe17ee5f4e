#!/bin/bash
# round 2, call 50: the whole GPU suite on the final tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c50_tests.log 2>&1
