#!/bin/bash
# round 2, call 47: the new two-rank eps = 1e-6 fixture test, then the whole GPU suite once more
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reference_pin.py -m gpu -x -q -s -k two_ranks > gpurun_out/r02_c47_two_ranks.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c47_tests.log 2>&1
