#!/bin/bash
# round 2, call 36: full GPU suite incl. the new factorisation tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c36_tests.log 2>&1
