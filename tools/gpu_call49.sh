#!/bin/bash
# round 2, call 49: smoke + a fast subset on the final binary
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c49_smoke.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02_c49_tests.log 2>&1
