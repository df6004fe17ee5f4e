#!/bin/bash
# round 2, call 44: final validation of the committed build (tests, smoke, short bench)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c44_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c44_smoke.log 2>&1
(timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02_c44_bench.err | tail -1) > gpurun_out/r02_c44_bench.json
