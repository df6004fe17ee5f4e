#!/bin/bash
# round 2, call 27: measurement pass on the current code (1 GPU): tests, bench line, sweep with the reference generator's points,
# optimiser rates, smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c27_tests.log 2>&1
(timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/r02_c27_bench.err | tail -1) > gpurun_out/r02_c27_bench.json
(timeout 900 python tools/sweep.py --golden 2> gpurun_out/r02_c27_sweep.err | grep "^{") > gpurun_out/r02_c27_sweep.jsonl
(timeout 600 python tools/optimizer_rate.py 200 2>&1 | grep "^{" ) > gpurun_out/r02_c27_optimizer_rate.jsonl
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c27_smoke.log 2>&1
