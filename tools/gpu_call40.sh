#!/bin/bash
# round 2, call 40: final measurement pass (1 GPU): tests, bench line, sweep with the reference generator's points, optimiser
# rates, potf2 phases, factorisation accuracy, smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c40_tests.log 2>&1
(timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/r02_c40_bench.err | tail -1) > gpurun_out/r02_c40_bench.json
(timeout 900 python tools/sweep.py --golden 2> gpurun_out/r02_c40_sweep.err | grep "^{") > gpurun_out/r02_c40_sweep.jsonl
(timeout 600 python tools/optimizer_rate.py 200 2>&1 | grep "^{" ) > gpurun_out/r02_c40_optimizer_rate.jsonl
timeout 120 python tools/potf2_bench.py > gpurun_out/r02_c40_potf2.log 2>&1
timeout 300 python tools/chol_accuracy.py > gpurun_out/r02_c40_chol_accuracy.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c40_smoke.log 2>&1
(timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2> gpurun_out/r02_c40_benchref.err | tail -1) > gpurun_out/r02_c40_benchref.json
