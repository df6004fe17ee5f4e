set -x
cd $GRAFT_REPO_ROOT
(timeout 600 python tools/optimizer_rate.py 200 2>&1 | grep "^{" ) > gpurun_out/r02_c13_optimizer_rate.jsonl
(timeout 900 python tools/sweep.py --golden 2>&1 | grep "^{") > gpurun_out/r02_c13_sweep.jsonl
(timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/r02_c13_bench.err | tail -1) > gpurun_out/r02_c13_bench.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c13_smoke.log 2>&1
