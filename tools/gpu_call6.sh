set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -30) > gpurun_out/r02_c6_tests.log
(PIGP_PROF_DUMP=gpurun_out/r02_c6_prof.csv timeout 300 python tools/one_step.py 20000 2>&1 | tail -5) > gpurun_out/r02_c6_onestep.log
(timeout 600 python tools/lookahead_sweep.py 20000 0,8,16,32 2>&1 | tail -8) > gpurun_out/r02_c6_lookahead.log
(timeout 300 python tools/lookahead_sweep.py 5018 0,2,4,8 2>&1 | tail -8) >> gpurun_out/r02_c6_lookahead.log
(timeout 300 python tools/lookahead_sweep.py 2640 0,2,4 2>&1 | tail -8) >> gpurun_out/r02_c6_lookahead.log
(timeout 900 python bench.py --steps 3 --warmup 3 2> gpurun_out/r02_c6_bench.err | tail -2) > gpurun_out/r02_c6_bench.json
