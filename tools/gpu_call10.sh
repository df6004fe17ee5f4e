set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -15) > gpurun_out/r02_c10_tests.log
(PIGP_PROF_DUMP=gpurun_out/r02_c10_prof.csv timeout 300 python tools/one_step.py 20000 2>&1 | tail -3) > gpurun_out/r02_c10_onestep.log
(timeout 300 python tools/sweep.py --sizes 498,1180,2640,5018 --no-library 2>&1 | tail -6) > gpurun_out/r02_c10_sweep.jsonl
