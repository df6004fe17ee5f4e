set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -30) > gpurun_out/r02_c7_tests.log
(PIGP_PROF_DUMP=gpurun_out/r02_c7_prof.csv timeout 300 python tools/one_step.py 20000 2>&1 | tail -5) > gpurun_out/r02_c7_onestep.log
