set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=25 2>&1 | tail -80) > gpurun_out/r02_c3_tests.log
(PIGP_PROF_DUMP=gpurun_out/r02_c3_prof.csv timeout 300 python tools/one_step.py 20000 2>&1 | tail -5) > gpurun_out/r02_c3_onestep.log
(timeout 300 python tools/chol_accuracy.py 2>&1 | tail -12) > gpurun_out/r02_c3_cholacc.log
