set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=12 2>&1 | tail -70) > gpurun_out/r02_c4_tests.log
(timeout 300 python tools/chol_accuracy.py 2>&1 | tail -12) > gpurun_out/r02_c4_cholacc.log
(PIGP_PROF_DUMP=gpurun_out/r02_c4_prof.csv timeout 300 python tools/one_step.py 20000 2>&1 | tail -5) > gpurun_out/r02_c4_onestep.log
(timeout 600 python tools/sweep.py --golden --sizes 498,1180,2640,5018 2>&1 | tail -12) > gpurun_out/r02_c4_sweep.jsonl
