set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60) > gpurun_out/r02_c2_tests.log
(PIGP_PROF_DUMP=gpurun_out/r02_c2_prof.csv timeout 300 python tools/one_step.py 20000 2>&1 | tail -5) > gpurun_out/r02_c2_onestep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blocks -c 2 -o gpurun_out/r02_c2_kblocks python tools/one_step.py 20000 > gpurun_out/r02_c2_ncu.log 2>&1
