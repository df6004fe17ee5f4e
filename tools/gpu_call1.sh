set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -40) > gpurun_out/r02_c1_tests.log
(timeout 300 python tools/one_step.py 20000 2>&1 | tail -5) > gpurun_out/r02_c1_onestep.log
(PIGP_PROF_DUMP=gpurun_out/r02_c1_prof.csv timeout 300 python tools/one_step.py 20000 2>&1 | tail -5) >> gpurun_out/r02_c1_onestep.log
(timeout 600 python tools/sweep.py --golden 2>&1 | tail -20) > gpurun_out/r02_c1_sweep.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blocks -c 4 -o gpurun_out/r02_c1_kblocks python tools/one_step.py 20000 > gpurun_out/r02_c1_ncu.log 2>&1
ls -la gpurun_out | tail
