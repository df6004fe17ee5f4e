set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python bench.py --steps 3 --warmup 3 2> gpurun_out/r02_c5_bench.err | tail -2) > gpurun_out/r02_c5_bench.json
(timeout 600 python -m pytest tests/test_gpu_next_rows.py -m gpu -q --timeout 600 2>&1 | tail -30) > gpurun_out/r02_c5_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blocks -c 2 -o gpurun_out/r02_c5_kblocks python tools/one_step.py 20000 > gpurun_out/r02_c5_ncu.log 2>&1
(timeout 300 python tools/potf2_bench.py 2>&1 | tail -12) > gpurun_out/r02_c5_potf2.log
