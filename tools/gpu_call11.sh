set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15) > gpurun_out/r02_c11_tests.log
(PIGP_PROF_DUMP=gpurun_out/r02_c11_prof.csv timeout 300 python tools/one_step.py 20000 2>&1 | tail -3) > gpurun_out/r02_c11_onestep.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_gemm --csv --log-file gpurun_out/r02_c11_gemm_launches.csv python tools/one_step.py 20000 > gpurun_out/r02_c11_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blocks -c 2 -o gpurun_out/r02_c11_kblocks python tools/one_step.py 20000 > gpurun_out/r02_c11_ncu_kb.log 2>&1
