set -x
cd $GRAFT_REPO_ROOT
for v in g00 g01 g10; do
  (PIGP_LIB=stopro_b200/libpigp_$v.so PIGP_PROF_DUMP=gpurun_out/r02_c8_prof_$v.csv timeout 300 python tools/one_step.py 20000 2>&1 | tail -2) > gpurun_out/r02_c8_onestep_$v.log
done
(PIGP_PROF_DUMP=gpurun_out/r02_c8_prof_g11.csv timeout 300 python tools/one_step.py 20000 2>&1 | tail -2) > gpurun_out/r02_c8_onestep_g11.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blocks -c 2 -o gpurun_out/r02_c8_kblocks python tools/one_step.py 20000 > gpurun_out/r02_c8_ncu.log 2>&1
