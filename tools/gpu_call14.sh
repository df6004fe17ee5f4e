set -x
cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -8) > gpurun_out/r02_c14_tests.log
(timeout 300 python tools/lookahead_sweep.py 20000 0,4,8,16,32 nll 2>&1 | tail -6) > gpurun_out/r02_c14_lookahead_nll.log
(timeout 300 python tools/lookahead_sweep.py 10570 0,4,8,16 nll 2>&1 | tail -6) >> gpurun_out/r02_c14_lookahead_nll.log
(timeout 300 python tools/lookahead_sweep.py 5018 0,2,4,8 nll 2>&1 | tail -6) >> gpurun_out/r02_c14_lookahead_nll.log
