set -x
cd $GRAFT_REPO_ROOT
NG=${NG:-8}
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 tools/dist_check.py 20000 0,4,8,16 2>&1 | grep "dist_check" | tail -20) > gpurun_out/r02_c12_dist$NG.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $NG --steps 3 --warmup 3 2> gpurun_out/r02_c12_bench$NG.err | tail -1) > gpurun_out/r02_c12_bench$NG.json
