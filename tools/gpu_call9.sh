set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 20000 0,8,16,32 2>&1 | grep -v "^W\|warn" | tail -20) > gpurun_out/r02_c9_dist2.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/r02_c9_bench2.err | tail -1) > gpurun_out/r02_c9_bench2.json
