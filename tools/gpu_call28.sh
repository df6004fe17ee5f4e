#!/bin/bash
# round 2, call 28 (NG GPUs): bench line and the drop-in classes on the sharded solver with the faster chain kernels
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
NG=${NG:-2}

(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $NG --steps 3 --warmup 3 2> gpurun_out/r02_c28_bench$NG.err | tail -1) > gpurun_out/r02_c28_bench$NG.json
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 tools/dist_check.py 20000 0,8 2>&1 | grep "dist_check" | tail -20) > gpurun_out/r02_c28_dist$NG.log
