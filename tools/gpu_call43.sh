#!/bin/bash
# round 2, call 43 (4 GPUs): bench line at 4 GPUs; ncu --set full of the chain kernels (k_potf2, k_trsm_blk) on GPU 0
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 3 --warmup 3 2> gpurun_out/r02_c43_bench4.err | tail -1) > gpurun_out/r02_c43_bench4.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_potf2 -s 20 -c 1 -o gpurun_out/r02_c43_potf2 python tools/potf2_bench.py > gpurun_out/r02_c43_ncu_potf2.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_trsm_blk -s 6 -c 1 -o gpurun_out/r02_c43_trsm python tools/one_step.py 1180 > gpurun_out/r02_c43_ncu_trsm.log 2>&1
