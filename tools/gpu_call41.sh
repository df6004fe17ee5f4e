#!/bin/bash
# round 2, call 41 (8 GPUs): one cooperative NLL + gradient evaluation at N = 40000 and N = 80000 (102 GB per rank with the replicated layout)
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
NG=${NG:-8}
for n in 40000 80000; do
  (timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 tools/dist_check.py $n 0 2>&1 | grep "dist_check\|Error\|error" | tail -6) >> gpurun_out/r02_c41_bigN_$NG.log
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> gpurun_out/r02_c41_bigN_$NG.log 2>&1
