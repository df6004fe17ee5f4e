"""Under torchrun: time the sharded NLL+grad evaluation (device-resident inputs), max over ranks.
    python -m torch.distributed.run --nproc-per-node P tools/dist_time.py [N] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from stopro_b200 import _lib, synthetic
from stopro_b200.dist import DistSolver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
_lib.check(_lib.lib().pigp_set_device(local))
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = synthetic.stokes2d_scaling(n, n_test=16)
gp = synthetic.make_model(cfg)
gp.set_constants(cfg["r_train"], cfg["delta_y"], cfg["eps"], only_training=True)
plan = gp._training_plan(cfg["r_train"])
ds = DistSolver(plan, rank, world)
ds.connect_ipc()
theta = torch.as_tensor(cfg["theta0"], device=dev)
y = torch.as_tensor(cfg["delta_y"], device=dev)
out = torch.zeros(1 + plan.theta_len, dtype=torch.float64, device=dev)


def step():
    ds.nll_grad(theta.data_ptr(), y.data_ptr(), cfg["eps"], out.data_ptr(), out.data_ptr() + 8, None, None)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(3):
    step()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
barrier()
ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"N={n} world={world} PIGP_SIDE_CHUNK={os.environ.get('PIGP_SIDE_CHUNK', 'default')}: {ms.item():.2f} ms per NLL+grad, nll={out[0].item():.6f}")
if world > 1:
    dist.destroy_process_group()
