"""Under torchrun: per-launch timeline (class, duration, start, stream) of one sharded NLL+grad evaluation per rank,
written to gpurun_out/timeline_w<world>_r<rank>.csv (two-stream mode, as timed by bench.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from stopro_b200 import _lib, synthetic
from stopro_b200.dist import DistSolver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
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
P = plan.theta_len
theta = torch.as_tensor(cfg["theta0"], device=dev)
y = torch.as_tensor(cfg["delta_y"], device=dev)
out = torch.zeros(1 + P, dtype=torch.float64, device=dev)


def step():
    ds.nll_grad(theta.data_ptr(), y.data_ptr(), cfg["eps"], out.data_ptr(), out.data_ptr() + 8, None, None)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(2):
    step()
barrier()
os.makedirs("gpurun_out", exist_ok=True)
os.environ["PIGP_PROF_DUMP"] = f"gpurun_out/timeline_w{world}_r{rank}.csv"
_lib.profile_start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
prof = _lib.profile_stop()
barrier()
print(f"rank {rank}: {e0.elapsed_time(e1):.2f} ms (with per-launch events)", {k: (round(v['ms'], 2), v['launches']) for k, v in prof.items()})
if world > 1:
    dist.destroy_process_group()
