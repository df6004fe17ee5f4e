"""Multi-GPU product path under torchrun (one rank per GPU):
  1. GPmodel(distributed=True) -- trainingFunction_all / d_trainingFunction_all / predict_many routed to the sharded solver --
     against the same model on one GPU (rank 0 evaluates both);
  2. NLL+gradient time for several look-ahead panel widths (pigp_set_lookahead).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py [N] [widths]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from stopro_b200 import _lib, synthetic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
_lib.check(_lib.lib().pigp_set_device(local))
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
widths = [int(w) for w in (sys.argv[2] if len(sys.argv) > 2 else "0,8,16,32").split(",")]

# ---- 1. drop-in classes on the sharded solver
cfg = dict(synthetic.stokes2d_scaling(3000, n_test=700), eps=1.0)
args = (cfg["r_train"], cfg["delta_y"], cfg["eps"])
pargs = (cfg["r_test"], cfg["mu_test"]) + args
gp = synthetic.make_model(cfg).enable_distributed()
gp.set_constants(*pargs)
th = cfg["theta0"] + 0.05
nll = gp.trainingFunction_all(th, *args)
grad = gp.d_trainingFunction_all(th, *args)
mus, vars_ = gp.predict_many([th, th - 0.1], *pargs)
mu_full, cov_full = gp.predictingFunction_all(th, *pargs)       # full covariance: every rank evaluates everything
if rank == 0:
    one = synthetic.make_model(cfg)
    one.set_constants(*pargs)
    nll1, grad1 = one.value_and_grad(th, *args)
    mus1, vars1 = one.predict_many([th, th - 0.1], *pargs)
    e = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(np.asarray(b))), 1e-300))
    print(f"[dist_check] world={world}: nll rel diff {abs(nll - nll1) / abs(nll1):.2e}, grad {e(grad, grad1):.2e}, "
          f"mu {max(e(a, b) for A, B in zip(mus, mus1) for a, b in zip(A, B)):.2e}, "
          f"var {max(e(a, b) for A, B in zip(vars_, vars1) for a, b in zip(A, B)):.2e}, "
          f"full-cov diag vs sharded var {max(e(np.diag(c), v) for c, v in zip(cov_full, vars_[0])):.2e}", flush=True)
    one.close()
gp.close()
dist.barrier()

# ---- 2. look-ahead widths at N points
from stopro_b200.plan import Solver

cfg = synthetic.stokes2d_scaling(n, n_test=16)
gp = synthetic.make_model(cfg)
gp.set_constants(cfg["r_train"], cfg["delta_y"], cfg["eps"], only_training=True)
plan = gp._training_plan(cfg["r_train"])
solver = Solver(plan, rank, world)
solver.connect_ipc()
P = plan.theta_len
theta = torch.as_tensor(cfg["theta0"], device=dev)
y = torch.as_tensor(cfg["delta_y"], device=dev)
out = torch.zeros(1 + P, dtype=torch.float64, device=dev)
for w in widths:
    _lib.check(_lib.lib().pigp_set_lookahead(w))
    best = 1e30
    for it in range(4):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        solver.nll_grad(theta.data_ptr(), y.data_ptr(), cfg["eps"], out.data_ptr(), out.data_ptr() + 8, None, None)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if it:
            best = min(best, float(ms.item()))
    if rank == 0:
        print(f"[dist_check] world={world} N={n} lookahead={w:3d}: {best:9.3f} ms  nll={out[0].item():.6f}", flush=True)
solver.close()
gp.close()
dist.destroy_process_group()
