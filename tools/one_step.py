"""One warm-up and one timed NLL+gradient evaluation at N points (default 20000): the command profiled by ncu
(profiles/) and the source of the per-launch event dump (PIGP_PROF_DUMP=<csv>)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stopro_b200 import _lib, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
cfg = synthetic.stokes2d_scaling(n, n_test=16)
if len(sys.argv) > 2:  # optional jitter override (eps = 1 makes the workload well conditioned at any size)
    cfg = dict(cfg, eps=float(sys.argv[2]))
gp = synthetic.make_model(cfg)
gp.set_constants(cfg["r_train"], cfg["delta_y"], cfg["eps"], only_training=True)
solver = gp._solver_for(cfg["r_train"])
P = solver.plan.theta_len
dev = torch.device("cuda:0")
theta = torch.as_tensor(cfg["theta0"], device=dev)
y = torch.as_tensor(cfg["delta_y"], device=dev)
out = torch.zeros(1 + P, dtype=torch.float64, device=dev)
if os.environ.get("PIGP_SERIAL"):  # every kernel on one stream: per-launch event times do not overlap
    _lib.check(_lib.lib().pigp_set_side_stream(0))
for it in range(2):
    if it == 1 and os.environ.get("PIGP_PROF_DUMP"):
        _lib.profile_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    solver.nll_grad(theta.data_ptr(), y.data_ptr(), cfg["eps"], out.data_ptr(), out.data_ptr() + 8, None, None)
    e1.record()
    torch.cuda.synchronize()
    if it == 1 and os.environ.get("PIGP_PROF_DUMP"):
        print({k: (round(v["ms"], 3), v["launches"]) for k, v in _lib.profile_stop().items()})
    print(f"eval {it}: {e0.elapsed_time(e1):.2f} ms  nll={out[0].item():.6f}")
