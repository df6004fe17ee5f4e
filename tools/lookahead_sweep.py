"""NLL+gradient time at N points on one GPU for several look-ahead panel widths, and agreement of the results."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stopro_b200 import _lib, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
widths = [int(w) for w in (sys.argv[2] if len(sys.argv) > 2 else "0,4,8,16,32").split(",")]
nll_only = len(sys.argv) > 3 and sys.argv[3] == "nll"
cfg = synthetic.stokes2d_scaling(n, n_test=16)
gp = synthetic.make_model(cfg)
gp.set_constants(cfg["r_train"], cfg["delta_y"], cfg["eps"], only_training=True)
solver = gp._solver_for(cfg["r_train"])
P = solver.plan.theta_len
dev = torch.device("cuda:0")
theta = torch.as_tensor(cfg["theta0"], device=dev)
y = torch.as_tensor(cfg["delta_y"], device=dev)
out = torch.zeros(1 + P, dtype=torch.float64, device=dev)
ref = None
for w in widths:
    _lib.check(_lib.lib().pigp_set_lookahead(w))
    best = 1e30
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if nll_only:
            solver.nll(theta.data_ptr(), y.data_ptr(), cfg["eps"], out.data_ptr(), None, None)
        else:
            solver.nll_grad(theta.data_ptr(), y.data_ptr(), cfg["eps"], out.data_ptr(), out.data_ptr() + 8, None, None)
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1))
    res = out.cpu().numpy().copy()
    if ref is None:
        ref = res
    import numpy as np
    print(f"N={n} {'NLL only' if nll_only else 'NLL+grad'} lookahead={w:3d}: {best:9.3f} ms   nll={res[0]:.9f}   max rel diff vs W=0: "
          f"{np.max(np.abs(res - ref) / np.maximum(np.abs(ref), 1e-300)):.2e}", flush=True)
