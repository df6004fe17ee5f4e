"""Optimiser steps per second at the BASELINE sizes C2 (N = 498), C3 (N = 1180) and C4 (N = 2640), reference generator's
points (tests/golden/ref_c*.npz): the device-resident Adam loop (pigp_adam_host), the host loop over the same library
(one ctypes call + synchronisation per step) and the CPU port of the reference's step (func + dfunc of the explicit-
derivative scripts, oracle/: numpy + LAPACK).  One JSON object per line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from conftest import oracle_for
from stopro_b200 import synthetic
from stopro_b200.solver.optimizers import optimize_by_adam
from stopro_b200.sub_modules.loss_modules import logposterior

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for name in ("ref_c2_poiseuille_additive", "ref_c3_sinusoidal", "ref_c4_drag3d"):
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    if not os.path.exists(path):
        continue
    cfg = synthetic.from_golden(path)
    gp = synthetic.make_model(cfg)
    args = (cfg["r_train"], cfg["delta_y"], cfg["eps"])
    gp.set_constants(*args, only_training=True)
    po = dict(maxiter_GD=steps, lr=1e-2, eps=0.0, loss_ridge_regression=False, index_fixed=None, print_process=False)
    f = logposterior(gp.trainingFunction_all, po)
    rec = {"workload": name[4:], "N": int(len(cfg["delta_y"])), "steps": steps}
    for label, dev_loop in (("device_loop", True), ("host_loop", False)):
        optimize_by_adam(f, f.grad, None, cfg["theta0"], dict(po, maxiter_GD=5, device_loop=dev_loop), *args)  # warm-up
        t0 = time.perf_counter()
        out = optimize_by_adam(f, f.grad, None, cfg["theta0"], dict(po, device_loop=dev_loop), *args)
        dt = time.perf_counter() - t0
        rec[label + "_steps_per_s"] = len(out[3]) / dt
        rec[label + "_final_loss"] = float(out[1][-1])
    gp.close()
    ref = oracle_for(cfg)
    t0 = time.perf_counter()
    ref.trainingFunction_all(cfg["theta0"], *args)
    ref.d_trainingFunction_all(cfg["theta0"], *args)
    rec["cpu_port_steps_per_s"] = 1.0 / (time.perf_counter() - t0)
    rec["cpu_threads"] = os.cpu_count()
    print(json.dumps(rec), flush=True)
