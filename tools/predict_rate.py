"""Posterior prediction (GP/gp.py:91-120, :226-256) at the BASELINE configurations C2-C4 with the
reference generator's points (tests/golden/ref_c*.npz): seconds per call of predictingFunction_all through the drop-in
class (host buffers in, host results out) with the full posterior covariance and with the variances only, beside the CPU
port of the reference (oracle/: numpy + LAPACK).  One JSON object per line."""
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


def best_of(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


for name in ("ref_c2_poiseuille_additive", "ref_c3_sinusoidal", "ref_c4_drag3d"):
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    if not os.path.exists(path):
        continue
    cfg = synthetic.from_golden(path)
    gp = synthetic.make_model(cfg)
    pargs = (cfg["r_test"], cfg["mu_test"], cfg["r_train"], cfg["delta_y"], cfg["eps"])
    gp.set_constants(*pargs)
    th = cfg["theta0"]
    rec = {"workload": name[4:], "N": int(len(cfg["delta_y"])), "M": int(sum(len(r) for r in cfg["r_test"]))}
    rec["gpu_full_cov_ms"] = 1e3 * best_of(lambda: gp.predictingFunction_all(th, *pargs))
    rec["gpu_variances_only_ms"] = 1e3 * best_of(lambda: gp.predictingFunction_all(th, *pargs, full_cov=False))
    mu, cov = gp.predictingFunction_all(th, *pargs)
    gp.close()
    ref = oracle_for(cfg)
    t0 = time.perf_counter()
    mu_ref, cov_ref = ref.predictingFunction_all(th, *pargs)
    rec["cpu_port_ms"] = 1e3 * (time.perf_counter() - t0)
    rec["cpu_threads"] = os.cpu_count()
    rec["mu_relerr"] = float(max(np.max(np.abs(a - b)) for a, b in zip(mu, mu_ref)) / max(np.max(np.abs(b)) for b in mu_ref))
    print(json.dumps(rec), flush=True)
