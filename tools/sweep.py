"""BASELINE.json's metric as stated -- "NLL+grad evals/s vs N; K-assembly GB/s; FP64 Cholesky TFLOP/s" -- on one GPU.

    python tools/sweep.py [--sizes 498,1180,...] [--reps 5] [--golden]

For every N: NLL+gradient evaluations/s (device-resident inputs, CUDA events), NLL-only time, the per-class kernel times
of one evaluation (assembly, GEMM, potf2, gradient, misc; serial pass with one event pair per launch), the assembly rate
in GB/s of lower-triangle bytes, the factorisation rate (pigp_potrf_lower alone on an SPD matrix of the padded size) and,
beside it, cuSOLVER potrf / cholesky_inverse through torch on the same matrix (the on-box library comparator of
SURVEY.md 8(d)).  One JSON object per line.  --golden adds the BASELINE configurations C2-C4 at their true sizes with the
reference generator's own points (tests/golden/ref_c*.npz).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from stopro_b200 import _lib, synthetic


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def measure(cfg, label, reps, with_library=True):
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    gp = synthetic.make_model(cfg)
    gp.set_constants(cfg["r_train"], cfg["delta_y"], cfg["eps"], only_training=True)
    solver = gp._solver_for(cfg["r_train"])
    plan = solver.plan
    N, P = plan.rows, plan.theta_len
    theta = torch.as_tensor(cfg["theta0"], device=dev)
    y = torch.as_tensor(cfg["delta_y"], device=dev)
    out = torch.zeros(1 + P, dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)

    def both():
        solver.nll_grad(theta.data_ptr(), y.data_ptr(), cfg["eps"], out.data_ptr(), out.data_ptr() + 8, info.data_ptr(), None)

    def nll_only():
        solver.nll(theta.data_ptr(), y.data_ptr(), cfg["eps"], out.data_ptr(), info.data_ptr(), None)

    ms_both = timeit(both, reps)
    ms_nll = timeit(nll_only, reps)
    finite = bool(torch.isfinite(out).all().item()) and int(info.item()) == 0
    # back-to-back throughput (launch overhead amortised the way an optimiser loop sees it)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = max(3, min(50, int(200.0 / max(ms_both, 0.05))))
    e0.record()
    for _ in range(k):
        both()
    e1.record()
    torch.cuda.synchronize()
    ms_stream = e0.elapsed_time(e1) / k
    _lib.check(lib.pigp_set_side_stream(0))
    both()
    torch.cuda.synchronize()
    _lib.profile_start()
    both()
    prof = _lib.profile_stop()
    _lib.check(lib.pigp_set_side_stream(1))
    rec = {"workload": label, "N": N, "P": P, "finite": finite, "evals_per_s": 1e3 / ms_stream, "ms_nll_grad": ms_both,
           "ms_nll_grad_back_to_back": ms_stream, "ms_nll": ms_nll,
           "tflops_nll_grad": float(N) ** 3 / (ms_stream * 1e-3) * 1e-12,
           "classes_ms": {c: round(v["ms"], 4) for c, v in prof.items()},
           "launches": {c: v["launches"] for c, v in prof.items()},
           "assembly_gb_per_s": 8.0 * N * (N + 1) / 2 / (max(prof["assemble"]["ms"], 1e-9) * 1e-3) * 1e-9,
           "gradient_kernel_gb_per_s": 8.0 * N * (N + 1) / 2 / (max(prof["gradient"]["ms"], 1e-9) * 1e-3) * 1e-9}
    gp.close()
    if with_library:
        n = (N + 127) // 128 * 128
        torch.manual_seed(0)
        X = torch.randn(n, 256, dtype=torch.float64, device=dev)
        S = X @ X.t()
        S.diagonal().add_(float(n))
        del X
        buf = torch.empty_like(S)
        invd = torch.empty(n // 128, 128, 128, dtype=torch.float64, device=dev)

        def ours():
            buf.copy_(S)
            _lib.check(lib.pigp_potrf_lower(buf.data_ptr(), n, n, 0, invd.data_ptr(), None, None))

        t_copy = timeit(lambda: buf.copy_(S), reps)
        t_ours = max(timeit(ours, reps) - t_copy, 1e-6)
        t_lib = timeit(lambda: torch.linalg.cholesky(S), reps)
        L = torch.linalg.cholesky(S)
        t_inv = timeit(lambda: torch.cholesky_inverse(L), max(1, reps // 2))
        rec["potrf"] = {"n": n, "ours_ms": t_ours, "ours_tflops": n ** 3 / 3 / t_ours * 1e-9, "cusolver_ms": t_lib,
                        "cusolver_tflops": n ** 3 / 3 / t_lib * 1e-9, "torch_cholesky_inverse_ms": t_inv,
                        "library_nll_grad_floor_ms": t_lib + t_inv,
                        # the product path beside it: a whole NLL-only evaluation (assembly + factorisation + solve +
                        # log-det) and a whole NLL + gradient evaluation at N <= n
                        "nll_only_eval_ms": ms_nll, "nll_grad_eval_ms": ms_stream}
        del S, buf, L
        torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="498,1180,2640,5018,10570,20000")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--golden", action="store_true")
    ap.add_argument("--no-library", action="store_true")
    a = ap.parse_args()
    if a.golden:
        for name in ("ref_c2_poiseuille_additive", "ref_c2_poiseuille_product", "ref_c3_sinusoidal", "ref_c4_drag3d"):
            path = os.path.join(ROOT, "tests", "golden", name + ".npz")
            if os.path.exists(path):
                cfg = synthetic.from_golden(path)
                print(json.dumps(measure(cfg, name[4:] + " (reference generator's points)", a.reps, not a.no_library)), flush=True)
    for n in [int(s) for s in a.sizes.split(",") if s]:
        cfg = synthetic.stokes2d_scaling(n, n_test=16)
        print(json.dumps(measure(cfg, f"C5 synthetic 2-D Stokes N={n}", a.reps if n < 15000 else 3, not a.no_library)), flush=True)


if __name__ == "__main__":
    main()
