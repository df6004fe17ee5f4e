"""Generate the golden vectors under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

JAX is not installed in this environment, so the reference itself cannot be executed; the vectors come from the
literal restatement of its construction (oracle backend="autodiff": nested torch.func grad / hessian of the scalar
kernel under a double vmap, jacfwd for dK/dtheta, numpy LAPACK with the reference's op sequence).  They pin both
the closed-form oracle (CPU test) and the CUDA path (GPU test) to that restatement.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import oracle_for  # noqa: E402
from stopro_b200 import synthetic  # noqa: E402

CASES = {
    "poiseuille_additive": lambda: dict(synthetic.poiseuille(u_num=5, p_num=5, f_num=5, n_test=4, kernel_form="additive"), eps=1e-4),
    "poiseuille_product": lambda: dict(synthetic.poiseuille(u_num=5, p_num=5, f_num=5, n_test=4, kernel_form="product"), eps=1e-4),
    "sinusoidal": lambda: dict(synthetic.sinusoidal(u_num=5, f_nx=5, f_ny=4, dif_num=4, n_test=4), eps=1e-4),
    "drag3d": lambda: dict(synthetic.drag3d(n_u=3, n_f=4, n_test=5), eps=1e-4),
    "sin1d_naive": lambda: synthetic.sin_1d_naive(n=16, n_test=11),
    "sin1d_laplacian": lambda: synthetic.sin_1d_laplacian(ly_num=10, n_test=11),
    # other live classes of the reference
    "sinusoidal_infer_difp": lambda: dict(synthetic.sinusoidal_without_difp("infer_difp", u_num=5, f_nx=5, f_ny=4, dif_num=4, n_test=4), eps=1e-4),
    "sinusoidal_infer_gov_without_difp": lambda: dict(synthetic.sinusoidal_without_difp("infer_gov_without_difp", u_num=5, f_nx=5, f_ny=4, dif_num=4, n_test=3), eps=1e-4),
    "stokes3d_infer_difp": lambda: dict(synthetic.drag3d_variant("stokes3d_infer_difp", n_u=3, n_f=3, n_test=4), eps=1e-4),
    "stokes2d2c_surface": lambda: dict(synthetic.drag3d_variant("stokes2d2c_surface", n_u=3, n_f=3, n_test=4), eps=1e-3),
}


def theta_of(cfg, seed=7):
    rng = np.random.default_rng(seed)
    th = cfg["theta0"].copy()
    nk = len(th) - (1 if cfg["model_kwargs"].get("index_optimize_noise") else 0)
    th[:nk] += 0.15 * rng.standard_normal(nk)
    return th


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, make in CASES.items():
        if "--only-missing" in sys.argv and os.path.exists(os.path.join(out_dir, f"{name}.npz")):
            continue
        cfg = make()
        gp = oracle_for(cfg, backend="autodiff")
        th = theta_of(cfg)
        thk, _ = gp.split_hyp_and_noise(gp._theta(th))
        args = (cfg["r_train"], cfg["delta_y"], cfg["eps"])
        pargs = (cfg["r_test"], cfg["mu_test"]) + args
        mu, cov = gp.predictingFunction_all(th, *pargs)
        np.savez_compressed(
            os.path.join(out_dir, f"{name}.npz"), theta=th, eps=cfg["eps"],
            K_train=gp._np(gp.trainingK_all(thk, gp._pts(cfg["r_train"]))),
            K_mixed=gp._np(gp.mixedK_all(thk, gp._pts(cfg["r_test"]), gp._pts(cfg["r_train"]))),
            K_test=gp._np(gp.testK_all(thk, gp._pts(cfg["r_test"]))),
            sigma=gp.training_sigma(th, cfg["r_train"], cfg["eps"]),
            nll=gp.trainingFunction_all(th, *args), grad=gp.d_trainingFunction_all(th, *args),
            mu=np.concatenate(mu), var=np.concatenate([np.diag(c) for c in cov]))
        print(name, "N =", len(cfg["delta_y"]))


if __name__ == "__main__":
    main()
