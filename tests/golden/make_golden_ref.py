"""Generate tests/golden/ref_*.npz by EXECUTING THE REFERENCE ITSELF (/root/reference, unmodified source) on the CPU.

    python tests/golden/make_golden_ref.py [--only NAME[,NAME..]] [--list] [--skip-existing]

JAX is not installed here, so the reference's files are imported byte for byte as package `stopro` on top of
tests/jax_shim (a `jax` / `jax.numpy` stand-in on torch.func, float64).  Everything numerical in a fixture therefore
comes out of the reference's own code: its data generators (data_generator/*.py, through DataPreparer.make_data) give the
BASELINE-size inputs, its GP classes (GP/*.py) give K, the NLL (trainingFunction_all, GP/gp.py:213-224), the explicit
gradient (d_trainingFunction_all, GP/gp.py:412-488: jacfwd re-assembly per hyper-parameter) and the posterior
(predictingFunction_all, GP/gp.py:226-256).  Next to them every fixture holds a numpy.longdouble truth of the same
quantities (oracle/extended.py) so that, for the ill-conditioned eps = 1e-6 cases, the tests can tell which float64
evaluation is closer instead of widening the tolerance.

Runs only where /root/reference exists (this container); the fixtures travel with the repo.  Small cases store the full
matrices; BASELINE-size cases store a seeded sample of 20000 entries plus per-block sums of every matrix.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "jax_shim")):
    if p not in sys.path:
        sys.path.insert(0, p)

from load_reference import load_reference  # noqa: E402

N_SAMPLE = 20000
FULL_MATRIX_LIMIT = 260  # cases with N <= this store K in full


# ---------------------------------------------------------------------------------------------- reference side
def ref_model(cfg):
    """Instantiate the reference's own class for a configuration dict (same keys as stopro_b200.synthetic)."""
    import torch  # noqa: F401
    from stopro.GP.kernels import define_kernel

    kernel = define_kernel(dict(cfg["kernel"]))
    name, kw = cfg["model"], dict(cfg["model_kwargs"])
    if "lbox" in kw and kw["lbox"] is not None:
        import jax.numpy as jnp
        kw["lbox"] = jnp.array(np.asarray(kw["lbox"], dtype=np.float64))
    if name == "naive":
        from stopro.GP.gp_naive import GPmodelNaive
        return GPmodelNaive(Kernel=kernel, index_optimize_noise=kw.get("index_optimize_noise"))
    if name == "laplacian1d":
        from stopro.GP.gp_1D_laplacian import GPmodel1DLaplacian
        return GPmodel1DLaplacian(Kernel=kernel)
    if name == "poiseuille":
        from stopro.GP.gp_poiseuille_independent import GPPoiseuilleIndependent
        return GPPoiseuilleIndependent(Kernel=kernel)
    if name in ("sinusoidal", "sinusoidal_infer_gov"):
        from stopro.GP.gp_sinusoidal_independent import GPSinusoidalWithoutPIndependent
        return GPSinusoidalWithoutPIndependent(Kernel=kernel, **kw)
    if name.startswith("sinusoidal_"):
        import stopro.GP.gp_sinusoidal_infer_difp as m
        cls = dict(sinusoidal_infer_difp=m.GPSinusoidalInferDifP, sinusoidal_infer_u_without_difp=m.GPSinusoidalInferUWithoutDifP,
                   sinusoidal_infer_gov_without_difp=m.GPSinusoidalInferGovWithoutDifP)[name]
        return cls(Kernel=kernel, **kw)
    if name in ("stokes3d", "stokes3d_infer_difp"):
        from stopro.GP.gp_stokes_3D import GPStokes3D
        return GPStokes3D(Kernel=kernel, **kw)
    if name == "stokes3d_naive":
        from stopro.GP.gp_stokes_3D_naive import GPStokes3DNaive
        return GPStokes3DNaive(Kernel=kernel, **kw)
    if name in ("stokes2d2c", "stokes2d2c_surface"):
        import stopro.GP.gp_stokes_3D_2D2C as m
        return dict(stokes2d2c=m.GPStokes2D2C, stokes2d2c_surface=m.GPStokes2D2CSurface)[name](Kernel=kernel, **kw)
    raise KeyError(name)


def from_generator(system, generator_cls, model, model_kwargs, kernel_form=None, modify=None, theta0=None):
    """Inputs of a BASELINE configuration from the reference's own generator (test/test_*_prepare.py flow)."""
    from stopro.data_preparer.data_preparer import DataPreparer

    dp = DataPreparer("/tmp/stopro_ref_proj", system, class_data_generator=generator_cls)
    dp.load_params(system_name=system)
    if modify:
        modify(dp)
    dp.update_params()
    r_train, f_train, r_test, f_test = dp.make_data(show_train_plot=False, show_test_plot=False, save_data=False,
                                                    save_train_plot=False, save_test_plot=False, return_data=True)
    pm = dict(dp.params_main["model"])
    if kernel_form:
        pm["kernel_form"] = kernel_form
    from stopro.sub_modules.init_modules import get_init
    if theta0 is None:
        theta0 = np.asarray(get_init(dict(pm["init_kernel_hyperparameter"]) if isinstance(pm["init_kernel_hyperparameter"], dict)
                                     else list(pm["init_kernel_hyperparameter"]), pm["kernel_type"],
                                     system_type=pm["system_type"]), dtype=np.float64)
    npf = lambda xs: [np.asarray(x, dtype=np.float64) for x in xs]
    r_train, f_train, r_test, f_test = npf(r_train), npf(f_train), npf(r_test), npf(f_test)
    kernel = dict(kernel_type=pm["kernel_type"], kernel_form=pm["kernel_form"], input_dim=pm["input_dim"],
                  distance_func=pm.get("distance_func", False))
    return dict(name=system, model=model, model_kwargs=model_kwargs, kernel=kernel, r_train=r_train, f_train=f_train,
                r_test=r_test, f_test=f_test, mu_test=[np.zeros(len(r)) for r in r_test], delta_y=np.concatenate(f_train),
                theta0=np.asarray(theta0, dtype=np.float64), eps=float(pm["epsilon"]), from_generator=True)


def baseline_cases():
    def c1():
        from stopro.data_generator.sin_1D_naive import Sin1DNaive

        def mod(dp):  # test/test_2_sin_1D_prepare.py:25-37, 61
            dp.params_generate_training["y_num"] = 32
            dp.params_generate_training["sigma2_noise"] = 1.0e-02
            dp.params_main["model"]["index_optimize_noise"] = [0]
            dp.params_main["model"]["init_kernel_hyperparameter"].append(float(np.log(0.0004)))
        np.random.seed(0)
        return from_generator("sin_1D_naive", Sin1DNaive, "naive", dict(index_optimize_noise=[0]), modify=mod)

    def c2(form):
        def f():
            from stopro.data_generator.poiseuille import Poiseuille
            return from_generator("poiseuille", Poiseuille, "poiseuille", {}, kernel_form=form)
        return f

    def c3():
        from stopro.data_generator.sinusoidal import Sinusoidal

        def mod(dp):  # test/test_0_sinusoidal_direct_prepare.py:21-26
            dp.params_main["model"]["init_kernel_hyperparameter"] = {"uxux": [0.0, -1.0, -1.0], "uyuy": [0.0, -1.0, -1.0],
                                                                     "pp": [0.0, -1.0, -1.0]}
        return from_generator("sinusoidal", Sinusoidal, "sinusoidal",
                              dict(lbox=np.array([2.5, 0.0]), use_difp=True, use_difu=True), modify=mod)

    def c4():
        from stopro.data_generator.drag3D import Drag3D

        def mod(dp):  # test/test_10_drag3D_prepare.py:25-26
            dp.params_setting["particle_radius"] = 0.4
            dp.params_generate_test["test_num"] = 40
        return from_generator("drag3D", Drag3D, "stokes3d", {}, modify=mod)

    return {"ref_c1_sin1d_naive": c1, "ref_c2_poiseuille_additive": c2("additive"), "ref_c2_poiseuille_product": c2("product"),
            "ref_c3_sinusoidal": c3, "ref_c4_drag3d": c4}


def small_cases():
    """The small configurations of tests/golden/make_golden.py, plus multi-block noise ranges (GP/gp.py:44-70)."""
    from make_golden import CASES
    from stopro_b200 import synthetic

    out = {"ref_" + k: v for k, v in CASES.items()}

    def noise_sin():
        c = dict(synthetic.sinusoidal(u_num=5, f_nx=5, f_ny=4, dif_num=4, n_test=4), eps=1e-4)
        c["model_kwargs"] = dict(c["model_kwargs"], index_optimize_noise=[1, 2])
        c["theta0"] = np.append(c["theta0"], np.log(3e-2))
        return c

    def noise_3d():
        c = dict(synthetic.drag3d(n_u=3, n_f=4, n_test=5), eps=1e-4)
        c["model_kwargs"] = dict(c["model_kwargs"], index_optimize_noise=[3, 5])
        c["theta0"] = np.append(c["theta0"], np.log(5e-2))
        return c

    out["ref_sinusoidal_noise_blocks_1_2"] = noise_sin
    out["ref_drag3d_noise_blocks_3_5"] = noise_3d
    return out


def theta_of(cfg, seed=7):
    rng = np.random.default_rng(seed)
    th = cfg["theta0"].copy()
    nk = len(th) - (1 if cfg["model_kwargs"].get("index_optimize_noise") else 0)
    th[:nk] += 0.15 * rng.standard_normal(nk)
    return th


def block_sums(K, sec_r, sec_c):
    nr, nc = len(sec_r) - 1, len(sec_c) - 1
    s = np.zeros((nr, nc))
    a = np.zeros((nr, nc))
    for i in range(nr):
        for j in range(nc):
            B = K[sec_r[i]:sec_r[i + 1], sec_c[j]:sec_c[j + 1]]
            s[i, j], a[i, j] = B.sum(), np.abs(B).sum()
    return s, a


def matrix_record(prefix, K, sec_r, sec_c, full, rng):
    K = np.asarray(K)
    rec = {}
    if full:
        rec[prefix] = K
    else:
        flat = rng.integers(0, K.size, N_SAMPLE)
        rec[prefix + "_idx"] = flat
        rec[prefix + "_val"] = K.reshape(-1)[flat]
        s, a = block_sums(K, sec_r, sec_c)
        rec[prefix + "_blocksum"], rec[prefix + "_blockabs"] = s, a
        if K.shape[0] == K.shape[1]:
            rec[prefix + "_diag"] = np.diag(K).copy()
    rec[prefix + "_shape"] = np.array(K.shape)
    return rec


def run_case(name, make, out_dir):
    import torch

    t_start = time.time()
    cfg = make()
    th = theta_of(cfg)
    gp = ref_model(cfg)
    T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    r_train, r_test = [T(r) for r in cfg["r_train"]], [T(r) for r in cfg["r_test"]]
    mu_test = [T(m) for m in cfg["mu_test"]]
    dy, eps, tht = T(cfg["delta_y"]), cfg["eps"], T(th)
    gp.set_constants(r_test, mu_test, r_train, dy, eps)
    N, M = len(cfg["delta_y"]), sum(len(r) for r in cfg["r_test"])
    full = N <= FULL_MATRIX_LIMIT
    rng = np.random.default_rng(12345)
    sec_tr, sec_te = np.asarray(gp.sec_tr), np.asarray(gp.sec_te)
    thk, noise = gp.split_hyp_and_noise(tht)
    rec = dict(theta=th, eps=eps, n_train=N, n_test=M, sec_tr=sec_tr, sec_te=sec_te,
               meta=json.dumps(dict(source="reference executed through tests/jax_shim", model=cfg["model"],
                                    kernel=cfg["kernel"], model_kwargs={k: (np.asarray(v).tolist() if v is not None else None)
                                                                        for k, v in cfg["model_kwargs"].items()})))
    if cfg.get("from_generator"):
        for i, (r, f) in enumerate(zip(cfg["r_train"], cfg["f_train"])):
            rec[f"r_train_{i}"], rec[f"f_train_{i}"] = r, f
        for i, (r, f) in enumerate(zip(cfg["r_test"], cfg["f_test"])):
            rec[f"r_test_{i}"], rec[f"f_test_{i}"] = r, f
        rec["theta0"] = cfg["theta0"]
    t0 = time.time()
    K = gp.trainingK_all(thk, r_train)
    print(f"  {name}: N={N} M={M}  trainingK_all {time.time() - t0:.1f}s", flush=True)
    rec.update(matrix_record("K_train", K.numpy(), sec_tr, sec_tr, full, rng))
    S = gp.add_eps_to_sigma(K, eps, noise_parameter=noise).numpy()
    rec["sigma_diag"] = np.diag(S).copy()
    if full:
        rec["sigma"] = S
    rec["cond"] = float(np.linalg.cond(S))
    del K, S
    Kab = gp.mixedK_all(thk, r_test, r_train)
    rec.update(matrix_record("K_mixed", Kab.numpy(), sec_te, sec_tr, full, rng))
    del Kab
    Kaa = gp.testK_all(thk, r_test)
    rec.update(matrix_record("K_test", Kaa.numpy(), sec_te, sec_te, full, rng))
    del Kaa
    args = (r_train, dy, eps)
    t0 = time.time()
    rec["nll"] = float(gp.trainingFunction_all(tht, *args))
    rec["nll_theta0"] = float(gp.trainingFunction_all(T(cfg["theta0"]), *args))
    print(f"  {name}: NLL {rec['nll']:.12g}  ({time.time() - t0:.1f}s)  cond {rec['cond']:.2e}", flush=True)
    t0 = time.time()
    rec["grad"] = gp.d_trainingFunction_all(tht, *args).numpy()  # explicit dK/dtheta path (GP/gp.py:412-488)
    print(f"  {name}: d_trainingFunction_all {time.time() - t0:.1f}s", flush=True)
    if N <= 600:
        import jax
        from stopro.sub_modules.loss_modules import logposterior
        func = logposterior(gp.trainingFunction_all, {"loss_ridge_regression": False})
        rec["grad_autodiff_posterior"] = jax.grad(func, 0)(tht, *args).numpy()  # jit(grad(func, 0)), test_1:77
        rec["logposterior"] = float(func(tht, *args))
    t0 = time.time()
    mu, cov = gp.predictingFunction_all(tht, r_test, [m.clone() for m in mu_test], *args)
    rec["mu"] = np.concatenate([m.numpy() for m in mu])
    rec["var"] = np.concatenate([np.diag(c.numpy()) for c in cov])
    if full:
        for i, c in enumerate(cov):
            rec[f"cov_{i}"] = c.numpy()
    else:
        for i, c in enumerate(cov):
            c = c.numpy()
            flat = rng.integers(0, c.size, 4000)
            rec[f"cov_{i}_idx"], rec[f"cov_{i}_val"] = flat, c.reshape(-1)[flat]
    print(f"  {name}: predictingFunction_all {time.time() - t0:.1f}s", flush=True)
    del mu, cov

    # higher-precision truth (oracle/extended.py)
    from oracle import extended
    from oracle.gp_ref import GPRef
    kw = cfg["model_kwargs"]
    ld = GPRef(cfg["model"], kernel_form=cfg["kernel"]["kernel_form"], dim=cfg["kernel"]["input_dim"], lbox=kw.get("lbox"),
               index_optimize_noise=kw.get("index_optimize_noise"), dtype=np.longdouble)
    ld.set_constants(cfg["r_test"], cfg["mu_test"], cfg["r_train"], cfg["delta_y"], eps)
    t0 = time.time()
    tr = extended.truth(ld, th, cfg["r_train"], cfg["delta_y"], eps, r_test=cfg["r_test"])
    print(f"  {name}: longdouble truth {time.time() - t0:.1f}s", flush=True)
    for k in ("nll", "grad", "mu", "var"):
        rec["truth_" + k] = tr[k]
    e_n = abs(rec["nll"] - tr["nll"]) / abs(tr["nll"])
    e_g = np.max(np.abs(rec["grad"] - tr["grad"])) / np.max(np.abs(tr["grad"]))
    e_m = np.max(np.abs(rec["mu"] - tr["mu"])) / max(np.max(np.abs(tr["mu"])), 1e-300)
    e_v = np.max(np.abs(rec["var"] - tr["var"]))
    rec["ref_vs_truth"] = np.array([e_n, e_g, e_m, e_v])
    print(f"  {name}: reference(float64) vs truth: nll {e_n:.2e} grad {e_g:.2e} mu {e_m:.2e} var(abs) {e_v:.2e}", flush=True)
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
    print(f"{name}: done in {time.time() - t_start:.0f}s, {os.path.getsize(os.path.join(out_dir, name + '.npz')) / 1e3:.0f} kB", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--skip-existing", action="store_true")
    a = ap.parse_args()
    if load_reference() is None:
        raise SystemExit("the reference tree is not available here: fixtures can only be regenerated where /root/reference exists")
    import torch
    torch.set_num_threads(max(1, (os.cpu_count() or 2) - 1))
    cases = {}
    cases.update(small_cases())
    cases.update(baseline_cases())
    if a.list:
        print("\n".join(cases))
        return
    only = [s for s in a.only.split(",") if s]
    for name, make in cases.items():
        if only and name not in only:
            continue
        if a.skip_existing and os.path.exists(os.path.join(HERE, name + ".npz")):
            continue
        run_case(name, make, HERE)


if __name__ == "__main__":
    main()
