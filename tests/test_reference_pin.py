"""Parity pinned to the reference's own source.

tests/golden/ref_*.npz were produced by EXECUTING /root/reference unmodified (tests/golden/make_golden_ref.py: the
reference's files imported as package `stopro` on the torch-backed jax shim of tests/jax_shim; BASELINE-size inputs from
the reference's own data generators).  Each fixture also carries a numpy.longdouble truth (oracle/extended.py).

  * test_oracle_reproduces_reference_fixture   (CPU)  the closed-form oracle == the executed reference
  * test_cuda_reproduces_reference_fixture     (GPU)  the CUDA path (through the C ABI) == the executed reference
  * test_live_reference_matches_oracle         (CPU, only where /root/reference exists) runs the reference live

Tolerances (north star): 1e-10 relative on K entries, 1e-8 relative on NLL / gradient / predictions.  A value x passes
when |x - reference| <= 1e-8.  Only where that fails -- the eps = 1e-6 Stokes cases, cond(K) = 1e9 .. 1e10, where the
reference's OWN float64 result sits up to 2e-8 from the higher-precision truth and the reference's op sequence run on two
LAPACK builds (torch/MKL in the fixture, numpy/OpenBLAS in the oracle) differs by 1.0e-8 on the C3 NLL -- x is judged
against the long-double truth instead: at least as close to it as REF_SLACK times the reference's own distance, or within
COND_FRAC = 5 % of the first-order forward-error bound cond_2(K) * 2^-53.  All three distances are printed per quantity.
"""
import glob
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.join(HERE, "jax_shim"))
from conftest import oracle_for  # noqa: E402
from stopro_b200 import synthetic  # noqa: E402

K_TOL, F_TOL = 1e-10, 1e-8
REF_SLACK = 3.0   # accepted distance to the truth, in units of the reference's own distance
COND_FRAC = 0.05  # ... or in units of cond_2(K) * u (u = 2^-53)
FIXTURES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "ref_*.npz")))
BIG = {"ref_c3_sinusoidal", "ref_c4_drag3d"}  # seconds of CPU work each for the oracle: still inside the CPU suite


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def config_of(name):
    if name.startswith("ref_c"):
        return synthetic.from_golden(os.path.join(HERE, "golden", name + ".npz"))
    from make_golden_ref import small_cases
    return small_cases()[name]()


def check_matrix(prefix, K, g, tol):
    K = np.asarray(K)
    assert tuple(g[prefix + "_shape"]) == K.shape
    if prefix in g.files:
        want = g[prefix]
        assert np.max(np.abs(K - want)) <= tol * max(np.max(np.abs(want)), 1e-300), prefix
        return
    scale = max(np.max(np.abs(g[prefix + "_val"])), 1e-300)
    assert np.max(np.abs(K.reshape(-1)[g[prefix + "_idx"]] - g[prefix + "_val"])) <= tol * scale, prefix + " samples"
    sec_r = g["sec_tr"] if prefix == "K_train" else g["sec_te"]
    sec_c = g["sec_te"] if prefix == "K_test" else g["sec_tr"]
    s, a = g[prefix + "_blocksum"], g[prefix + "_blockabs"]
    for i in range(len(sec_r) - 1):
        for j in range(len(sec_c) - 1):
            B = K[sec_r[i]:sec_r[i + 1], sec_c[j]:sec_c[j + 1]]
            assert abs(B.sum() - s[i, j]) <= 1e-11 * max(a[i, j], 1e-300) + 1e-300, (prefix, i, j)
            assert abs(np.abs(B).sum() - a[i, j]) <= 1e-11 * max(a[i, j], 1e-300) + 1e-300, (prefix, i, j)
    if prefix + "_diag" in g.files:
        assert np.max(np.abs(np.diag(K) - g[prefix + "_diag"])) <= tol * scale


def accept(label, x, ref, truth, scale, report, cond=0.0):
    """|x - ref| <= F_TOL * scale; failing that, x judged against the long-double truth (module docstring)."""
    x, ref, truth = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64), np.asarray(truth, dtype=np.float64)
    d_ref = np.max(np.abs(x - ref)) / scale
    e_x = np.max(np.abs(x - truth)) / scale
    e_ref = np.max(np.abs(ref - truth)) / scale
    report.append(f"{label}: |x-ref|={d_ref:.2e} |x-truth|={e_x:.2e} |ref-truth|={e_ref:.2e}")
    assert d_ref <= F_TOL or e_x <= max(F_TOL, REF_SLACK * e_ref, COND_FRAC * cond * 2.0 ** -53), report[-1]


def compare_all(name, g, cfg, K_train, K_mixed, K_test, sigma_diag, nll, grad, mu, var, k_tol):
    report = []
    check_matrix("K_train", K_train, g, k_tol)
    check_matrix("K_mixed", K_mixed, g, k_tol)
    check_matrix("K_test", K_test, g, k_tol)
    assert np.max(np.abs(sigma_diag - g["sigma_diag"])) <= k_tol * np.max(np.abs(g["sigma_diag"]))
    cond = float(g["cond"])
    accept("nll", nll, g["nll"], g["truth_nll"], abs(float(g["truth_nll"])), report, cond)
    accept("grad", grad, g["grad"], g["truth_grad"], np.max(np.abs(g["truth_grad"])), report, cond)
    accept("mu", mu, g["mu"], g["truth_mu"], max(np.max(np.abs(g["truth_mu"])), 1e-300), report, cond)
    # the posterior variance K_aa - V^T V is a difference of O(|K_aa|) terms: judged on that scale
    kaa = np.max(np.abs(g["K_test_diag"])) if "K_test_diag" in g.files else np.max(np.abs(np.diag(g["K_test"])))
    accept("var", var, g["var"], g["truth_var"], max(kaa, 1.0), report, cond)
    print(f"\n[{name}] N={int(g['n_train'])} cond={float(g['cond']):.2e}  " + "; ".join(report))


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_reproduces_reference_fixture(name):
    g, cfg = load(name), config_of(name)
    cfg = dict(cfg, eps=float(g["eps"]))
    gp = oracle_for(cfg, "closed")
    th = g["theta"]
    thk, _ = gp.split_hyp_and_noise(th)
    args = (cfg["r_train"], cfg["delta_y"], cfg["eps"])
    S = gp.training_sigma(th, cfg["r_train"], cfg["eps"])
    mu, cov = gp.predictingFunction_all(th, cfg["r_test"], cfg["mu_test"], *args)
    compare_all(name, g, cfg,
                gp.trainingK_all(thk, gp._pts(cfg["r_train"])),
                gp.mixedK_all(thk, gp._pts(cfg["r_test"]), gp._pts(cfg["r_train"])),
                gp.testK_all(thk, gp._pts(cfg["r_test"])), np.diag(S),
                gp.trainingFunction_all(th, *args), gp.d_trainingFunction_all(th, *args),
                np.concatenate(mu), np.concatenate([np.diag(c) for c in cov]), 1e-12)
    if "grad_autodiff_posterior" in g.files:
        # jit(grad(logposterior)) (test_1:76-77) == explicit d_logposterior (GP/gp.py:491-493) in the reference itself
        assert np.max(np.abs(g["grad_autodiff_posterior"] - (g["grad"] + 1.0))) <= 1e-6 * np.max(np.abs(g["grad"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_cuda_reproduces_reference_fixture(cuda_device, name):
    g, cfg = load(name), config_of(name)
    cfg = dict(cfg, eps=float(g["eps"]))
    gp = synthetic.make_model(cfg)
    th, eps = g["theta"], cfg["eps"]
    thk, _ = gp.split_hyp_and_noise(th)
    args = (cfg["r_train"], cfg["delta_y"], eps)
    pargs = (cfg["r_test"], cfg["mu_test"]) + args
    gp.set_constants(*pargs)
    K = gp.trainingK_all(thk, cfg["r_train"])
    assert np.array_equal(K, K.T)
    S = gp.training_sigma(th, cfg["r_train"], eps)
    nll = gp.trainingFunction_all(th, *args)
    grad = gp.d_trainingFunction_all(th, *args)
    mu, var = gp.predictingFunction_all(th, *pargs, full_cov=False)
    compare_all(name, g, cfg, K, gp.mixedK_all(thk, cfg["r_test"], cfg["r_train"]), gp.testK_all(thk, cfg["r_test"]),
                np.diag(S), nll, grad, np.concatenate(mu), np.concatenate(var), K_TOL)
    # full posterior covariance blocks against the reference's (stored in full for small cases, sampled otherwise)
    mu2, cov = gp.predictingFunction_all(th, *pargs)
    kaa = max(np.max(np.abs(np.diag(c))) for c in cov)
    for i, c in enumerate(cov):
        if f"cov_{i}" in g.files:
            d = np.max(np.abs(c - g[f"cov_{i}"]))
        else:
            d = np.max(np.abs(c.reshape(-1)[g[f"cov_{i}_idx"]] - g[f"cov_{i}_val"]))
        assert d <= max(F_TOL, REF_SLACK * float(g["ref_vs_truth"][3]), COND_FRAC * float(g["cond"]) * 2.0 ** -53) * max(kaa, 1.0), (i, d)
    gp.close()


def test_fixtures_present():
    # every BASELINE configuration at its true size, and multi-block noise ranges, are pinned
    for need in ("ref_c1_sin1d_naive", "ref_c2_poiseuille_additive", "ref_c2_poiseuille_product", "ref_c3_sinusoidal",
                 "ref_c4_drag3d", "ref_sinusoidal_noise_blocks_1_2", "ref_drag3d_noise_blocks_3_5"):
        assert need in FIXTURES, need
    sizes = {n: int(load(n)["n_train"]) for n in FIXTURES if n.startswith("ref_c")}
    assert sizes == {"ref_c1_sin1d_naive": 32, "ref_c2_poiseuille_additive": 498, "ref_c2_poiseuille_product": 498,
                     "ref_c3_sinusoidal": 1180, "ref_c4_drag3d": 2640}


def test_live_reference_matches_oracle():
    """Where the reference tree exists (this container, not the GPU box): import it unmodified and compare live."""
    from load_reference import load_reference

    if load_reference() is None:
        pytest.skip("reference tree not present")
    import torch
    from make_golden_ref import ref_model, small_cases, theta_of

    name = "ref_sinusoidal_noise_blocks_1_2"
    cfg = small_cases()[name]()
    th = theta_of(cfg)
    gp = ref_model(cfg)
    T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    r_train, r_test = [T(r) for r in cfg["r_train"]], [T(r) for r in cfg["r_test"]]
    gp.set_constants(r_test, [T(m) for m in cfg["mu_test"]], r_train, T(cfg["delta_y"]), cfg["eps"])
    thk, noise = gp.split_hyp_and_noise(T(th))
    K = gp.trainingK_all(thk, r_train).numpy()
    nll = float(gp.trainingFunction_all(T(th), r_train, T(cfg["delta_y"]), cfg["eps"]))
    g = load(name)
    assert np.array_equal(K, g["K_train"])            # the committed fixture is what the reference computes
    assert nll == float(g["nll"])
    ora = oracle_for(cfg, "closed")
    assert np.max(np.abs(ora.trainingK_all(th[:-1], ora._pts(cfg["r_train"])) - K)) <= 1e-13 * np.max(np.abs(K))
    assert abs(ora.trainingFunction_all(th, cfg["r_train"], cfg["delta_y"], cfg["eps"]) - nll) <= 1e-10 * abs(nll)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ref_c2_poiseuille_product", "ref_c3_sinusoidal"])
def test_two_ranks_reproduce_reference_fixture_at_eps_1e_6(cuda_device, name):
    """The block-cyclic solver on 2 ranks (peer stores + epoch flags, here two ranks on one device) against what the executed
    reference computed for the BASELINE configurations at eps = 1e-6 (cond(K) = 1e9 ... 1e10): NLL and explicit gradient under
    the same acceptance rule as the single-GPU path, and bitwise agreement of the two ranks."""
    from test_gpu_dist import run_ranks

    g, cfg = load(name), config_of(name)
    cfg = dict(cfg, eps=float(g["eps"]))
    assert cfg["eps"] == 1e-6
    gp = synthetic.make_model(cfg)
    out = run_ranks(gp, cfg, g["theta"], 2)
    gp.close()
    (nll, grad, info), (nll1, grad1, info1) = out
    assert info == 0 and info1 == 0
    assert nll == nll1 and np.array_equal(grad, grad1)
    report, cond = [], float(g["cond"])
    accept("nll", nll, g["nll"], g["truth_nll"], abs(float(g["truth_nll"])), report, cond)
    accept("grad", grad, g["grad"], g["truth_grad"], np.max(np.abs(g["truth_grad"])), report, cond)
    print(f"\n[{name}, 2 ranks] N={int(g['n_train'])} cond={cond:.2e}  " + "; ".join(report))
