"""Matern-5/2, 7/2, 9/2 kernels (GP/kernels.py:127-205) under the reference's operators -- SURVEY.md 8(f) row n4.

  * CPU, where /root/reference exists: the closed-form oracle (oracle/matern_ref.py) against the reference's own block
    functions executed through tests/jax_shim, including coincident coordinates (the reference's autodiff through jnp.abs
    gives 0 for every derivative of order >= 1 at s = 0; parity reproduces that), and against its NLL / gradient /
    posterior for the plain GP.
  * GPU: the CUDA evaluator (csrc/pigp_matern.cu, through the C ABI) against that oracle.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "jax_shim"))
sys.path.insert(0, os.path.join(HERE, "golden"))
from conftest import oracle_for  # noqa: E402
from stopro_b200 import operators, synthetic  # noqa: E402

KINDS = ("mt52", "mt72", "mt92")


def grid_and_random_points(rng, dim, n_grid=3, n_rand=5):
    g = np.linspace(0.0, 1.0, n_grid)
    mesh = np.stack([m.ravel() for m in np.meshgrid(*([g] * dim))], 1)   # many pairs share a coordinate: s_d = 0 exactly
    return mesh, np.concatenate([mesh[:4], rng.random((n_rand, dim))])


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("form", ["product", "additive"])
def test_oracle_matches_the_executed_reference_blocks(kind, form):
    from load_reference import load_reference

    if load_reference() is None:
        pytest.skip("reference tree not present")
    import torch
    from oracle import matern_ref
    from stopro.GP.gp_poiseuille_independent import GPPoiseuilleIndependent
    from stopro.GP.kernels import define_kernel

    rng = np.random.default_rng(1)
    gp = GPPoiseuilleIndependent(Kernel=define_kernel(dict(kernel_type=kind, kernel_form=form, distance_func=False, input_dim=2)))
    obs, fields = operators.stokes_observables(2)
    r, rp = grid_and_random_points(rng, 2)
    theta = rng.normal(0.0, 0.3, 9)
    T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    checked = 0
    for na in ("ux", "uy", "p", "fx", "fy", "div"):
        for nb in ("ux", "uy", "p", "fx", "fy", "div"):
            if not hasattr(gp, f"K{na}{nb}"):
                continue
            want = getattr(gp, f"K{na}{nb}")(T(r), T(rp), T(theta)).numpy()
            got = matern_ref.eval_terms(kind, operators.block_terms(obs[na], obs[nb], fields, 2, form == "product"), r, rp,
                                        theta, 2, form == "product")
            assert np.max(np.abs(got - want)) <= 1e-12 * max(np.max(np.abs(want)), 1e-12), (na, nb)
            checked += 1
    assert checked >= 30


def naive_matern_config(kind, dim=2, n=40, m=15, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.random((n, dim))
    xt = rng.random((m, dim))
    y = np.sin(3.0 * x[:, 0]) * np.cos(2.0 * x[:, 1]) + 0.05 * rng.standard_normal(n)
    kernel = dict(kernel_type=kind, kernel_form="product", input_dim=dim, distance_func=False)
    return synthetic._pack("naive_" + kind, "naive", dict(index_optimize_noise=[0]), kernel, [x], [y], [xt], [np.zeros(m)],
                           [0.2, -0.5, -0.3, np.log(1e-2)], eps=1e-6)


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_matches_the_executed_reference_gp(kind):
    from load_reference import load_reference

    if load_reference() is None:
        pytest.skip("reference tree not present")
    import torch
    from make_golden_ref import ref_model

    cfg = naive_matern_config(kind)
    gp = ref_model(cfg)
    T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    r_train, r_test = [T(r) for r in cfg["r_train"]], [T(r) for r in cfg["r_test"]]
    dy, th = T(cfg["delta_y"]), T(cfg["theta0"])
    gp.set_constants(r_test, [T(m) for m in cfg["mu_test"]], r_train, dy, cfg["eps"])
    nll = float(gp.trainingFunction_all(th, r_train, dy, cfg["eps"]))
    grad = gp.d_trainingFunction_all(th, r_train, dy, cfg["eps"]).numpy()
    mu, cov = gp.predictingFunction_all(th, r_test, [T(m) for m in cfg["mu_test"]], r_train, dy, cfg["eps"])
    ora = oracle_for(cfg)
    args = (cfg["r_train"], cfg["delta_y"], cfg["eps"])
    assert abs(ora.trainingFunction_all(cfg["theta0"], *args) - nll) <= 1e-10 * abs(nll)
    assert np.max(np.abs(ora.d_trainingFunction_all(cfg["theta0"], *args) - grad)) <= 1e-8 * np.max(np.abs(grad))
    mu_o, cov_o = ora.predictingFunction_all(cfg["theta0"], cfg["r_test"], cfg["mu_test"], *args)
    assert np.max(np.abs(mu_o[0] - mu[0].numpy())) <= 1e-9 * np.max(np.abs(mu[0].numpy()))
    assert np.max(np.abs(cov_o[0] - cov[0].numpy())) <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("form", ["product", "additive"])
def test_cuda_blocks_match_oracle(cuda_device, kind, form):
    """Every block of the Poiseuille table (up to 4th derivatives), training / mixed / test matrices, grid points."""
    cfg = dict(synthetic.poiseuille(u_num=5, p_num=5, f_num=4, n_test=3, kernel_form=form), eps=1e-3)
    cfg["kernel"] = dict(cfg["kernel"], kernel_type=kind)
    gp = synthetic.make_model(cfg)
    ref = oracle_for(cfg)
    rng = np.random.default_rng(2)
    th = cfg["theta0"] + 0.2 * rng.standard_normal(9)
    gp.set_constants(cfg["r_test"], cfg["mu_test"], cfg["r_train"], cfg["delta_y"], cfg["eps"])
    for ours, want in ((gp.trainingK_all(th, cfg["r_train"]), ref.trainingK_all(th, ref._pts(cfg["r_train"]))),
                       (gp.mixedK_all(th, cfg["r_test"], cfg["r_train"]), ref.mixedK_all(th, ref._pts(cfg["r_test"]), ref._pts(cfg["r_train"]))),
                       (gp.testK_all(th, cfg["r_test"]), ref.testK_all(th, ref._pts(cfg["r_test"])))):
        assert ours.shape == want.shape
        assert np.max(np.abs(ours - want)) <= 1e-10 * np.max(np.abs(want))
    K = gp.trainingK_all(th, cfg["r_train"])
    assert np.array_equal(K, K.T)
    gp.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,dim", [("mt52", 2), ("mt72", 2), ("mt92", 2), ("mt92", 3)])
def test_cuda_gp_matches_oracle(cuda_device, kind, dim):
    """NLL, dK/dtheta trace gradient (incl. the noise parameter) and posterior of the plain GP with a Matern kernel."""
    cfg = naive_matern_config(kind, dim=dim)
    if dim == 3:
        cfg["theta0"] = np.array([0.2, -0.5, -0.3, -0.4, np.log(1e-2)])
    gp = synthetic.make_model(cfg)
    ref = oracle_for(cfg)
    args = (cfg["r_train"], cfg["delta_y"], cfg["eps"])
    pargs = (cfg["r_test"], cfg["mu_test"]) + args
    gp.set_constants(*pargs)
    th = cfg["theta0"]
    nll, grad = gp.value_and_grad(th, *args)
    nll_ref, g_ref = ref.trainingFunction_all(th, *args), ref.d_trainingFunction_all(th, *args)
    assert abs(nll - nll_ref) <= 1e-8 * abs(nll_ref)
    assert np.max(np.abs(grad - g_ref)) <= 1e-8 * np.max(np.abs(g_ref))
    mu, cov = gp.predictingFunction_all(th, *pargs)
    mu_ref, cov_ref = ref.predictingFunction_all(th, *pargs)
    assert np.max(np.abs(mu[0] - mu_ref[0])) <= 1e-8 * np.max(np.abs(mu_ref[0]))
    assert np.max(np.abs(cov[0] - cov_ref[0])) <= 1e-8 * max(np.max(np.abs(cov_ref[0])), 1.0)
    gp.close()
