"""GPU tests of the "next" rows of SURVEY.md 8(f): the device-resident optimiser loop (n1), diagonal-only and batched
posterior (n3).  The reference behaviour they are held to: solver/optimizers.py:94-263 (through the host loop of
stopro_b200.solver.optimizers, itself checked against the reference's semantics on the CPU) and
test/test_1_sinusoidal_direct_main.py:111-131 (interval_check loop)."""
import numpy as np
import pytest

from stopro_b200 import synthetic
from stopro_b200.solver.optimizers import optimize_by_adam
from stopro_b200.sub_modules.loss_modules import logposterior

pytestmark = pytest.mark.gpu


def _setup(cfg):
    gp = synthetic.make_model(cfg)
    args = (cfg["r_train"], cfg["delta_y"], cfg["eps"])
    gp.set_constants(*args, only_training=True)
    return gp, args


@pytest.mark.parametrize("case", ["poiseuille", "sin1d_noise", "sinusoidal_fixed"])
def test_device_adam_loop_equals_host_loop(cuda_device, case):
    if case == "poiseuille":
        cfg = dict(synthetic.poiseuille(u_num=7, p_num=7, f_num=8, n_test=4, kernel_form="product"), eps=1e-4)
        po = dict(maxiter_GD=40, lr=1e-2, eps=1e-9, loss_ridge_regression=False, index_fixed=None)
    elif case == "sin1d_noise":
        cfg = synthetic.sin_1d_naive()
        po = dict(maxiter_GD=60, lr=2e-2, eps=1e-9, loss_ridge_regression=True, ridge_alpha=1e-3, index_fixed=None)
    else:
        cfg = dict(synthetic.sinusoidal(u_num=10, f_nx=8, f_ny=5, dif_num=6, n_test=4), eps=1e-4)
        po = dict(maxiter_GD=30, lr=1e-2, eps=1e-9, loss_ridge_regression=False, index_fixed=[0, 3, 6])
    gp, args = _setup(cfg)
    f = logposterior(gp.trainingFunction_all, po)
    init = cfg["theta0"].copy()
    dev = optimize_by_adam(f, f.grad, None, init, dict(po, device_loop=True), *args)
    host = optimize_by_adam(f, f.grad, None, init, dict(po, device_loop=False), *args)
    assert len(dev[1]) == len(host[1]) and len(dev[2]) == len(host[2]) and len(dev[3]) == len(host[3])
    assert len(dev[3]) == po["maxiter_GD"]
    assert np.allclose(np.array(dev[2]), np.array(host[2]), rtol=0, atol=1e-9)      # theta trajectories
    assert np.allclose(dev[1], host[1], rtol=1e-10, atol=1e-12)                      # losses (entry 0 = entry 1)
    assert np.allclose(dev[3], host[3], rtol=1e-8, atol=1e-12)                       # gradient norms
    assert dev[1][0] == dev[1][1]
    if po["index_fixed"]:
        assert np.all(np.array(dev[2])[:, po["index_fixed"]] == init[po["index_fixed"]])
    gp.close()


def test_device_adam_loop_stops_on_plateau_and_reports_nan(cuda_device):
    cfg = dict(synthetic.poiseuille(u_num=7, p_num=7, f_num=8, n_test=4, kernel_form="product"), eps=1e-4)
    gp, args = _setup(cfg)
    po = dict(maxiter_GD=400, lr=1e-2, eps=8e-3, loss_ridge_regression=False, index_fixed=None)
    f = logposterior(gp.trainingFunction_all, po)
    dev = optimize_by_adam(f, f.grad, None, cfg["theta0"], dict(po, device_loop=True), *args)
    host = optimize_by_adam(f, f.grad, None, cfg["theta0"], dict(po, device_loop=False), *args)
    assert len(dev[3]) == len(host[3]) < 400                                         # same two-in-a-row plateau step
    assert np.allclose(dev[0], host[0], atol=1e-9)
    bad = cfg["theta0"].copy()
    bad[1] = 40.0                                                                     # length scale e^40: K is singular
    with pytest.raises(Exception):
        optimize_by_adam(f, f.grad, None, bad, dict(po, device_loop=True), args[0], args[1], 0.0)
    gp.close()


def test_batched_and_diagonal_only_posterior(cuda_device):
    cfg = dict(synthetic.sinusoidal(u_num=12, f_nx=10, f_ny=6, dif_num=7, n_test=9), eps=1e-3)
    gp = synthetic.make_model(cfg)
    args = (cfg["r_test"], cfg["mu_test"], cfg["r_train"], cfg["delta_y"], cfg["eps"])
    gp.set_constants(*args)
    rng = np.random.default_rng(3)
    thetas = [cfg["theta0"] + 0.1 * rng.standard_normal(len(cfg["theta0"])) for _ in range(4)]
    mus, vars_ = gp.predict_many(thetas, *args)                                       # one library call
    for th, mu_b, var_b in zip(thetas, mus, vars_):
        mu1, cov1 = gp.predictingFunction_all(th, *args)                              # reference-shaped call, full covariance
        for a, b in zip(mu_b, mu1):
            assert np.array_equal(a, b)
        for v, c in zip(var_b, cov1):
            assert np.max(np.abs(v - np.diag(c))) <= 1e-12 * max(np.max(np.abs(np.diag(c))), 1.0)
    # the diagonal-only kernel against the full test matrix
    import torch
    test = gp._test_plan(cfg["r_test"])
    th = torch.as_tensor(thetas[0], device=cuda_device)
    out = torch.empty(test.rows, dtype=torch.float64, device=cuda_device)
    test.assemble_diag(th.data_ptr(), 0.0, 0, out.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), np.diag(gp.testK_all(thetas[0], cfg["r_test"])))
    gp.close()
