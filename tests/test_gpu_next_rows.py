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


def test_block_with_four_hyperparameter_groups(cuda_device):
    """A block may draw on every hyper-parameter group of the model (PIGP_MAX_GROUPS = 4): K and the trace gradient of a
    hand-made 3-D block  sum_g c_g D_g k_g  equal the sum of the four single-group blocks (the old evaluator
    silently truncated a block to three groups)."""
    import ctypes as C

    import torch

    from stopro_b200 import _lib, operators

    lib = _lib.lib()
    rng = np.random.default_rng(5)
    n, dim, ng = 200, 3, 4
    pts = np.ascontiguousarray(rng.random((n, dim)))
    sec = np.array([0, n], dtype=np.int64)
    orders = [(0, 0, 0), (2, 0, 0), (1, 1, 0), (0, 2, 2)]          # one monomial per group, mixed parities
    coefs = [1.0, -0.7, 0.4, 0.25]

    def make_plan(terms):
        table = (_lib.BlockDesc * 1)()
        table[0].n_terms = len(terms)
        for k, (g, c, o) in enumerate(terms):
            table[0].terms[k].group, table[0].terms[k].coef = g, c
            for d in range(3):
                table[0].terms[k].order[d] = o[d]
        d = _lib.PlanDesc()
        d.dim, d.product_form, d.n_groups, d.symmetric, d.n_row_blocks, d.n_col_blocks = dim, 1, ng, 1, 1, 1
        d.sec_row = sec.ctypes.data_as(C.POINTER(C.c_int64))
        d.sec_col = d.sec_row
        d.pts_row_host = pts.ctypes.data_as(C.POINTER(C.c_double))
        d.pts_col_host = d.pts_row_host
        d.table = table
        d.noise_lo_block = d.noise_hi_block = -1
        h = C.c_void_p()
        _lib.check(lib.pigp_plan_create(C.byref(d), C.byref(h)))
        return h, table

    theta = rng.normal(0.0, 0.3, ng * (1 + dim))
    th_dev = torch.as_tensor(theta, device=cuda_device)

    def assemble(h):
        K = torch.empty(n, n, dtype=torch.float64, device=cuda_device)
        _lib.check(lib.pigp_assemble(h, th_dev.data_ptr(), 0.0, 0, K.data_ptr(), n, _lib.LAYOUT_FULL, None))
        torch.cuda.synchronize()
        return K.cpu().numpy()

    h_all, keep = make_plan([(g, coefs[g], orders[g]) for g in range(ng)])
    K_all = assemble(h_all)
    K_sum = np.zeros((n, n))
    for g in range(ng):
        h_g, keep_g = make_plan([(g, coefs[g], orders[g])])
        K_sum += assemble(h_g)
        lib.pigp_plan_destroy(h_g)
    assert np.max(np.abs(K_all - K_sum)) <= 1e-13 * np.max(np.abs(K_sum))
    # every group's gamma really is its own: scaling gamma_3 scales only the fourth monomial's share
    theta2 = theta.copy()
    theta2[3 * (1 + dim)] += np.log(2.0)
    th_dev.copy_(torch.as_tensor(theta2))
    h_3, keep_3 = make_plan([(3, coefs[3], orders[3])])
    th_dev.copy_(torch.as_tensor(theta))
    K_3 = assemble(h_3)
    th_dev.copy_(torch.as_tensor(theta2))
    assert np.max(np.abs(assemble(h_all) - (K_all + K_3))) <= 1e-13 * np.max(np.abs(K_all))
    lib.pigp_plan_destroy(h_3)
    lib.pigp_plan_destroy(h_all)
