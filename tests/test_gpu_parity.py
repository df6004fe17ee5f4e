"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances are the north star's: 1e-10 relative on K entries (relative to the largest entry of the matrix),
1e-8 relative on NLL, gradient and predictions.
"""
import numpy as np
import pytest

from conftest import oracle_for
from stopro_b200 import synthetic

pytestmark = pytest.mark.gpu

K_TOL = 1e-10
F_TOL = 1e-8

CONFIGS = {
    "sin1d_naive": lambda: synthetic.sin_1d_naive(),
    "sin1d_laplacian": lambda: synthetic.sin_1d_laplacian(),
    "poiseuille_additive": lambda: synthetic.poiseuille(kernel_form="additive"),
    "poiseuille_product": lambda: synthetic.poiseuille(kernel_form="product"),
    "sinusoidal": lambda: synthetic.sinusoidal(u_num=16, f_nx=14, f_ny=8, dif_num=9, n_test=10),
    "drag3d": lambda: synthetic.drag3d(n_u=4, n_f=5, n_test=12),
    # larger jitter (params_model["epsilon"] is an input): well conditioned, so the strict 1e-8 bar applies
    "poiseuille_additive_eps1e-2": lambda: dict(synthetic.poiseuille(kernel_form="additive"), eps=1e-2),
    "poiseuille_product_eps1e-2": lambda: dict(synthetic.poiseuille(kernel_form="product"), eps=1e-2),
    "sinusoidal_eps1e-2": lambda: dict(synthetic.sinusoidal(u_num=16, f_nx=14, f_ny=8, dif_num=9, n_test=10), eps=1e-2),
    # the other live classes of the reference, through the same descriptor compiler
    "sinusoidal_infer_gov": lambda: dict(synthetic.sinusoidal(u_num=12, f_nx=10, f_ny=6, dif_num=7, n_test=8), eps=1e-2,
                                         model="sinusoidal_infer_gov",
                                         model_kwargs=dict(lbox=np.array([2.5, 0.0]), use_difp=True, use_difu=True,
                                                           infer_governing_eqs=True),
                                         **_gov_test(8)),
    "sinusoidal_infer_difp": lambda: dict(synthetic.sinusoidal_without_difp("infer_difp", u_num=12, f_nx=10, f_ny=6, dif_num=7, n_test=8), eps=1e-2),
    "sinusoidal_infer_u_without_difp": lambda: dict(synthetic.sinusoidal_without_difp("infer_u_without_difp", u_num=12, f_nx=10, f_ny=6, dif_num=7, n_test=8), eps=1e-2),
    "sinusoidal_infer_gov_without_difp": lambda: dict(synthetic.sinusoidal_without_difp("infer_gov_without_difp", u_num=12, f_nx=10, f_ny=6, dif_num=7, n_test=8), eps=1e-2),
    "stokes3d_infer_difp": lambda: dict(synthetic.drag3d_variant("stokes3d_infer_difp"), eps=1e-2),
    "stokes3d_naive": lambda: dict(synthetic.drag3d_variant("stokes3d_naive"), eps=1e-2),
    "stokes2d2c": lambda: dict(synthetic.drag3d_variant("stokes2d2c"), eps=1e-2),
    "stokes2d2c_surface": lambda: dict(synthetic.drag3d_variant("stokes2d2c_surface"), eps=1e-2),
}


def _gov_test(n_test):
    """test arrays for infer_governing_eqs: three blocks [fx, fy, div] on the sinusoidal test points"""
    base = synthetic.sinusoidal(u_num=12, f_nx=10, f_ny=6, dif_num=7, n_test=n_test)
    r_t = base["r_test"][0]
    return dict(r_test=[r_t, r_t.copy(), r_t.copy()], mu_test=[np.zeros(len(r_t))] * 3,
                f_test=[np.full(len(r_t), 12.0), np.zeros(len(r_t)), np.zeros(len(r_t))])

STRICT = {"sin1d_naive", "sin1d_laplacian", "drag3d", "poiseuille_additive_eps1e-2", "poiseuille_product_eps1e-2",
          "sinusoidal_eps1e-2"}
U = 2.0 ** -52
COND_FRAC = 0.05  # of the first-order forward-error bound cond(K) * u (the rule of tests/test_reference_pin.py)


def solve_tol(ref, cfg, th, name):
    """1e-8 where the conditioning allows it.  Two correct FP64 evaluations of K differ by ~u per entry, and a
    solve against K amplifies that by up to cond(K): with eps = 1e-6 the schema-faithful 2-D Stokes instances have
    cond(K) ~ 1e9, where no pair of independent FP64 implementations can agree to 1e-8.
    Those cases are held to 5 % of cond(K) * u instead (the reference-pinned fixtures of tests/test_reference_pin.py, with
    their long-double truth, are the authoritative gate for them); the STRICT cases must be well conditioned enough for
    1e-8."""
    cond = np.linalg.cond(ref.training_sigma(th, cfg["r_train"], cfg["eps"]))
    if name in STRICT:
        assert cond * U < F_TOL, f"{name}: cond={cond:.2e} is too large for a strict case"
        return F_TOL
    return max(F_TOL, COND_FRAC * cond * U)


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def theta_for(cfg, seed=1):
    rng = np.random.default_rng(seed)
    th = cfg["theta0"].copy()
    nk = len(th) - (1 if cfg["model_kwargs"].get("index_optimize_noise") else 0)
    th[:nk] += 0.2 * rng.standard_normal(nk)  # generic theta: every group gets its own gamma and length scales
    return th


@pytest.mark.parametrize("name", list(CONFIGS))
def test_covariance_blocks(cuda_device, name):
    cfg = CONFIGS[name]()
    gp = synthetic.make_model(cfg)
    ref = oracle_for(cfg)
    th = theta_for(cfg)
    thk, _ = ref.split_hyp_and_noise(th)
    gp.set_constants(cfg["r_test"], cfg["mu_test"], cfg["r_train"], cfg["delta_y"], cfg["eps"])
    K = gp.trainingK_all(thk, cfg["r_train"])
    Kref = ref.trainingK_all(thk, ref._pts(cfg["r_train"]))
    assert K.shape == Kref.shape
    assert relerr(K, Kref) < K_TOL
    assert np.array_equal(K, K.T)
    Kab = gp.mixedK_all(thk, cfg["r_test"], cfg["r_train"])
    assert relerr(Kab, ref.mixedK_all(thk, ref._pts(cfg["r_test"]), ref._pts(cfg["r_train"]))) < K_TOL
    Kaa = gp.testK_all(thk, cfg["r_test"])
    assert relerr(Kaa, ref.testK_all(thk, ref._pts(cfg["r_test"]))) < K_TOL
    S = gp.training_sigma(th, cfg["r_train"], cfg["eps"])
    assert relerr(S, ref.training_sigma(th, cfg["r_train"], cfg["eps"])) < K_TOL


@pytest.mark.parametrize("name", list(CONFIGS))
def test_nll_and_gradient(cuda_device, name):
    cfg = CONFIGS[name]()
    gp = synthetic.make_model(cfg)
    ref = oracle_for(cfg)
    th = theta_for(cfg)
    args = (cfg["r_train"], cfg["delta_y"], cfg["eps"])
    gp.set_constants(*args, only_training=True)
    tol = solve_tol(ref, cfg, th, name)
    nll = gp.trainingFunction_all(th, *args)
    nll_ref = ref.trainingFunction_all(th, *args)
    assert abs(nll - nll_ref) <= tol * abs(nll_ref)
    g = gp.d_trainingFunction_all(th, *args)
    g_ref = ref.d_trainingFunction_all(th, *args)
    assert np.max(np.abs(g - g_ref)) <= tol * np.max(np.abs(g_ref))
    assert np.allclose(gp.d_logposterior(th, *args), g + 1.0)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_prediction(cuda_device, name):
    cfg = CONFIGS[name]()
    gp = synthetic.make_model(cfg)
    ref = oracle_for(cfg)
    th = theta_for(cfg)
    args = (cfg["r_test"], cfg["mu_test"], cfg["r_train"], cfg["delta_y"], cfg["eps"])
    gp.set_constants(*args)
    tol = solve_tol(ref, cfg, th, name)
    mu, cov = gp.predictingFunction_all(th, *args)
    mu_ref, cov_ref = ref.predictingFunction_all(th, *args)
    scale_mu = max(np.max(np.abs(m)) for m in mu_ref)
    # the posterior covariance K_aa - V^T V is a difference of O(|K_aa|) terms: errors are judged on that scale
    thk, _ = ref.split_hyp_and_noise(th)
    scale_cov = max(np.max(np.abs(ref.testK_all(thk, ref._pts(cfg["r_test"])))), max(np.max(np.abs(c)) for c in cov_ref))
    for a, b in zip(mu, mu_ref):
        assert np.max(np.abs(a - b)) <= tol * scale_mu
    for a, b in zip(cov, cov_ref):
        assert a.shape == b.shape
        assert np.max(np.abs(a - b)) <= tol * scale_cov
    _, var = gp.predictingFunction_all(th, *args, full_cov=False)
    for v, c in zip(var, cov):
        assert np.max(np.abs(v - np.diag(c))) <= tol * scale_cov


def test_dense_building_blocks(cuda_device):
    import torch
    from stopro_b200 import _lib

    lib = _lib.lib()
    torch.manual_seed(0)
    n, m = 384, 256
    # GEMM, all four operand orientations
    A = torch.randn(m, n, dtype=torch.float64, device=cuda_device)
    B = torch.randn(m, n, dtype=torch.float64, device=cuda_device)
    for akc in (1, 0):
        for bkc in (1, 0):
            Am = A if akc else A.t().contiguous()     # A(m,k): [m][k] or [k][m]
            Bm = B if bkc else B.t().contiguous()
            Cm = torch.randn(m, m, dtype=torch.float64, device=cuda_device)
            want = -0.5 * A @ B.t() + 2.0 * Cm
            _lib.check(lib.pigp_dgemm(m, m, n, -0.5, Am.data_ptr(), Am.stride(0), akc, Bm.data_ptr(), Bm.stride(0), bkc,
                                      2.0, Cm.data_ptr(), m, 0, None))
            torch.cuda.synchronize()
            assert relerr(Cm.cpu().numpy(), want.cpu().numpy()) < 1e-13
    # Cholesky with extra rows, and the inverse
    n = 640
    X = torch.randn(n, n + 50, dtype=torch.float64, device=cuda_device)
    S = X @ X.t() / n + torch.eye(n, dtype=torch.float64, device=cuda_device)
    E = torch.randn(128, n, dtype=torch.float64, device=cuda_device)
    buf = torch.cat([S, E], 0).contiguous()
    invd = torch.empty(n // 128, 128, 128, dtype=torch.float64, device=cuda_device)
    info = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    _lib.check(lib.pigp_potrf_lower(buf.data_ptr(), n, n, 128, invd.data_ptr(), info.data_ptr(), None))
    torch.cuda.synchronize()
    assert int(info.item()) == 0
    L = torch.tril(buf[:n])
    Lref = torch.linalg.cholesky(S)
    assert relerr(L.cpu().numpy(), Lref.cpu().numpy()) < 1e-12
    Eref = torch.linalg.solve_triangular(Lref, E.t(), upper=False).t()
    assert relerr(buf[n:].cpu().numpy(), Eref.cpu().numpy()) < 1e-11
    W = torch.zeros(n, n, dtype=torch.float64, device=cuda_device)
    Xo = torch.zeros(n, n, dtype=torch.float64, device=cuda_device)
    _lib.check(lib.pigp_potri_lower(buf.data_ptr(), n, n, invd.data_ptr(), W.data_ptr(), Xo.data_ptr(), None))
    torch.cuda.synchronize()
    Sinv = torch.linalg.inv(S)
    assert relerr(torch.tril(W).cpu().numpy(), torch.linalg.inv(Lref).cpu().numpy()) < 1e-11
    assert relerr(torch.tril(Xo).cpu().numpy(), torch.tril(Sinv).cpu().numpy()) < 1e-11
    # non-positive-definite input: info set, NaNs out (what jnp.linalg.cholesky gives the reference)
    bad = -torch.eye(128, dtype=torch.float64, device=cuda_device)
    _lib.check(lib.pigp_potrf_lower(bad.data_ptr(), 128, 128, 0, invd.data_ptr(), info.data_ptr(), None))
    torch.cuda.synchronize()
    assert int(info.item()) == 1 and torch.isnan(bad[0, 0])


def test_panel_schedule_of_the_standalone_factorisation(cuda_device):
    """pigp_potrf_lower takes the panel schedule with look-ahead from 12 tiles on (two streams, coarse panels): factor, extra
    rows and inverse tiles against torch / cuSOLVER at 16 tiles, in a caller-provided (uninitialised) workspace."""
    import torch

    from stopro_b200 import _lib

    lib = _lib.lib()
    n, extra = 2048, 256
    g = torch.Generator(device=cuda_device).manual_seed(5)
    X = torch.randn(n, n + 64, dtype=torch.float64, device=cuda_device, generator=g)
    S = X @ X.t() / n + torch.eye(n, dtype=torch.float64, device=cuda_device)
    E = torch.randn(extra, n, dtype=torch.float64, device=cuda_device, generator=g)
    buf = torch.cat([S, E], 0).contiguous()
    invd = torch.full((n // 128, 128, 128), float("nan"), dtype=torch.float64, device=cuda_device)
    info = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    for _ in range(2):  # twice: the second call reuses the cached bulk stream and events
        buf.copy_(torch.cat([S, E], 0))
        _lib.check(lib.pigp_potrf_lower(buf.data_ptr(), n, n, extra, invd.data_ptr(), info.data_ptr(), None))
    torch.cuda.synchronize()
    assert int(info.item()) == 0
    Lref = torch.linalg.cholesky(S)
    assert relerr(torch.tril(buf[:n]).cpu().numpy(), Lref.cpu().numpy()) < 1e-12
    Eref = torch.linalg.solve_triangular(Lref, E.t(), upper=False).t()
    assert relerr(buf[n:].cpu().numpy(), Eref.cpu().numpy()) < 1e-11
    eye = torch.eye(128, dtype=torch.float64, device=cuda_device)
    for k in (0, 7, 15):
        W = invd[k]
        assert not torch.isnan(W).any() and float(torch.triu(W, 1).abs().max()) == 0.0
        assert float((W @ Lref[128 * k:128 * (k + 1), 128 * k:128 * (k + 1)] - eye).abs().max()) < 1e-11


@pytest.mark.parametrize("col", [0, 37, 200, 201, 255])
def test_first_failing_pivot_is_reported(cuda_device, col):
    """info = 1-based index of the first non-positive pivot, as LAPACK's potrf reports it -- for even and odd columns of
    the 2 x 2 pivot blocks, and across 32-column blocks and 128-column tiles; everything from that column on is NaN."""
    import torch

    from stopro_b200 import _lib

    lib = _lib.lib()
    n = 256
    g = torch.Generator(device=cuda_device).manual_seed(col)
    X = torch.randn(n, n, dtype=torch.float64, device=cuda_device, generator=g)
    S = X @ X.t() / n + torch.eye(n, dtype=torch.float64, device=cuda_device)
    Lref = torch.linalg.cholesky(S)
    # make the Schur complement at `col` negative: lower the diagonal entry below the sum of squares of its row of L
    S[col, col] = (Lref[col, :col] ** 2).sum() - 0.5
    buf = S.clone()
    invd = torch.empty(n // 128, 128, 128, dtype=torch.float64, device=cuda_device)
    info = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    _lib.check(lib.pigp_potrf_lower(buf.data_ptr(), n, n, 0, invd.data_ptr(), info.data_ptr(), None))
    torch.cuda.synchronize()
    assert int(info.item()) == col + 1
    assert torch.isnan(buf[col, col])
    if col > 0:
        assert relerr(torch.tril(buf[:col, :col]).cpu().numpy(), Lref[:col, :col].cpu().numpy()) < 1e-12
