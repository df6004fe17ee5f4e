"""GPU tests of the block-cyclic multi-rank evaluation (csrc/pigp_dist.cu).

Several ranks run inside ONE process on ONE device (one stream and one host thread per rank, slabs connected by raw
pointers), so the complete protocol -- peer stores from the kernels, epoch flags, partial-gradient exchange -- is
exercised on the single-GPU box; the multi-process CUDA-IPC wiring is exercised by bench.py under torchrun.
Checked against the CPU oracle and against the single-GPU solver.
"""
import threading

import numpy as np
import pytest

from conftest import oracle_for
from stopro_b200 import synthetic
from stopro_b200.dist import DistSolver

pytestmark = pytest.mark.gpu
F_TOL = 1e-8


def run_ranks(gp, cfg, theta, world, repeats=1):
    r, y, eps = cfg["r_train"], cfg["delta_y"], cfg["eps"]
    gp.set_constants(r, y, eps, only_training=True)
    plan = gp._training_plan(r)
    solvers = [DistSolver(plan, k, world) for k in range(world)]
    slabs = [s.slab()[0] for s in solvers]
    for s in solvers:
        s.connect_pointers(slabs)
    out = [None] * world

    def work(k):
        try:
            for _ in range(repeats):
                out[k] = solvers[k].nll_grad_host(theta, y, eps)
        except Exception as exc:  # noqa: BLE001
            out[k] = exc

    th = [threading.Thread(target=work, args=(k,)) for k in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for s in solvers:
        s.close()
    for o in out:
        if isinstance(o, Exception):
            raise o
    return out


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


CASES = {
    "poiseuille_product": lambda: dict(synthetic.poiseuille(kernel_form="product"), eps=1e-2),        # N = 498, 4 tiles
    "poiseuille_additive": lambda: dict(synthetic.poiseuille(kernel_form="additive"), eps=1e-2),
    "sinusoidal": lambda: dict(synthetic.sinusoidal(u_num=16, f_nx=14, f_ny=8, dif_num=9, n_test=10), eps=1e-2),
    "sin1d_naive": lambda: synthetic.sin_1d_naive(),                                                   # N = 32: one tile, noise theta
    # eps = 1 keeps cond(K) * u below 1e-8 (the 4th-derivative blocks have entries ~ 1e5), so the strict bar applies
    "scaling_1280": lambda: dict(synthetic.stokes2d_scaling(1280, n_test=8), eps=1.0),                 # N multiple of 128
    "scaling_1500": lambda: dict(synthetic.stokes2d_scaling(1500, n_test=8), eps=1.0),
}


@pytest.mark.parametrize("world", [1, 2, 3, 4])
@pytest.mark.parametrize("name", list(CASES))
def test_sharded_matches_oracle(cuda_device, name, world):
    cfg = CASES[name]()
    gp = synthetic.make_model(cfg)
    ref = oracle_for(cfg)
    rng = np.random.default_rng(3)
    th = cfg["theta0"].copy()
    nk = len(th) - (1 if cfg["model_kwargs"].get("index_optimize_noise") else 0)
    th[:nk] += 0.1 * rng.standard_normal(nk)
    args = (cfg["r_train"], cfg["delta_y"], cfg["eps"])
    nll_ref = ref.trainingFunction_all(th, *args)
    g_ref = ref.d_trainingFunction_all(th, *args)
    out = run_ranks(gp, cfg, th, world, repeats=2)
    for nll, grad, info in out:
        assert info == 0
        assert abs(nll - nll_ref) <= F_TOL * abs(nll_ref)
        assert relerr(grad, g_ref) <= F_TOL
    # every rank returns the same bits
    for nll, grad, _ in out[1:]:
        assert nll == out[0][0]
        assert np.array_equal(grad, out[0][1])
    gp.close()


def test_sharded_matches_single_solver(cuda_device):
    cfg = dict(synthetic.stokes2d_scaling(3000, n_test=8), eps=1.0)
    gp = synthetic.make_model(cfg)
    args = (cfg["r_train"], cfg["delta_y"], cfg["eps"])
    nll1, g1 = gp.value_and_grad(cfg["theta0"], *args)
    for world in (2, 4):  # in-process ranks use 4 streams each and CUDA offers at most 32 hardware queues per process
        out = run_ranks(gp, cfg, cfg["theta0"], world)
        assert abs(out[0][0] - nll1) <= 1e-10 * abs(nll1)
        assert relerr(out[0][1], g1) <= 1e-8
    gp.close()


def test_not_positive_definite_gives_nan(cuda_device):
    cfg = dict(synthetic.poiseuille(kernel_form="product"), eps=-10.0)  # K - 10 I is indefinite
    gp = synthetic.make_model(cfg)
    out = run_ranks(gp, cfg, cfg["theta0"], 2)
    for nll, grad, info in out:
        assert info != 0 and np.isnan(nll) and np.all(np.isnan(grad))
    gp.close()


@pytest.mark.parametrize("world,width", [(1, 2), (1, 4), (2, 2), (3, 3), (4, 2)])
def test_lookahead_panel_schedule_matches_recursion(cuda_device, world, width):
    """The panel schedule with look-ahead (pigp_set_lookahead; bulk trailing updates off the dependency chain) computes the
    same factorisation, inverse and gradient as the plain recursion."""
    from stopro_b200 import _lib

    cfg = dict(synthetic.stokes2d_scaling(1500, n_test=8), eps=1.0)   # 12 tiles: several panels for every width
    gp = synthetic.make_model(cfg)
    try:
        _lib.check(_lib.lib().pigp_set_lookahead(0))
        base = run_ranks(gp, cfg, cfg["theta0"], world)
        _lib.check(_lib.lib().pigp_set_lookahead(width))
        out = run_ranks(gp, cfg, cfg["theta0"], world, repeats=2)
    finally:
        _lib.check(_lib.lib().pigp_set_lookahead(-1))
    for (nll, grad, info), (nll0, grad0, _) in zip(out, base):
        assert info == 0
        assert abs(nll - nll0) <= 1e-11 * abs(nll0)
        assert relerr(grad, grad0) <= 1e-9
    gp.close()


def test_automatic_panel_schedule_for_nll_only(cuda_device):
    """Default (automatic) width: a single-GPU NLL-only evaluation of >= 12 tiles takes the panel schedule; the value must
    agree with the plain recursion (width 0) and with the NLL half of the NLL+gradient call."""
    import torch

    from stopro_b200 import _lib

    cfg = dict(synthetic.stokes2d_scaling(2000, n_test=8), eps=1.0)   # 16 tiles -> automatic width 2
    gp = synthetic.make_model(cfg)
    gp.set_constants(cfg["r_train"], cfg["delta_y"], cfg["eps"], only_training=True)
    solver = gp._solver_for(cfg["r_train"])
    dev = torch.device("cuda:0")
    theta = torch.as_tensor(cfg["theta0"], device=dev)
    y = torch.as_tensor(cfg["delta_y"], device=dev)
    out = torch.zeros(1 + solver.plan.theta_len, dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    vals = {}
    try:
        for w in (0, -1):
            _lib.check(_lib.lib().pigp_set_lookahead(w))
            for _ in range(2):
                solver.nll(theta.data_ptr(), y.data_ptr(), cfg["eps"], out.data_ptr(), info.data_ptr(), None)
            torch.cuda.synchronize()
            assert int(info.item()) == 0
            vals[w] = float(out[0].item())
    finally:
        _lib.check(_lib.lib().pigp_set_lookahead(-1))
    solver.nll_grad(theta.data_ptr(), y.data_ptr(), cfg["eps"], out.data_ptr(), out.data_ptr() + 8, info.data_ptr(), None)
    torch.cuda.synchronize()
    both = float(out[0].item())
    assert abs(vals[-1] - vals[0]) <= 1e-11 * abs(vals[0])
    assert abs(vals[-1] - both) <= 1e-11 * abs(both)
    gp.close()


def test_lost_peer_times_out_and_reset_recovers(cuda_device):
    """a rank that never shows up must not hang a GPU.  Rank 0 evaluates alone: its opening barrier runs into
    the (test-shortened) time-out, the host call reports it, results are NaN; after pigp_dsolver_reset on both ranks the
    pair evaluates correctly."""
    import time

    from stopro_b200 import _lib

    cfg = dict(synthetic.poiseuille(kernel_form="product"), eps=1e-2)
    gp = synthetic.make_model(cfg)
    ref = oracle_for(cfg)
    r, y, eps = cfg["r_train"], cfg["delta_y"], cfg["eps"]
    gp.set_constants(r, y, eps, only_training=True)
    plan = gp._training_plan(r)
    solvers = [DistSolver(plan, k, 2) for k in range(2)]
    slabs = [s.slab()[0] for s in solvers]
    for s in solvers:
        s.connect_pointers(slabs)
    t0 = time.perf_counter()
    with pytest.raises(_lib.PigpError, match="timed out"):
        solvers[0].nll_grad_host(cfg["theta0"], y, eps)
    assert time.perf_counter() - t0 < 90.0
    for s in solvers:
        s.reset()
    out = [None, None]

    def work(k):
        out[k] = solvers[k].nll_grad_host(cfg["theta0"], y, eps)

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    nll_ref = ref.trainingFunction_all(cfg["theta0"], r, y, eps)
    for nll, grad, info in out:
        assert info == 0 and abs(nll - nll_ref) <= F_TOL * abs(nll_ref)
    for s in solvers:
        s.close()
    gp.close()
