"""CPU tests of the host-side bookkeeping of stopro_b200.GP.gp.GPmodel with the device classes replaced by stubs:
plan compilation in set_constants, plan reuse / point replacement / rebuild, the one-factorisation cache behind
func(theta) + dfunc(theta) (solver/optimizers.py:148-150), the +1.0 prior gradient, the per-variable split of the
posterior, and the reference's calling conventions (GP/gp.py:213-256, :263-285, :491-493)."""
import numpy as np
import pytest

from stopro_b200 import synthetic
from stopro_b200.GP import gp as gp_module


class FakePlan:
    created = []

    def __init__(self, dim, product_form, fields, row_obs, row_pts, col_obs=None, col_pts=None, lbox=None,
                 noise_blocks=None, zero_blocks=(), kernel_type="se"):
        self.symmetric = col_obs is None
        self.row_pts = [np.array(p) for p in row_pts]
        self.col_pts = None if col_pts is None else [np.array(p) for p in col_pts]
        self.rows = sum(len(p) for p in row_pts)
        self.cols = self.rows if col_pts is None else sum(len(p) for p in col_pts)
        self.theta_len = len(fields) * (1 + dim) + (1 if noise_blocks else 0)
        self.set_calls, self.closed = [], False
        self.zero_blocks = zero_blocks
        FakePlan.created.append(self)

    def same_points(self, row_pts, col_pts=None):
        eq = lambda a, b: len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
        return eq(self.row_pts, row_pts) and (col_pts is None or eq(self.col_pts, col_pts))

    def set_points(self, side, pts):
        self.set_calls.append(side)
        if side == 0:
            self.row_pts = [np.array(p) for p in pts]
        else:
            self.col_pts = [np.array(p) for p in pts]

    def assemble_host(self, theta, eps=0.0, add_diag=False, layout=0):
        return np.full((self.rows, self.cols), float(np.sum(theta)) + eps * add_diag)

    def close(self):
        self.closed = True


class FakeSolver:
    created = []

    def __init__(self, plan, rank=0, world=1):
        self.plan, self.evals, self.closed, self.rank, self.world = plan, 0, False, rank, world
        FakeSolver.created.append(self)

    def connect_ipc(self, group=None):
        pass

    def nll_grad_host(self, th, y, eps, want_grad=True, pts=None):
        self.evals += 1
        return float(np.sum(th) + np.sum(y)), (np.arange(len(th), dtype=float) if want_grad else None), 0

    def predict_batch_host(self, mixed, test, thetas, y, eps, full_cov=True):
        m, nb = mixed.rows, len(thetas)
        mu = np.tile(np.arange(m, dtype=float), (nb, 1)) + np.arange(nb)[:, None]
        cov = np.tile(np.eye(m), (nb, 1, 1)) if full_cov else np.ones((nb, m))
        return mu, cov, np.zeros(nb, dtype=np.int32)

    def close(self):
        self.closed = True


@pytest.fixture
def model(monkeypatch):
    FakePlan.created, FakeSolver.created = [], []
    monkeypatch.setattr(gp_module, "Plan", FakePlan)
    monkeypatch.setattr(gp_module, "Solver", FakeSolver)
    cfg = synthetic.poiseuille(u_num=4, p_num=4, f_num=3, n_test=3, kernel_form="product")
    return cfg, synthetic.make_model(cfg)


def test_set_constants_compiles_three_plans_and_sections(model):
    cfg, gp = model
    gp.set_constants(cfg["r_test"], cfg["mu_test"], cfg["r_train"], cfg["delta_y"], cfg["eps"])
    assert len(FakePlan.created) == 3
    assert gp.num_tr == 6 and gp.num_te == 3
    assert gp.sec_tr.tolist() == np.concatenate([[0], np.cumsum([len(r) for r in cfg["r_train"]])]).tolist()
    train, mixed, test = FakePlan.created
    assert train.symmetric and test.symmetric and not mixed.symmetric
    assert (mixed.rows, mixed.cols) == (test.rows, train.rows)
    gp.set_constants(cfg["r_train"], cfg["delta_y"], cfg["eps"], only_training=True)   # three-argument form
    assert len(FakePlan.created) == 3                                                    # same points: nothing rebuilt


def test_func_then_dfunc_costs_one_factorisation(model):
    cfg, gp = model
    args = (cfg["r_train"], cfg["delta_y"], cfg["eps"])
    th = cfg["theta0"] + 0.1
    f = gp.trainingFunction_all(th, *args)               # func(theta): NLL only
    g = gp.d_trainingFunction_all(th, *args)             # dfunc(theta): needs the gradient -> second evaluation
    solver = FakeSolver.created[-1]
    assert solver.evals == 2
    nll, g2 = gp.value_and_grad(th, *args)               # cached now
    assert solver.evals == 2 and nll == f and np.array_equal(g, g2)
    assert np.array_equal(gp.d_logposterior(th, *args), g + 1.0) and solver.evals == 2
    f2 = gp.trainingFunction_all(th + 2e-3, *args)       # a gradient has been asked for before: value AND gradient now,
    g3 = gp.d_logposterior(th + 2e-3, *args)             # so the pair func(theta), dfunc(theta) costs ONE evaluation
    assert solver.evals == 3 and np.isfinite(f2) and g3.shape == th.shape
    gp.value_and_grad(th + 1e-3, *args)                  # new theta -> new evaluation
    solver.evals -= 1
    assert solver.evals == 3
    gp.value_and_grad(th + 1e-3, args[0], args[1] * 2.0, args[2])   # new delta_y -> new evaluation
    assert solver.evals == 4
    g[0] = 123.0                                         # callers may modify what they get: the cache keeps its own copy
    assert gp.d_trainingFunction_all(th + 1e-3, args[0], args[1] * 2.0, args[2])[0] != 123.0


def test_new_points_replace_coordinates_new_sizes_rebuild(model):
    cfg, gp = model
    args = (cfg["r_train"], cfg["delta_y"], cfg["eps"])
    gp.set_constants(*args, only_training=True)
    gp.value_and_grad(cfg["theta0"], *args)
    plan, solver = FakePlan.created[-1], FakeSolver.created[-1]
    moved = [r + 0.01 for r in cfg["r_train"]]
    gp.value_and_grad(cfg["theta0"], moved, cfg["delta_y"], cfg["eps"])
    assert plan.set_calls == [0] and len(FakePlan.created) == 1 and solver.evals == 2   # same plan, points replaced, cache dropped
    smaller = [r[:-1] for r in cfg["r_train"]]
    gp.value_and_grad(cfg["theta0"], smaller, cfg["delta_y"][:-6], cfg["eps"])
    assert len(FakePlan.created) == 2 and plan.closed and solver.closed                 # other block sizes: rebuilt
    with pytest.raises(ValueError):
        gp._training_plan(cfg["r_train"][:5] + [cfg["r_train"][0]] * 2)                 # more blocks than the class has


def test_prediction_is_split_per_variable_and_shifted_by_the_prior_mean(model):
    cfg, gp = model
    mu_test = [np.full(len(r), 10.0 * (i + 1)) for i, r in enumerate(cfg["r_test"])]
    args = (cfg["r_test"], mu_test, cfg["r_train"], cfg["delta_y"], cfg["eps"])
    gp.set_constants(*args)
    mus, covs = gp.predictingFunction_all(cfg["theta0"], *args)
    m = len(cfg["r_test"][0])
    assert [len(x) for x in mus] == [m, m, m] and [c.shape for c in covs] == [(m, m)] * 3
    assert mus[1][0] == m + 20.0                           # second block starts at flat index m, plus its prior mean
    _, var = gp.predictingFunction_all(cfg["theta0"], *args, full_cov=False)
    assert [v.shape for v in var] == [(m,)] * 3
    K = gp.trainingK_all(cfg["theta0"], cfg["r_train"])
    assert K.shape == (gp.sec_tr[-1], gp.sec_tr[-1])


def test_kernel_is_required_and_close_releases_everything(model):
    cfg, gp = model
    gp.set_constants(cfg["r_test"], cfg["mu_test"], cfg["r_train"], cfg["delta_y"], cfg["eps"])
    gp.value_and_grad(cfg["theta0"], cfg["r_train"], cfg["delta_y"], cfg["eps"])
    gp.close()
    assert all(p.closed for p in FakePlan.created) and all(s.closed for s in FakeSolver.created)
    with pytest.raises(ValueError):
        type(gp)(Kernel=None)


def test_predict_many_is_one_call_per_list_and_splits_per_block(model):
    cfg, gp = model
    args = (cfg["r_test"], cfg["mu_test"], cfg["r_train"], cfg["delta_y"], cfg["eps"])
    gp.set_constants(*args)
    thetas = [cfg["theta0"], cfg["theta0"] + 0.1, cfg["theta0"] - 0.2]
    mus, vars_ = gp.predict_many(thetas, *args)
    assert len(mus) == 3 and len(vars_) == 3
    sizes = [len(r) for r in cfg["r_test"]]
    for b in range(3):
        assert [len(m) for m in mus[b]] == sizes and [len(v) for v in vars_[b]] == sizes
        flat = np.concatenate(mus[b])
        assert np.allclose(flat, np.arange(sum(sizes)) + b)           # per-theta result, blocks in order, mu_test = 0 added


def test_shard_points_partitions_every_block():
    pts = [np.arange(10.0).reshape(5, 2), np.arange(14.0).reshape(7, 2), np.zeros((0, 2))]
    seen = [np.zeros(len(p), dtype=int) for p in pts]
    for rank in range(3):
        part, slices = gp_module.GPmodel.shard_points(pts, rank, 3)
        for i, ((lo, hi), q) in enumerate(zip(slices, part)):
            assert len(q) == hi - lo and np.array_equal(q, pts[i][lo:hi])
            seen[i][lo:hi] += 1
    assert all(np.all(s == 1) for s in seen)
