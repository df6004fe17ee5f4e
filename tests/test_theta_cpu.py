"""CPU tests of the theta plumbing around the path: get_init (sub_modules/init_modules.py:5-54) and the handling of the
unused cross-covariance groups of params_main.yaml by the model classes (no GPU: the solver is replaced by a stub)."""
import numpy as np
import pytest
import yaml

from stopro_b200 import synthetic
from stopro_b200.sub_modules.init_modules import get_init
from stopro_b200.sub_modules.load_modules import load_data, load_params

YAML_6 = """
model:
  init_kernel_hyperparameter:
    uyuy: [1.1, -1.2, -1.3]
    uxux: [1.0, -1.0, -1.0]
    pp: [1.2, -1.4, -1.5]
    uxuy: [0.0, -1.2, -1.2]
    uxp: [0.1, -1.2, -1.2]
    uyp: [0.2, -1.2, -1.2]
  kernel_form: product
  kernel_type: se
  distance_func: false
  epsilon: 1.0e-06
  input_dim: 2
  system_type: Stokes_2D
  index_optimize_noise: null
optimization: {eps: 0.0001, loss_ridge_regression: false, lr: 1.0e-05, maxiter_GD: 0, maxiter_scipy: [500],
               method_GD: adam, method_scipy: [Nelder-Mead], print_process: true, interval_check: null, index_fixed: null}
"""


def test_get_init_layouts():
    pm = yaml.safe_load(YAML_6)["model"]
    th = get_init(pm["init_kernel_hyperparameter"], pm["kernel_type"], system_type=pm["system_type"])
    assert th.tolist() == [1.0, -1.0, -1.0, 1.1, -1.2, -1.3, 1.2, -1.4, -1.5, 0.0, -1.2, -1.2, 0.1, -1.2, -1.2, 0.2, -1.2, -1.2]
    three = {k: [0.0, -1.0, -1.0] for k in ("pp", "uxux", "uyuy")}
    assert get_init(dict(three), "se").tolist() == [0.0, -1.0, -1.0] * 3
    hp = dict(three, noise=[-3.0])
    th = get_init(hp, "se")
    assert th.tolist() == [0.0, -1.0, -1.0] * 3 + [-3.0] and "noise" not in hp       # the key is consumed, noise goes last
    d3 = {k: [0.0, -1.0, -1.0, -1.0] for k in ("uxux", "uyuy", "uzuz", "pp")}
    assert get_init(d3, "se", system_type="Stokes_3D").shape == (16,)
    assert get_init([0.0, 0.5], "se", system_type="1D").tolist() == [0.0, 0.5]
    with pytest.raises(NotImplementedError):
        get_init({}, "sm")


def test_load_params_and_load_data(tmp_path):
    (tmp_path / "params_main.yaml").write_text(YAML_6)
    (tmp_path / "params_prepare.yaml").write_text("system_name: sinusoidal\n")
    (tmp_path / "lbls.yaml").write_text("train: [ux, uy]\ntest: [ux]\n")
    params_main, params_prepare, lbls = load_params(str(tmp_path))
    assert params_main["model"]["epsilon"] == 1e-6 and params_prepare["system_name"] == "sinusoidal"

    class FakeHdf:
        def load_train_data(self, lb, vn):
            return [np.zeros((3, 2)), np.ones((2, 2))], [np.arange(3.0), np.arange(2.0)]

        def load_test_data(self, lb, vn):
            return [np.zeros((4, 2))], [np.arange(4.0)]

    r_test, mu_test, r_train, mu_train, f_train = load_data(lbls, {"train": "tr", "test": "te"}, FakeHdf())
    assert [m.tolist() for m in mu_train] == [[0.0] * 3, [0.0] * 2] and mu_test[0].shape == (4,)
    assert len(r_train) == 2 and len(r_test) == 1 and f_train[1].tolist() == [0.0, 1.0]


class StubSolver:
    def __init__(self, theta_len):
        class P:
            pass
        self.plan = P()
        self.plan.theta_len = theta_len
        self.calls = []

    def nll_grad_host(self, th, y, eps, want_grad=True):
        self.calls.append(np.array(th))
        return float(np.sum(th)), (np.arange(1.0, th.size + 1.0) if want_grad else None), 0


@pytest.mark.parametrize("noise", [False, True])
def test_unused_theta_groups_are_dropped_and_get_zero_gradient(noise):
    cfg = synthetic.sinusoidal(u_num=6, f_nx=5, f_ny=3, dif_num=4, n_test=4)
    cfg["model_kwargs"] = dict(cfg["model_kwargs"], index_optimize_noise=[4, 5] if noise else None)
    gp = synthetic.make_model(cfg)
    n_plan = 9 + int(noise)
    stub = StubSolver(n_plan)
    gp._solver_for = lambda r: stub
    th9 = np.linspace(-1.0, 1.0, n_plan)
    # exact length: passed through untouched
    nll, g = gp.value_and_grad(th9, cfg["r_train"], cfg["delta_y"], 1e-6)
    assert stub.calls[-1].tolist() == th9.tolist() and g.shape == (n_plan,)
    # 6 groups from the YAML (+ noise last): the three unused groups are dropped, their gradient is zero
    th18 = np.concatenate([th9[:9], 7.0 + np.arange(9.0), th9[9:]])
    nll18, g18 = gp.value_and_grad(th18, cfg["r_train"], cfg["delta_y"], 1e-6)
    assert nll18 == nll and g18.shape == th18.shape
    assert g18[:9].tolist() == g[:9].tolist() and not g18[9:18].any()
    if noise:
        assert g18[-1] == g[-1]
    assert gp.d_logposterior(th18, cfg["r_train"], cfg["delta_y"], 1e-6).tolist() == (g18 + 1.0).tolist()
    with pytest.raises(ValueError):
        gp.value_and_grad(th9[:5], cfg["r_train"], cfg["delta_y"], 1e-6)
