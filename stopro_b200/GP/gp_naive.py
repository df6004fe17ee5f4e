"""GPmodelNaive: plain single-block GP  [[Kyy]]  (reference: GP/gp_naive.py:4-45)."""
from .gp import GPmodel


class GPmodelNaive(GPmodel):
    system = "scalar"
    train_observables = ("y",)
    test_observables = ("y",)

    def __init__(self, Kernel=None, index_optimize_noise=None):
        super().__init__(Kernel=Kernel, index_optimize_noise=index_optimize_noise)
