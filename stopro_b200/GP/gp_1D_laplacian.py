"""GPmodel1DLaplacian: observations of y and of y'' (reference: GP/gp_1D_laplacian.py:7-59)."""
from .gp import GPmodel


class GPmodel1DLaplacian(GPmodel):
    system = "scalar"
    train_observables = ("y", "ly")
    test_observables = ("y",)

    def __init__(self, Kernel=None):
        super().__init__(Kernel=Kernel)
