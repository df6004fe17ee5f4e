"""GPPoiseuilleIndependent (reference: GP/gp_poiseuille_independent.py:7-42).

Training blocks [ux, uy, p, fx, fy, div], inference of [ux, uy, p]; the 6 x 6 / 3 x 6 / 3 x 3 block tables of the
reference follow from those observables (stopro_b200.operators).
"""
from .gp import GPmodel


class GPPoiseuilleIndependent(GPmodel):
    train_observables = ("ux", "uy", "p", "fx", "fy", "div")
    test_observables = ("ux", "uy", "p")

    def __init__(self, Kernel=None):
        super().__init__(Kernel=Kernel)
