"""GPSinusoidalWithoutPIndependent (reference: GP/gp_sinusoidal_independent.py:9-181).

Training blocks [ux, uy, difux, difuy, fx, fy, div, difp] where dif* are periodic differences over ``lbox``
(the pressure is only seen through its inlet/outlet difference).  Inference of [ux, uy], or of the governing
equations [fx, fy, div] with ``infer_governing_eqs`` (:92-124, :171-176).  Without ``use_difp`` the caller simply
passes seven training arrays and the leading 7 x 7 part of the table is used, as in the reference (:148-168).
"""
from .gp import GPmodel


class GPSinusoidalWithoutPIndependent(GPmodel):
    train_observables = ("ux", "uy", "difux", "difuy", "fx", "fy", "div", "difp")

    def __init__(self, lbox=None, use_difp=False, use_difu=False, infer_governing_eqs=False, Kernel=None,
                 index_optimize_noise=None):
        self.use_difp = use_difp
        self.use_difu = use_difu
        self.infer_governing_eqs = infer_governing_eqs
        self.test_observables = ("fx", "fy", "div") if infer_governing_eqs else ("ux", "uy")
        super().__init__(Kernel=Kernel, index_optimize_noise=index_optimize_noise, lbox=lbox)
