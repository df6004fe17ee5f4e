"""GPStokes3DNaive (reference: GP/gp_stokes_3D_naive.py:9-128): velocity-only 3-D model.

Training blocks [ux, uy, uz], inference of [ux, uy, uz]; no governing-equation observations.  The reference leaves
``infer_governing_eqs`` and ``use_difp`` without tables (:64-96 ``pass``), so they are rejected here.
"""
from .gp import GPmodel


class GPStokes3DNaive(GPmodel):
    train_observables = ("ux", "uy", "uz")
    test_observables = ("ux", "uy", "uz")

    def __init__(self, lbox=None, use_difp=False, use_difu=False, infer_governing_eqs=False, Kernel=None,
                 index_optimize_noise=None):
        if infer_governing_eqs or use_difp:
            raise NotImplementedError("GPStokes3DNaive has no tables for infer_governing_eqs / use_difp (gp_stokes_3D_naive.py:64-96)")
        self.use_difp, self.use_difu, self.infer_governing_eqs = use_difp, use_difu, infer_governing_eqs
        super().__init__(Kernel=Kernel, index_optimize_noise=index_optimize_noise, lbox=lbox)
