"""GPStokes3D (reference: GP/gp_stokes_3D.py:9-172).

Training blocks [ux, uy, uz, fx, fy, fz, div] (no pressure observations: theta_pp enters through the f-f blocks);
inference of [ux, uy, uz], or of the periodic pressure difference with ``infer_difp`` (:112-123, :163-166).
"""
from .gp import GPmodel


class GPStokes3D(GPmodel):
    train_observables = ("ux", "uy", "uz", "fx", "fy", "fz", "div")

    def __init__(self, lbox=None, use_difp=False, use_difu=False, infer_governing_eqs=False, Kernel=None,
                 index_optimize_noise=None, infer_difp=False):
        if infer_governing_eqs:
            raise NotImplementedError("the reference leaves infer_governing_eqs unimplemented for GPStokes3D (gp_stokes_3D.py:85-111)")
        self.use_difp = use_difp
        self.use_difu = use_difu
        self.infer_difp = infer_difp
        self.infer_governing_eqs = infer_governing_eqs
        self.test_observables = ("difp",) if infer_difp else ("ux", "uy", "uz")
        super().__init__(Kernel=Kernel, index_optimize_noise=index_optimize_noise, lbox=lbox)
