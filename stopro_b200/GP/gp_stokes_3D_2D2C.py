"""GPStokes2D2C / GPStokes2D2CSurface (reference: GP/gp_stokes_3D_2D2C.py:9-188): 3-D Stokes model trained on
two-component velocity data (2D2C: planar PIV-like measurements), inferring all three components.

  GPStokes2D2C         training [ux, uy, fx, fy, fz, div]                    (:14-37)
  GPStokes2D2CSurface  training [ux, uy, ux, uy, uz, fx, fy, fz, div]        (:86-136: in-plane data + surface data)
  both                 inference of [ux, uy, uz]                             (:39-78, :138-188)
"""
from .gp_stokes_3D import GPStokes3D


class GPStokes2D2C(GPStokes3D):
    train_observables = ("ux", "uy", "fx", "fy", "fz", "div")

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.infer_difp or self.use_difp:
            raise NotImplementedError("GPStokes2D2C has no tables for use_difp / infer_difp (gp_stokes_3D_2D2C.py:42-45)")


class GPStokes2D2CSurface(GPStokes3D):
    train_observables = ("ux", "uy", "ux", "uy", "uz", "fx", "fy", "fz", "div")

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.infer_difp or self.use_difp:
            raise NotImplementedError("GPStokes2D2CSurface has no tables for use_difp / infer_difp (gp_stokes_3D_2D2C.py:141-144)")
