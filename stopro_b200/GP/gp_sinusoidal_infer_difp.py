"""Variants of the sinusoidal-channel model trained WITHOUT the pressure-difference block
(reference: GP/gp_sinusoidal_infer_difp.py:7-101).  Training blocks [ux, uy, difux, difuy, fx, fy, div] (:44-59).

  GPSinusoidalInferDifP            infers the periodic pressure difference  p(r + lbox) - p(r)     (:60-68)
  GPSinusoidalInferUWithoutDifP    infers [ux, uy]                                                   (:71-82)
  GPSinusoidalInferGovWithoutDifP  infers [fx, fy, div]                                              (:85-101)

Reference quirk kept: the test table of GPSinusoidalInferGovWithoutDifP holds ``Kfxuy`` (zero) in the (fx, fy) slot
(:97) where ``Kfxfy`` would be expected.  Only ``testK_all`` can see it -- the posterior is returned per variable.
"""
from .gp_sinusoidal_independent import GPSinusoidalWithoutPIndependent


class _WithoutDifP(GPSinusoidalWithoutPIndependent):
    train_observables = ("ux", "uy", "difux", "difuy", "fx", "fy", "div")


class GPSinusoidalInferDifP(_WithoutDifP):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.test_observables = ("difp",)


class GPSinusoidalInferUWithoutDifP(_WithoutDifP):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.test_observables = ("ux", "uy")


class GPSinusoidalInferGovWithoutDifP(_WithoutDifP):
    test_zero_blocks = frozenset({(0, 1)})

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.test_observables = ("fx", "fy", "div")
