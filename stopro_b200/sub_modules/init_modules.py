"""get_init with the reference's signature (sub_modules/init_modules.py:5-54): the flat log-space theta vector from the
``init_kernel_hyperparameter`` entry of params_main.yaml.

Layout (the contract the CUDA evaluator honours):
  Stokes_2D: [uxux(3), uyuy(3), pp(3)] (+ uxuy(3) (+ uxp(3), uyp(3)) when the YAML carries 4 / 6 groups -- the
             *independent* model classes ignore those, exactly like the reference's theta slices ind_uxux / ind_uyuy /
             ind_pp), noise appended last;
  Stokes_3D: [uxux(4), uyuy(4), uzuz(4), pp(4)], noise last;
  otherwise: the flat YAML list, noise last.
Keys are read by name, so their order in the YAML file is irrelevant.  Like the reference, the "noise" key is removed
from the dictionary that is passed in.
"""
import numpy as np


def get_init(hyperparams, kernel_type, use_gradp_training=False, system_type="Stokes_2D"):
    noise = None
    if isinstance(hyperparams, dict) and "noise" in hyperparams:
        noise = np.asarray(hyperparams["noise"], dtype=np.float64)
        del hyperparams["noise"]
    if kernel_type == "sm":
        raise NotImplementedError("the spectral-mixture kernel is not on the B200 path (only kernel_type == 'se')")
    group = lambda k: np.atleast_1d(np.asarray(hyperparams[k], dtype=np.float64))
    if system_type == "Stokes_3D":
        init = np.concatenate([group(k) for k in ("uxux", "uyuy", "uzuz", "pp")])
    elif system_type == "Stokes_2D":
        if use_gradp_training:
            raise ValueError("Not implemented yet")
        keys = {3: ("uxux", "uyuy", "pp"), 4: ("uxux", "uyuy", "pp", "uxuy"),
                6: ("uxux", "uyuy", "pp", "uxuy", "uxp", "uyp")}.get(len(hyperparams))
        if keys is None:
            raise ValueError(f"Stokes_2D expects 3, 4 or 6 hyper-parameter groups, got {len(hyperparams)}")
        init = np.concatenate([group(k) for k in keys])
    else:
        init = np.atleast_1d(np.asarray(hyperparams, dtype=np.float64))
    if noise is not None and noise.size and np.any(noise):  # the reference tests ``if noise:`` (init_modules.py:52)
        init = np.append(init, noise)
    return init
