"""Registration of libpigp's entry points as JAX FFI custom calls (BASELINE.json's north star), for installations that
have JAX >= 0.4.31.  This image has no JAX: importing this module without it raises ImportError, nothing else in
stopro_b200 depends on it, and the numpy + ctypes classes in stopro_b200.GP are the tested host side here.

    import jax_ffi                       # this file, integration/xla_ffi/jax_ffi.py
    f = jax_ffi.make_training_function(gp_model, r_train, eps)     # replaces gp_model.trainingFunction_all
    loss = jax.jit(lambda th, dy: f(th, dy) + jnp.sum(th))          # logposterior (sub_modules/loss_modules.py:5-13)
    dloss = jax.jit(jax.grad(loss, 0))                              # as test/test_1_sinusoidal_direct_main.py:76-78

The C++ handlers are pigp_xla_ffi.cc next to this file (`make JAX_FFI_INCLUDE=...` here).
"""
import ctypes
import os

import jax
import jax.numpy as jnp

_HERE = os.path.dirname(os.path.abspath(__file__))
_TARGETS = ("PigpAssemble", "PigpNll", "PigpNllGrad", "PigpPredict")
_registered = False


def register(path=None):
    """dlopen libpigp_xla_ffi.so and register its handlers for the CUDA platform (idempotent)."""
    global _registered
    if _registered:
        return
    lib = ctypes.CDLL(path or os.path.join(_HERE, "libpigp_xla_ffi.so"))
    for name in _TARGETS:
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform="CUDA")
    _registered = True


def make_training_function(gp_model, r_train, eps):
    """NLL(theta, delta_y) with a custom VJP whose backward pass is the fused dK/dtheta trace gradient
    (GP/gp.py:213-224 and :412-488 from ONE factorisation).  `gp_model` is a stopro_b200.GP model after set_constants."""
    register()
    solver = gp_model._solver_for(r_train)
    addr, P = int(solver.handle.value), solver.plan.theta_len
    out_types = (jax.ShapeDtypeStruct((1,), jnp.float64), jax.ShapeDtypeStruct((P,), jnp.float64),
                 jax.ShapeDtypeStruct((1,), jnp.int32))
    call = jax.ffi.ffi_call("PigpNllGrad", out_types)

    @jax.custom_vjp
    def nll(theta, delta_y):
        return call(theta, delta_y, solver=addr, eps=float(eps))[0][0]

    def fwd(theta, delta_y):
        value, grad, _info = call(theta, delta_y, solver=addr, eps=float(eps))
        return value[0], grad

    def bwd(grad, ct):
        return ct * grad, None  # the reference differentiates with respect to theta only: grad(func, 0)

    nll.defvjp(fwd, bwd)
    return nll
