"""Minimal `jax` stand-in on torch.func (float64) -- test infrastructure, see tests/jax_shim/README.md.

Only what the reference's hot path uses (GP/*.py, sub_modules/loss_modules.py, solver/optimizers.py):
grad, value_and_grad, hessian, jacfwd, jacrev, vmap, jit, config.  Semantics follow JAX:
  * argnums may be an int or a sequence; a sequence gives a tuple of results,
  * hessian(f, argnums) = jacfwd(jacrev(f, argnums), argnums): for argnums=[0, 1] the result is a tuple of tuples
    with out[i][j][a, b] = d^2 f / d arg_i[a] d arg_j[b]   (GP/gp_2D.py:24-34 reads [0][1][i, j]),
  * vmap(f, in_axes) with None entries for unmapped arguments (GP/gp.py:19-21).
"""
import torch
import torch.func as _tf

from . import numpy  # noqa: F401  (also patches torch.Tensor with .at / no-arg .transpose())
from . import config as _config_mod
from .numpy import _asarray

config = _config_mod.config


def _argnums(argnums):
    if isinstance(argnums, int):
        return argnums
    return tuple(int(a) for a in argnums)


def jit(fun=None, **_kw):
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    if isinstance(in_axes, list):
        in_axes = tuple(in_axes)

    def tensor_out(*args):
        # JAX lets a mapped function return Python scalars (GP/gp_2D_stokes_independent.py:14 Kzero = lambda: 0.0)
        out = fun(*args)
        return _tree_map_scalars(out)

    return _tf.vmap(tensor_out, in_dims=in_axes, out_dims=out_axes)


def _tree_map_scalars(out):
    if isinstance(out, (float, int)):
        return torch.tensor(float(out), dtype=torch.float64)
    if isinstance(out, tuple):
        return tuple(_tree_map_scalars(o) for o in out)
    if isinstance(out, list):
        return [_tree_map_scalars(o) for o in out]
    return out


def _tensor_args(args, argnums):
    """Differentiated arguments must be float64 tensors."""
    nums = (argnums,) if isinstance(argnums, int) else argnums
    args = list(args)
    for k in nums:
        if not isinstance(args[k], torch.Tensor):
            args[k] = _asarray(args[k])
    return args


def grad(fun, argnums=0, has_aux=False):
    an = _argnums(argnums)
    g = _tf.grad(fun, argnums=an, has_aux=has_aux)
    return lambda *args: g(*_tensor_args(args, an))


def value_and_grad(fun, argnums=0, has_aux=False):
    an = _argnums(argnums)
    g = _tf.grad_and_value(fun, argnums=an, has_aux=has_aux)

    def inner(*args):
        gr, val = g(*_tensor_args(args, an))
        return val, gr

    return inner


def jacfwd(fun, argnums=0, has_aux=False):
    an = _argnums(argnums)
    j = _tf.jacfwd(fun, argnums=an, has_aux=has_aux)
    return lambda *args: j(*_tensor_args(args, an))


def jacrev(fun, argnums=0, has_aux=False):
    an = _argnums(argnums)
    j = _tf.jacrev(fun, argnums=an, has_aux=has_aux)
    return lambda *args: j(*_tensor_args(args, an))


def hessian(fun, argnums=0):
    an = _argnums(argnums)
    h = _tf.jacfwd(_tf.jacrev(fun, argnums=an), argnums=an)
    return lambda *args: h(*_tensor_args(args, an))
