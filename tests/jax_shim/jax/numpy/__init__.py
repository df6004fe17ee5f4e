"""Minimal `jax.numpy` stand-in on torch (float64) -- test infrastructure, see tests/jax_shim/README.md.

Arrays are torch.Tensor.  Importing this module adds two JAX-isms to torch.Tensor (both are no-ops for code that does not
use them): the functional-update property `x.at[idx].set(v)` and `x.transpose()` without arguments (GP/gp.py:109).
"""
import builtins as _b
import math

import numpy as _np
import torch

from . import linalg  # noqa: F401

pi = math.pi
inf = math.inf
nan = math.nan
e = math.e
newaxis = None
ndarray = torch.Tensor
float64 = torch.float64
float32 = torch.float32
int64 = torch.int64
int32 = torch.int32
bool_ = torch.bool

_F64 = torch.float64


def _asarray(x, dtype=None):
    """Anything array-like -> tensor; floating data becomes float64 (jax_enable_x64)."""
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, (list, tuple)) and len(x) > 0 and _b.any(isinstance(v, torch.Tensor) for v in x):
        t = torch.stack([_asarray(v) for v in x])
        return t if dtype is None else t.to(dtype)
    a = _np.asarray(x)
    if a.dtype.kind == "f":
        a = a.astype(_np.float64)
    t = torch.from_numpy(_np.ascontiguousarray(a)) if a.ndim else torch.tensor(a.item(), dtype=_F64 if a.dtype.kind == "f" else None)
    return t if dtype is None else t.to(dtype)


def _shape(shape):
    if isinstance(shape, (int, _np.integer)):
        return (int(shape),)
    return tuple(int(s) for s in shape)


# ------------------------------------------------------------------ creation
def array(x, dtype=None):
    t = _asarray(x, dtype)
    return t.clone() if isinstance(x, torch.Tensor) else t


asarray = _asarray


def zeros(shape, dtype=None):
    return torch.zeros(_shape(shape), dtype=dtype or _F64)


def ones(shape, dtype=None):
    return torch.ones(_shape(shape), dtype=dtype or _F64)


def empty(shape, dtype=None):
    return torch.zeros(_shape(shape), dtype=dtype or _F64)


def zeros_like(x):
    return torch.zeros_like(_asarray(x))


def ones_like(x):
    return torch.ones_like(_asarray(x))


def eye(n, dtype=None):
    return torch.eye(int(n), dtype=dtype or _F64)


def arange(*args, **kw):
    t = torch.arange(*args)
    return t


def linspace(a, b, num=50, endpoint=True):
    return torch.from_numpy(_np.linspace(a, b, num, endpoint=endpoint))


# ------------------------------------------------------------------ elementwise
def _un(fn):
    def f(x):
        return fn(_asarray(x))

    return f


exp = _un(torch.exp)
log = _un(torch.log)
sqrt = _un(torch.sqrt)
abs = _un(torch.abs)
sin = _un(torch.sin)
cos = _un(torch.cos)
tanh = _un(torch.tanh)
arcsin = _un(torch.asin)
square = _un(torch.square)
round = _un(torch.round)
isnan = _un(torch.isnan)


def power(x, y):
    return torch.pow(_asarray(x), y)


def multiply(x, y):
    return _asarray(x) * y


def where(c, a, b):
    return torch.where(c, _asarray(a), _asarray(b))


# ------------------------------------------------------------------ reductions
def _axis_red(fn):
    def f(x, axis=None):
        x = _asarray(x)
        return fn(x) if axis is None else fn(x, dim=axis)

    return f


sum = _axis_red(torch.sum)
prod = _axis_red(torch.prod)
mean = _axis_red(torch.mean)


def max(x, axis=None):
    x = _asarray(x)
    return torch.max(x) if axis is None else torch.max(x, dim=axis).values


def min(x, axis=None):
    x = _asarray(x)
    return torch.min(x) if axis is None else torch.min(x, dim=axis).values


def argmax(x, axis=None):
    return torch.argmax(_asarray(x)) if axis is None else torch.argmax(_asarray(x), dim=axis)


def any(x):
    return torch.any(_asarray(x))


def all(x):
    return torch.all(_asarray(x))


# ------------------------------------------------------------------ shape / indexing
def transpose(x, axes=None):
    x = _asarray(x)
    if axes is None:
        axes = tuple(reversed(range(x.dim())))
    return x.permute(*axes)


def concatenate(xs, axis=0):
    return torch.cat([torch.atleast_1d(_asarray(x)) for x in xs], dim=axis)


def append(a, b):
    return torch.cat([_asarray(a).reshape(-1).to(_F64), _asarray(b).reshape(-1).to(_F64)])


def delete(x, idx, axis=None):
    x = _asarray(x)
    i = int(idx)
    return torch.cat([x[:i], x[i + 1:]])


def split(x, indices_or_sections, axis=0):
    x = _asarray(x)
    if isinstance(indices_or_sections, (int, _np.integer)):
        return list(torch.tensor_split(x, int(indices_or_sections), dim=axis))
    return list(torch.tensor_split(x, [int(i) for i in indices_or_sections], dim=axis))


def diag(x, k=0):
    return torch.diag(_asarray(x), k)


def diagonal(x, offset=0, axis1=0, axis2=1):
    return torch.diagonal(_asarray(x), offset, axis1, axis2)


def reshape(x, shape):
    return _asarray(x).reshape(shape)


def setdiff1d(a, b):
    return torch.from_numpy(_np.setdiff1d(_np.asarray(a), _np.asarray(b)))


# ------------------------------------------------------------------ products
def dot(a, b):
    return torch.matmul(_asarray(a), _asarray(b))


def matmul(a, b):
    return torch.matmul(_asarray(a), _asarray(b))


def einsum(spec, *ops):
    return torch.einsum(spec.replace(" ", ""), *[_asarray(o) for o in ops])


# ------------------------------------------------------------------ x.at[idx].set(v): functional scatter
class _AtIndex:
    def __init__(self, x, idx):
        self.x, self.idx = x, idx

    def _positions(self):
        x = self.x
        return torch.arange(x.numel()).reshape(x.shape)[self.idx]

    def _apply(self, v, combine):
        x = self.x
        pos = self._positions()
        v = _asarray(v)
        if v.dtype != x.dtype:
            v = v.to(x.dtype)
        flat = x.reshape(-1)
        if combine is not None:
            v = combine(flat[pos.reshape(-1)].reshape(pos.shape), v)
        v = torch.broadcast_to(v, pos.shape)
        return flat.scatter(0, pos.reshape(-1), v.reshape(-1)).reshape(x.shape)

    def set(self, v):
        return self._apply(v, None)

    def add(self, v):
        return self._apply(v, lambda old, new: old + new)

    def multiply(self, v):
        return self._apply(v, lambda old, new: old * new)


class _At:
    def __init__(self, x):
        self.x = x

    def __getitem__(self, idx):
        return _AtIndex(self.x, idx)


if not hasattr(torch.Tensor, "at"):
    torch.Tensor.at = property(lambda self: _At(self))

if not getattr(torch.Tensor.transpose, "_jax_shim", False):
    _orig_transpose = torch.Tensor.transpose

    def _transpose(self, *dims):
        if not dims:
            return self.permute(*reversed(range(self.dim())))
        return _orig_transpose(self, *dims)

    _transpose._jax_shim = True
    torch.Tensor.transpose = _transpose
