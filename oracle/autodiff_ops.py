"""Differential operators on the scalar kernel, built by nested autodiff.

This is the reference's construction restated one-to-one with ``torch.func``
(JAX is not installed here): each operator is a nested grad / hessian of the
scalar kernel, wrapped by ``outermap`` = vmap(vmap(f)) into a dense block.

Follows /root/reference/GP/gp.py:19-21 (outermap), GP/gp_2D.py:16-86 (2-D
operators), GP/gp_3D.py:12-35 (3-D additions), GP/gp_1D_laplacian.py:35-46.
"""
import torch
from torch.func import grad, hessian, vmap


def outermap(f):
    # gp.py:19-21  vmap(vmap(f, (None, 0, None)), (0, None, None))
    return vmap(vmap(f, in_dims=(None, 0, None)), in_dims=(0, None, None))


def _trace(h):
    return torch.sum(torch.diagonal(h))


def scalar_operators(kernel, dim):
    """name -> scalar function (r, rp, theta) for the given base kernel."""
    ops = {"K": kernel}
    if dim == 1:
        # gp_1D_laplacian.py:35-46
        def lap(f, i):
            return grad(grad(f, i), i)

        ops["L0K"] = lap(kernel, 0)
        ops["L1K"] = lap(kernel, 1)
        ops["LLK"] = lap(ops["L0K"], 1)
        return ops

    # gp_2D.py:18-71
    def cross(i, j):
        # jax.hessian(K, [0, 1])(r, rp, th)[0][1][i, j]
        return lambda r, rp, th: hessian(kernel, argnums=(0, 1))(r, rp, th)[0][1][i, j]

    def d_first(arg, i):
        return lambda r, rp, th: grad(kernel, arg)(r, rp, th)[i]

    ops["L0"] = lambda r, rp, th: _trace(hessian(kernel, 0)(r, rp, th))
    ops["L1"] = lambda r, rp, th: _trace(hessian(kernel, 1)(r, rp, th))
    ops["LL"] = lambda r, rp, th: _trace(hessian(ops["L1"], 0)(r, rp, th))
    for i in range(dim):
        ops[f"d0{i}"] = d_first(0, i)  # d/dr_i      (_d00, _d01)
        ops[f"d1{i}"] = d_first(1, i)  # d/dr'_i     (_d10, _d11, _d12)
        ops[f"d{i}L"] = (lambda i: lambda r, rp, th: grad(ops["L1"], 0)(r, rp, th)[i])(i)
        ops[f"Ld{i}"] = (lambda i: lambda r, rp, th: _trace(hessian(ops[f"d1{i}"], 0)(r, rp, th)))(i)
        for j in range(dim):
            ops[f"d{i}d{j}"] = cross(i, j)
    return ops


class AutodiffOps:
    """Dense-block operators: ``ops.block(name)(r, rp, theta_group)`` -> (n, m) tensor."""

    def __init__(self, kernel, dim):
        self.dim = dim
        self._scalar = scalar_operators(kernel, dim)
        self._dense = {}

    def names(self):
        return list(self._scalar)

    def scalar(self, name):
        return self._scalar[name]

    def block(self, name):
        if name not in self._dense:
            self._dense[name] = outermap(self._scalar[name])
        return self._dense[name]
