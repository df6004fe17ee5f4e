"""jax.numpy.linalg calls of the reference's path: cholesky (GP/gp.py:83, 106, 430), solve (general LU solve, also on the
triangular factor: GP/gp.py:84, 109, 118, 432-433), norm."""
import torch


def cholesky(a):
    # jnp.linalg.cholesky returns NaNs for a non-positive-definite input instead of raising
    L, info = torch.linalg.cholesky_ex(a)
    if int(info.max()) != 0:
        return torch.full_like(a, float("nan"))
    return L


def solve(a, b):
    return torch.linalg.solve(a, b)


def norm(x, ord=None, axis=None):
    return torch.linalg.norm(x, ord=ord, dim=axis)


def inv(a):
    return torch.linalg.inv(a)


def slogdet(a):
    return torch.linalg.slogdet(a)
