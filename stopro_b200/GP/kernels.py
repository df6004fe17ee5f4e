"""Kernel factory with the reference's signature: ``define_kernel(params_model)`` (GP/kernels.py:331-427).

The reference returns a Python callable that JAX differentiates; here the object only *names* the kernel
(family, form, input dimension) -- the derivatives are closed forms evaluated on the GPU -- but it stays callable
on scalars / points (numpy) with the reference's semantics so user code that probes ``Kernel(r1, r2, theta)`` works.
Only the squared-exponential family used by the reference's configurations is on the hot path.
"""
import numpy as np


class SquaredExponential:
    """K_1d_SquareExp / K_2d_SquareExp_{Add,Pro} / K_3d_SquareExp_Pro (GP/kernels.py:36-77); theta in log space."""

    def __init__(self, input_dim, form):
        self.input_dim = int(input_dim)
        self.form = form  # "product" | "additive"

    @property
    def product_form(self):
        return self.form == "product"

    def __call__(self, r1, r2, theta):
        r1 = np.atleast_1d(np.asarray(r1, dtype=np.float64))
        r2 = np.atleast_1d(np.asarray(r2, dtype=np.float64))
        theta = np.asarray(theta, dtype=np.float64)
        e = np.exp(-0.5 * ((r1 - r2) * np.exp(-theta[1:1 + self.input_dim])) ** 2)
        return np.exp(theta[0]) * (np.prod(e) if self.product_form else np.sum(e))

    def __repr__(self):
        return f"SquaredExponential(input_dim={self.input_dim}, form={self.form!r})"


def define_kernel(params_model, lbox=None):
    """Same keys as the reference: kernel_type, kernel_form, distance_func, input_dim."""
    kernel_type = params_model["kernel_type"]
    kernel_form = params_model["kernel_form"]
    input_dim = params_model["input_dim"]
    if kernel_type != "se":
        raise NotImplementedError(f"kernel_type={kernel_type!r}: only the squared-exponential family is implemented on the B200 path")
    if input_dim == 1:
        if params_model.get("distance_func"):
            raise NotImplementedError("distance_func (periodic distance) kernels are not on the B200 path")
        return SquaredExponential(1, "product")
    if input_dim == 2:
        if kernel_form not in ("product", "additive"):
            raise NotImplementedError(f"kernel_form={kernel_form!r} is not on the B200 path")
        return SquaredExponential(2, kernel_form)
    if input_dim == 3:
        return SquaredExponential(3, "product")  # the reference ignores kernel_form for 3-D inputs (kernels.py:419-426)
    raise ValueError(f"input_dim={input_dim}")
