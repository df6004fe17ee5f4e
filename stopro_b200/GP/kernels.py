"""Kernel factory with the reference's signature: ``define_kernel(params_model)`` (GP/kernels.py:331-427).

The reference returns a Python callable that JAX differentiates; here the object only *names* the kernel
(family, form, input dimension) -- the derivatives are closed forms evaluated on the GPU -- but it stays callable
on scalars / points (numpy) with the reference's semantics so user code that probes ``Kernel(r1, r2, theta)`` works.
The squared-exponential family used by the reference's configurations and the Matern-5/2, 7/2, 9/2 kernels are on the
GPU path.
"""
import numpy as np


class SquaredExponential:
    """K_1d_SquareExp / K_2d_SquareExp_{Add,Pro} / K_3d_SquareExp_Pro (GP/kernels.py:36-77); theta in log space."""

    kernel_type = "se"

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


class Matern:
    """K_{2d,3d}_Matern{52,72,92}_{Add,Pro} (GP/kernels.py:127-205): q(rho) exp(-rho) per dimension, rho = kappa |s| / l."""

    KAPPA = {"mt52": np.sqrt(5.0), "mt72": np.sqrt(7.0), "mt92": 3.0}
    Q = {"mt52": [1.0, 1.0, 1.0 / 3.0], "mt72": [1.0, 1.0, 2.0 / 5.0, 1.0 / 15.0],
         "mt92": [1.0, 1.0, 3.0 / 7.0, 2.0 / 21.0, 1.0 / 105.0]}

    def __init__(self, input_dim, form, kernel_type):
        self.input_dim, self.form, self.kernel_type = int(input_dim), form, kernel_type

    @property
    def product_form(self):
        return self.form == "product"

    def __call__(self, r1, r2, theta):
        r1 = np.atleast_1d(np.asarray(r1, dtype=np.float64))
        r2 = np.atleast_1d(np.asarray(r2, dtype=np.float64))
        theta = np.asarray(theta, dtype=np.float64)
        rho = self.KAPPA[self.kernel_type] * np.abs(r1 - r2) * np.exp(-theta[1:1 + self.input_dim])
        m = np.polynomial.polynomial.polyval(rho, self.Q[self.kernel_type]) * np.exp(-rho)
        return np.exp(theta[0]) * (np.prod(m) if self.product_form else np.sum(m))

    def __repr__(self):
        return f"Matern({self.kernel_type!r}, input_dim={self.input_dim}, form={self.form!r})"


def define_kernel(params_model, lbox=None):
    """Same keys as the reference: kernel_type, kernel_form, distance_func, input_dim (GP/kernels.py:331-427).
    On the B200 path: the squared exponential (1-D, 2-D additive / product, 3-D product) and the Matern-5/2, 7/2, 9/2
    kernels in the combinations the reference defines (2-D additive / product; 3-D mt92 product)."""
    kernel_type = params_model["kernel_type"]
    kernel_form = params_model["kernel_form"]
    input_dim = params_model["input_dim"]
    if kernel_type in ("mt52", "mt72", "mt92"):
        if input_dim == 2 and kernel_form in ("product", "additive"):
            return Matern(2, kernel_form, kernel_type)
        if input_dim == 3 and kernel_type == "mt92":
            return Matern(3, "product", kernel_type)  # K_3d_Matern92_Pro, kernel_form ignored (kernels.py:419-426)
        raise NotImplementedError(f"kernel_type={kernel_type!r} with input_dim={input_dim}, kernel_form={kernel_form!r} "
                                  "is not defined by the reference")
    if kernel_type != "se":
        raise NotImplementedError(f"kernel_type={kernel_type!r}: only the squared-exponential and Matern-5/2, 7/2, 9/2 "
                                  "families are implemented on the B200 path")
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
