"""Scalar covariance functions of the reference, restated with torch (float64).

Follows /root/reference/GP/kernels.py:
  K_SquareExp :36-37, K_1d_SquareExp :47-49, K_2d_SquareExp_Add :57-61,
  K_2d_SquareExp_Pro :64-68, K_3d_SquareExp_Pro :71-77, define_kernel :331-427.
Only the squared-exponential family the BASELINE configs use is restated.
theta is in log space: [log gamma, log l_x, (log l_y, (log l_z))].
"""
import torch


def se_1d_factor(x1, x2, logl):
    # kernels.py:36-37
    return torch.exp(-0.5 * ((x1 - x2) * torch.exp(-logl)) ** 2)


def k_1d_se(r1, r2, th):
    # kernels.py:47-49
    return torch.exp(th[0]) * se_1d_factor(r1, r2, th[1])


def k_2d_se_add(r1, r2, th):
    # kernels.py:57-61
    return torch.exp(th[0]) * (se_1d_factor(r1[0], r2[0], th[1]) + se_1d_factor(r1[1], r2[1], th[2]))


def k_2d_se_pro(r1, r2, th):
    # kernels.py:64-68
    return torch.exp(th[0]) * (se_1d_factor(r1[0], r2[0], th[1]) * se_1d_factor(r1[1], r2[1], th[2]))


def k_3d_se_pro(r1, r2, th):
    # kernels.py:71-77
    return torch.exp(th[0]) * (
        se_1d_factor(r1[0], r2[0], th[1]) * se_1d_factor(r1[1], r2[1], th[2]) * se_1d_factor(r1[2], r2[2], th[3])
    )


def define_kernel_ref(params_model):
    """kernels.py:331-427 restricted to kernel_type == 'se'.

    Quirk kept: for input_dim == 3 the kernel_form is ignored (:419-426).
    """
    kt, kf, dim = params_model["kernel_type"], params_model["kernel_form"], params_model["input_dim"]
    if kt != "se":
        raise NotImplementedError("oracle restates the squared-exponential kernels only")
    if dim == 1:
        if params_model.get("distance_func"):
            raise NotImplementedError("distance_func kernels are outside the hot path")
        return k_1d_se
    if dim == 2:
        return {"additive": k_2d_se_add, "product": k_2d_se_pro}[kf]
    if dim == 3:
        return k_3d_se_pro
    raise ValueError(dim)
