"""Closed-form numpy restatement of the reference's Matern-5/2, 7/2 and 9/2 kernels under its autodiff operators
(TEST INFRASTRUCTURE, see oracle/__init__.py).

GP/kernels.py:127-205:  M(x1, x2; logl) = q(rho) exp(-rho),  rho = kappa |x1 - x2| exp(-logl)
    mt52: kappa = sqrt 5, q = 1 + rho + rho^2 / 3
    mt72: kappa = sqrt 7, q = 1 + rho + 2 rho^2 / 5 + rho^3 / 15
    mt92: kappa = 3,      q = 1 + rho + 3 rho^2 / 7 + 2 rho^3 / 21 + rho^4 / 105     (the reference writes it in r = rho / 3)
With s = x1 - x2:  (d/ds)^n M = (kappa / l)^n sgn(s)^n q_n(rho) exp(-rho),  q_0 = q,  q_{n+1} = q_n' - q_n.

Autodiff quirk that parity must reproduce: the reference differentiates through jnp.abs, whose derivative at 0 is
sgn(0) = 0, and sgn itself has derivative 0 everywhere; so for s = 0 EVERY derivative of order >= 1 evaluates to 0
(also the even ones, whose true value is not 0).  Here: sgn(s)^n with sgn(0) = 0 for n >= 1.
d/dlog l [(kappa/l)^n q_n(rho) e^-rho] = (kappa/l)^n e^-rho (-n q_n + rho (q_n - q_n')).
"""
import numpy as np
from numpy.polynomial import polynomial as Poly

KINDS = {
    "mt52": (np.sqrt(5.0), [1.0, 1.0, 1.0 / 3.0]),
    "mt72": (np.sqrt(7.0), [1.0, 1.0, 2.0 / 5.0, 1.0 / 15.0]),
    "mt92": (3.0, [1.0, 1.0, 3.0 / 7.0, 2.0 / 21.0, 1.0 / 105.0]),
}


def q_poly(kind, n):
    """Coefficients (ascending powers of rho) of q_n."""
    q = np.asarray(KINDS[kind][1], dtype=np.float64)
    for _ in range(n):
        q = Poly.polysub(Poly.polyder(q) if len(q) > 1 else np.zeros(1), q)
    return q


def u_poly(kind, n):
    """Coefficients of u_n = -n q_n + rho (q_n - q_n'):  d/dlog l of the n-th derivative factor, over its prefactor."""
    q = q_poly(kind, n)
    dq = Poly.polyder(q) if len(q) > 1 else np.zeros(1)
    return Poly.polyadd(-n * q, Poly.polymul([0.0, 1.0], Poly.polysub(q, dq)))


def factor(kind, n, s, logl, want_dlogl=False):
    """(d/ds)^n M(s) and optionally its derivative with respect to log l (arrays like s)."""
    kappa = KINDS[kind][0]
    c = kappa * np.exp(-logl)
    rho = c * np.abs(s)
    sg = np.sign(s) ** n if n > 0 else np.ones_like(s)
    pref = c ** n * sg * np.exp(-rho)
    val = pref * Poly.polyval(rho, q_poly(kind, n))
    if not want_dlogl:
        return val
    return val, pref * Poly.polyval(rho, u_poly(kind, n))


def eval_terms(kind, terms, r, rp, theta, dim, product, with_grad=False):
    """Block from monomial terms [(group, coef, orders)] (stopro_b200.operators.block_terms): dense (n, m) array, and
    optionally d/dtheta as (len(theta), n, m)."""
    r = np.asarray(r, dtype=np.float64).reshape(len(r), -1)
    rp = np.asarray(rp, dtype=np.float64).reshape(len(rp), -1)
    s = [r[:, None, d] - rp[None, :, d] for d in range(dim)]
    out = np.zeros_like(s[0])
    grad = np.zeros((len(theta),) + s[0].shape) if with_grad else None
    for g, coef, order in terms:
        th = theta[g * (1 + dim):(g + 1) * (1 + dim)]
        gamma = np.exp(th[0])
        dims = [d for d in range(dim) if order[d] >= 0] if not product else list(range(dim))
        fac, dfac = {}, {}
        for d in dims:
            fac[d], dfac[d] = factor(kind, max(order[d], 0), s[d], th[1 + d], want_dlogl=True)
        term = coef * gamma * np.prod([fac[d] for d in dims], axis=0)
        out += term
        if with_grad:
            grad[g * (1 + dim)] += term
            for e in dims:
                grad[g * (1 + dim) + 1 + e] += coef * gamma * dfac[e] * np.prod([fac[d] for d in dims if d != e] + [np.ones_like(s[0])], axis=0)
    return (out, grad) if with_grad else out


def eval_operator(kind, name, r, rp, theta, form, dim, with_grad=False):
    """Same contract as oracle.closed_form.eval_operator (one differential operator of GP/gp_2D.py / gp_3D.py applied to the
    kernel of one hyper-parameter group), for a Matern kernel."""
    from . import closed_form

    terms = []
    for coef, alpha, beta in closed_form.operator_polynomial(name, dim):
        sign = coef * (-1.0) ** sum(beta)
        order = tuple(a + b for a, b in zip(alpha, beta))
        if form == "product":
            terms.append((0, sign, order))
        else:
            active = [d for d in range(dim) if order[d] > 0]
            if len(active) == 0:
                terms += [(0, sign, tuple(0 if d == e else -1 for d in range(dim))) for e in range(dim)]
            elif len(active) == 1:
                terms.append((0, sign, tuple(order[d] if d == active[0] else -1 for d in range(dim))))
    out = eval_terms(kind, terms, r, rp, np.asarray(theta, dtype=np.float64), dim, form == "product", with_grad=with_grad)
    return out
