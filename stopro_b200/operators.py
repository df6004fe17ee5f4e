"""Block-descriptor compiler: observables -> covariance blocks -> monomial tables for the CUDA evaluator.

The reference writes every block of its library by hand as a sum of autodiff operators
(GP/gp_2D_stokes_independent.py:22-246, GP/gp_3D_stokes_independent.py:25-239, operators in
GP/gp_2D.py:16-86 and GP/gp_3D.py:12-35).  Here a block is *derived*: each observed quantity is a linear
differential operator acting on independent latent fields,

    u_c   = u_c                      f_c = -Laplace(u_c) + d_c p        (Stokes momentum balance)
    p     = p                        div = sum_c d_c u_c                (continuity)
    difA  = A(r + lbox) - A(r)                                           (periodic difference, GP/gp.py:374-410)

and cov(A(r), B(r')) = sum_fields A_r B_r' k_field(r, r').  With d/dr'_d = -d/dr_d on a stationary kernel, a pair of
monomials with derivative orders (alpha, beta) contributes  (-1)^|beta| * gamma * prod_d G_{alpha_d + beta_d}(s_d)
for the product-form squared exponential (GP/kernels.py:64-77) and only its single-dimension part for the additive
form (GP/kernels.py:57-61).  The result is the `pigp_block_desc` table of include/pigp.h.
"""
from collections import OrderedDict

from . import _lib

ABSENT = -1


def _unit(i, dim, k=1):
    return tuple(k if d == i else 0 for d in range(dim))


class Observable:
    """Linear operator on the latent fields: {field: [(coef, alpha)]}, optionally followed by the periodic shift."""

    def __init__(self, parts, shift=False):
        self.parts = parts
        self.shift = shift

    def shifted(self):
        return Observable(self.parts, True)


def stokes_observables(dim):
    """Observables of the Stokes-independent models; latent fields ux, uy(, uz), p in theta order
    (ind_uxux, ind_uyuy, (ind_uzuz,) ind_pp: gp_2D_stokes_independent.py:17-19, gp_3D_stokes_independent.py:18-21)."""
    comps = "xyz"[:dim]
    zero = (0,) * dim
    obs = OrderedDict()
    for i, c in enumerate(comps):
        obs["u" + c] = Observable({"u" + c: [(1.0, zero)]})
    obs["p"] = Observable({"p": [(1.0, zero)]})
    for i, c in enumerate(comps):
        obs["f" + c] = Observable({"u" + c: [(-1.0, _unit(d, dim, 2)) for d in range(dim)], "p": [(1.0, _unit(i, dim))]})
    obs["div"] = Observable({"u" + c: [(1.0, _unit(i, dim))] for i, c in enumerate(comps)})
    for name in list(obs):
        obs["dif" + name] = obs[name].shifted()
    fields = ["u" + c for c in comps] + ["p"]
    return obs, fields


def scalar_observables(dim):
    """Single latent field y: plain GP (GP/gp_naive.py) and the 1-D y / y'' model (GP/gp_1D_laplacian.py:26-33)."""
    zero = (0,) * dim
    obs = OrderedDict(y=Observable({"y": [(1.0, zero)]}))
    obs["ly"] = Observable({"y": [(1.0, _unit(d, dim, 2)) for d in range(dim)]})
    return obs, ["y"]


def block_terms(obs_a, obs_b, fields, dim, product_form):
    """cov(A(r), B(r')) as a list of (group, coef, orders) monomials, merged and sorted by group."""
    acc = OrderedDict()
    for g, field in enumerate(fields):
        for ca, alpha in obs_a.parts.get(field, []):
            for cb, beta in obs_b.parts.get(field, []):
                coef = ca * cb * (-1.0) ** sum(beta)
                order = tuple(a + b for a, b in zip(alpha, beta))
                if product_form:
                    keys = [order]
                else:
                    active = [d for d in range(dim) if order[d] > 0]
                    if len(active) == 0:
                        keys = [tuple(0 if d == e else ABSENT for d in range(dim)) for e in range(dim)]
                    elif len(active) == 1:
                        keys = [tuple(order[d] if d == active[0] else ABSENT for d in range(dim))]
                    else:
                        keys = []  # mixed derivatives of a sum of 1-D kernels vanish
                for key in keys:
                    acc[(g, key)] = acc.get((g, key), 0.0) + coef
    terms = [(g, c, key) for (g, key), c in acc.items() if c != 0.0]
    terms.sort(key=lambda t: t[0])
    return terms


def make_desc(obs_a, obs_b, fields, dim, product_form):
    terms = block_terms(obs_a, obs_b, fields, dim, product_form)
    if len(terms) > _lib.MAX_TERMS:
        raise ValueError(f"block needs {len(terms)} monomials, more than PIGP_MAX_TERMS")
    d = _lib.BlockDesc()
    d.n_terms = len(terms)
    d.shift_first = 1 if (obs_a.shift and terms) else 0
    d.shift_second = 1 if (obs_b.shift and terms) else 0
    for k, (g, c, order) in enumerate(terms):
        d.terms[k].group = g
        d.terms[k].coef = c
        for i in range(3):
            d.terms[k].order[i] = order[i] if i < dim else ABSENT
    return d


def describe(desc, dim):
    """Human-readable form of a descriptor (for tests / debugging)."""
    out = []
    for k in range(desc.n_terms):
        t = desc.terms[k]
        out.append((t.group, t.coef, tuple(t.order[i] for i in range(dim))))
    return out, bool(desc.shift_first), bool(desc.shift_second)


def parse_block_name(name, obs):
    """'Kuxfx' -> (obs['ux'], obs['fx']): the naming scheme of the reference's block library (K<a><b>)."""
    if not name.startswith("K"):
        raise KeyError(name)
    rest = name[1:]
    for a in sorted(obs, key=len, reverse=True):
        if rest.startswith(a) and rest[len(a):] in obs:
            return obs[a], obs[rest[len(a):]]
    raise KeyError(name)

