"""Host-side model of the assembly kernel's arithmetic (csrc/pigp_assemble.cu): the descriptor -> dense polynomial
expansion of its prologue (herm_coef, the parity classes, the -2k rule for d/dlog l) and the evaluation order of rpoly_eval / upoly_eval,
restated in numpy and checked against the closed-form oracle for every block of the Stokes libraries.  It pins the
algebra the CUDA code relies on without needing a GPU."""
import itertools

import numpy as np
import pytest

from oracle import closed_form
from stopro_b200 import operators

MAX_DEG = 4


KEXP = {1: [(2,), (1,), (0,)],
        2: [(2, 0), (1, 1), (1, 0), (0, 2), (0, 1), (0, 0)],
        3: [(2, 0, 0), (1, 1, 0), (1, 0, 1), (1, 0, 0), (0, 2, 0), (0, 1, 1), (0, 1, 0), (0, 0, 2), (0, 0, 1), (0, 0, 0)]}


def rpoly(c, x, dim):
    """rpoly_eval of the kernel: R(x_0 .. x_{D-1}), total degree <= 2, coefficients in KEXP order."""
    if dim == 1:
        return (c[0] * x[0] + c[1]) * x[0] + c[2]
    if dim == 2:
        t0 = c[0] * x[0] + (c[1] * x[1] + c[2])
        t1 = c[3] * x[1] + c[4]
        return t0 * x[0] + (t1 * x[1] + c[5])
    t0 = c[0] * x[0] + (c[1] * x[1] + (c[2] * x[2] + c[3]))
    t1 = c[4] * x[1] + (c[5] * x[2] + c[6])
    return t0 * x[0] + (t1 * x[1] + ((c[7] * x[2] + c[8]) * x[2] + c[9]))


def herm_coef(n, i, a):
    if i > n or (n - i) & 1:
        return 0.0
    table = {0: {0: 1.0}, 1: {1: -a}, 2: {0: -a, 2: a * a}, 3: {1: 3 * a * a, 3: -a ** 3},
             4: {0: 3 * a * a, 2: -6 * a ** 3, 4: a ** 4}}
    return table[n][i]


def horner1(c, s):
    """upoly_eval of the kernel (additive form): c4 .. c0."""
    acc = c[0]
    for ci in c[1:]:
        acc = acc * s + ci
    return acc


def kernel_model(terms, theta, s, dim, product):
    """-> value, d/dtheta (per group: [log gamma, log l_0 ..]) of the block at separation s, the way k_blocks forms them."""
    n_groups = len(theta) // (1 + dim)
    val, grad = 0.0, np.zeros(len(theta))
    for g in range(n_groups):
        run = [t for t in terms if t[0] == g]
        if not run:
            continue
        gamma = np.exp(theta[g * (1 + dim)])
        a = np.exp(-2.0 * theta[g * (1 + dim) + 1:(g + 1) * (1 + dim)])
        if product:
            # runs of equal (group, parity pattern), as pigp_plan_create sorts them
            classes = sorted({tuple(max(o, 0) & 1 for o in t[2]) for t in run})
            E = gamma * np.exp(-0.5 * np.sum(a * s * s))
            x = s * s
            for pm in classes:
                sub = [t for t in run if tuple(max(o, 0) & 1 for o in t[2]) == pm]
                coef = np.zeros((1 + dim, len(KEXP[dim])))
                for m, kx in enumerate(KEXP[dim]):
                    for _, c, order in sub:
                        p, ks = c, []
                        for d in range(dim):
                            n = max(order[d], 0)
                            i = pm[d] + 2 * kx[d]
                            p *= herm_coef(n, i, a[d])
                            ks.append((n + i) >> 1)
                        coef[0, m] += p
                        for d in range(dim):
                            coef[1 + d, m] += -2.0 * ks[d] * p
                pref = np.prod([s[d] if pm[d] else 1.0 for d in range(dim)])
                P = rpoly(coef[0], x, dim)
                val += pref * P * E
                grad[g * (1 + dim)] += pref * P * E
                for d in range(dim):
                    grad[g * (1 + dim) + 1 + d] += pref * E * (a[d] * x[d] * P + rpoly(coef[1 + d], x, dim))
        else:
            for d in range(dim):
                c0, c1 = np.zeros(5), np.zeros(5)
                for m in range(5):
                    i = MAX_DEG - m
                    for _, c, order in run:
                        if order[d] < 0:
                            continue
                        p = c * herm_coef(order[d], i, a[d])
                        c0[m] += p
                        c1[m] += -2.0 * ((order[d] + i) >> 1) * p
                t = a[d] * s[d] * s[d]
                E = gamma * np.exp(-0.5 * t)
                P = horner1(c0, s[d])
                val += P * E
                grad[g * (1 + dim)] += P * E
                grad[g * (1 + dim) + 1 + d] += E * (t * P + horner1(c1, s[d]))
    return val, grad


@pytest.mark.parametrize("dim,product", [(1, True), (2, True), (2, False), (3, True), (3, False)])
def test_expansion_matches_closed_form(dim, product):
    rng = np.random.default_rng(dim * 10 + product)
    if dim == 1:
        obs, fields = operators.scalar_observables(1)
    else:
        obs, fields = operators.stokes_observables(dim)
    names = [n for n in obs if not n.startswith("dif")]
    theta = rng.normal(0.0, 0.4, len(fields) * (1 + dim))
    r, rp = rng.normal(size=(1, dim)), rng.normal(size=(1, dim))
    s = (r - rp)[0]
    form = "product" if product else "additive"
    n_checked = 0
    for na, nb in itertools.product(names, names):
        terms = operators.block_terms(obs[na], obs[nb], fields, dim, product)
        if not terms:
            continue
        # the same block from the closed-form oracle, monomial by monomial (oracle.closed_form.eval_operator semantics)
        want, want_g = 0.0, np.zeros(len(theta))
        for g, c, order in terms:
            th = theta[g * (1 + dim):(g + 1) * (1 + dim)]
            gamma, a = np.exp(th[0]), np.exp(-2.0 * th[1:])
            fac, dfac = [], []
            for d in range(dim):
                if order[d] < 0:
                    fac.append(None)
                    dfac.append(None)
                    continue
                E = np.exp(-0.5 * a[d] * s[d] ** 2)
                gn = closed_form._g(order[d], s[d:d + 1], a[d])[0]
                fac.append(gn * E)
                dfac.append(-2.0 * a[d] * (closed_form._dg_da(order[d], s[d:d + 1], a[d])[0] - 0.5 * s[d] ** 2 * gn) * E)
            if product:
                v = c * gamma * np.prod([f for f in fac])
                want += v
                want_g[g * (1 + dim)] += v
                for d in range(dim):
                    want_g[g * (1 + dim) + 1 + d] += c * gamma * dfac[d] * np.prod([fac[e] for e in range(dim) if e != d])
            else:
                (d,) = [d for d in range(dim) if order[d] >= 0]
                v = c * gamma * fac[d]
                want += v
                want_g[g * (1 + dim)] += v
                want_g[g * (1 + dim) + 1 + d] += c * gamma * dfac[d]
        got, got_g = kernel_model(terms, theta, s, dim, product)
        scale = max(abs(want), np.max(np.abs(want_g)), 1e-30)
        assert abs(got - want) <= 1e-12 * scale, (na, nb)
        assert np.max(np.abs(got_g - want_g)) <= 1e-12 * scale, (na, nb)
        n_checked += 1
    assert n_checked >= 4
