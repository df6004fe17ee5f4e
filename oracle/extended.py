"""Higher-precision truth for the ill-conditioned configurations (TEST INFRASTRUCTURE, see oracle/__init__.py).

With the schema's jitter eps = 1e-6 the Stokes covariance matrices have cond(K) ~ 1e9 .. 1e11, so two correct float64
evaluations of GP/gp.py:72-89 / :412-488 (the reference under JAX, the reference under the torch shim, this build's
CUDA path) can legitimately differ by cond(K) * 2^-53 > 1e-8.  To decide which side is closer instead of widening the
tolerance, this module evaluates the same quantities in numpy.longdouble (x87 80-bit: 64-bit
mantissa, u = 5.4e-20): closed-form blocks (oracle.closed_form, dtype=longdouble), a blocked Cholesky, the explicit
inverse and the trace formula  dNLL/dtheta_p = 1/2 sum_jk (K^-1 - alpha alpha^T)_jk dK_jk/dtheta_p  (GP/gp.py:480-486),
and the posterior mean / variance (GP/gp.py:91-120).  numpy has no LAPACK for longdouble; everything is built from
longdouble matmuls (slow: minutes at N = 2640, run once when the golden fixtures are generated).
"""
import numpy as np

LD = np.longdouble
BLOCK = 96


def _chol_unblocked(A):
    n = A.shape[0]
    L = np.tril(A).astype(LD)
    for j in range(n):
        d = L[j, j] - np.dot(L[j, :j], L[j, :j])
        if not d > 0:
            raise np.linalg.LinAlgError(f"non-positive pivot at {j}")
        L[j, j] = np.sqrt(d)
        if j + 1 < n:
            L[j + 1:, j] = (L[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    return L


def tri_inv_lower(L):
    """inv(L) for lower-triangular L (recursive 2 x 2 blocking; flops in matmul)."""
    n = L.shape[0]
    if n <= 32:
        X = np.zeros((n, n), dtype=LD)
        for j in range(n):
            X[j, j] = LD(1) / L[j, j]
            for i in range(j + 1, n):
                X[i, j] = -np.dot(L[i, j:i], X[j:i, j]) / L[i, i]
        return X
    h = n // 2
    A = tri_inv_lower(L[:h, :h])
    C = tri_inv_lower(L[h:, h:])
    X = np.zeros((n, n), dtype=LD)
    X[:h, :h] = A
    X[h:, h:] = C
    X[h:, :h] = -(C @ (L[h:, :h] @ A))
    return X


def cholesky_lower(A):
    """Blocked right-looking Cholesky in longdouble; returns L (lower)."""
    n = A.shape[0]
    A = np.array(A, dtype=LD, copy=True)
    for k in range(0, n, BLOCK):
        e = min(k + BLOCK, n)
        Lkk = _chol_unblocked(A[k:e, k:e])
        A[k:e, k:e] = Lkk
        if e < n:
            Wi = tri_inv_lower(Lkk)
            P = A[e:, k:e] @ Wi.T
            A[e:, k:e] = P
            A[e:, e:] -= P @ P.T
    return np.tril(A)


def truth(ref_ld, theta, r_train, delta_y, eps, r_test=None, want_grad=True):
    """ref_ld: oracle.gp_ref.GPRef(..., backend="closed", dtype=np.longdouble) after set_constants.
    Returns a dict of float64 arrays rounded from the longdouble results: nll, grad, alpha, mu, var."""
    th = np.asarray(theta, dtype=LD)
    y = np.asarray(delta_y, dtype=LD)
    S = ref_ld.training_sigma(th, r_train, LD(eps))
    assert S.dtype == LD
    n = len(y)
    L = cholesky_lower(S)
    W = tri_inv_lower(L)
    v = W @ y
    out = {}
    pi_ld = LD(4) * np.arctan(LD(1))
    nll = LD(0.5) * np.dot(v, v) + np.sum(np.log(np.diag(L))) + LD(0.5) * n * np.log(LD(2) * pi_ld)
    out["nll"] = float(nll)
    alpha = W.T @ v
    out["alpha"] = alpha.astype(np.float64)
    if want_grad:
        Kinv = W.T @ W
        M = Kinv - np.outer(alpha, alpha)
        dKs = ref_ld.dK_dtheta(th, r_train, LD(eps))
        out["grad"] = np.array([float(LD(0.5) * np.sum(M * dK)) for dK in dKs])
        del Kinv, M, dKs
    if r_test is not None:
        thk, _ = ref_ld.split_hyp_and_noise(th)
        Kab = ref_ld.mixedK_all(thk, ref_ld._pts(r_test), ref_ld._pts(r_train))
        out["mu"] = (Kab @ alpha).astype(np.float64)
        V = W @ Kab.T
        # only the per-variable diagonal of K_aa is needed: evaluate the test blocks' diagonals pointwise
        kdiag = []
        pts = ref_ld._pts(r_test)
        for i, p in enumerate(pts):
            name = ref_ld.table["test"][i][0]
            step = 256
            d = np.concatenate([np.diag(ref_ld.block(name, p[a:a + step], p[a:a + step], thk)) for a in range(0, len(p), step)])
            kdiag.append(d)
        out["var"] = (np.concatenate(kdiag) - np.sum(V * V, axis=0)).astype(np.float64)
    return out
