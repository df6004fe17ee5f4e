"""Host-side models of the diagonal-tile kernels (csrc/pigp_dense.cu), restated in numpy so that their algebra is pinned
without a GPU:

  * warp_factor32 / warp_inverse32: the 32 x 32 block is eliminated in column PAIRS.  Every lane applies the same
    formulas (lane J holds d0 in a[J], lane J + 1 holds b and d1), the next 2 x 2 pivot block is rebuilt by all lanes from
    seven exchanged numbers, det = d0 d1 - b^2 is formed with the rounding of d0 d1 compensated, a non-positive pivot is
    reported like LAPACK (1-based index of the first one, even or odd column), and the inverse is solved right-looking two
    columns at a time from the same published columns;
  * k_trsm_blk: X = A inv(L)^T by block forward substitution over 32-column blocks, each diagonal solve done with the
    block's explicit inverse plus one refinement step -- row-wise backward stable where a plain product with the
    inverse of the whole tile is not.
"""
import numpy as np
import pytest
import scipy.linalg


def rsqrt_pivot(d):
    with np.errstate(invalid="ignore", divide="ignore"):
        return 1.0 / np.sqrt(d)


def pivot_block(d0, d1, b, col, fail):
    dd = d0 * d1
    dde = np.float64(np.longdouble(d0) * np.longdouble(d1) - np.longdouble(dd))  # fma(d0, d1, -dd)
    det = (dd - b * b) + dde
    if fail == 0 and not d0 > 0.0:
        fail = col + 1
    if fail == 0 and not det > 0.0:
        fail = col + 2
    r0, rdet = rsqrt_pivot(d0), rsqrt_pivot(det)
    return (r0, b * r0, d0 * r0, rdet), fail


def factor32(S):
    """Lane-parallel model: a[i, :] is lane i's row, r[:, c] lane c's column of the inverse."""
    n = 32
    a = np.tril(S).astype(np.float64)
    piv = np.diag(S).astype(np.float64).copy()
    r = np.eye(n)
    L = np.zeros((n, n))
    W = np.zeros((n, n))
    lanes = np.arange(n)
    fail = 0
    sc, fail = pivot_block(piv[0], piv[1], a[1, 0], 0, fail)
    with np.errstate(invalid="ignore", over="ignore"):
        for j in range(0, n, 2):
            r0, l10, l00, rdet = sc
            l0 = a[:, j] * r0                                  # every lane, rows above the block compute garbage
            l1 = ((a[:, j + 1] - l0 * l10) * l00) * rdet
            piv = piv - l0 * l0 - l1 * l1
            if j + 3 < n:                                      # the seven numbers the next pivot block depends on
                b = a[j + 3, j + 2] - l0[j + 3] * l0[j + 2] - l1[j + 3] * l1[j + 2]
                sc, fail = pivot_block(piv[j + 2], piv[j + 3], b, j + 2, fail)
            L[:, j] = np.where(lanes >= j, l0, 0.0)            # masked store of the two finished columns
            L[:, j + 1] = np.where(lanes > j, l1, 0.0)
            r1 = rdet * l00
            w0 = r[j, :] * r0                                  # inverse: rows j, j + 1 of W for every column (lane)
            w1 = (r[j + 1, :] - l10 * w0) * r1
            W[j, :], W[j + 1, :] = w0, w1
            for k in range(j + 2, n):
                a[:, k] = a[:, k] - l0 * l0[k] - l1 * l1[k]
                r[k, :] = r[k, :] - l0[k] * w0 - l1[k] * w1
    return L, W, fail


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_pairwise_elimination_is_a_cholesky_factorisation(seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((32, 40))
    S = X @ X.T + 1e-3 * np.eye(32)
    L, W, fail = factor32(S)
    assert fail == 0
    ref = np.linalg.cholesky(S)
    assert np.max(np.abs(L - ref)) <= 1e-12 * np.max(np.abs(ref))
    assert np.max(np.abs(np.triu(L, 1))) == 0.0 and np.max(np.abs(np.triu(W, 1))) == 0.0
    assert np.max(np.abs(W @ ref - np.eye(32))) <= 1e-9


@pytest.mark.parametrize("col", [0, 1, 6, 7, 30, 31])
def test_first_non_positive_pivot_is_reported_like_lapack(col):
    rng = np.random.default_rng(col)
    X = rng.standard_normal((32, 32))
    S = X @ X.T / 32 + np.eye(32)
    ref = np.linalg.cholesky(S)
    S[col, col] = np.sum(ref[col, :col] ** 2) - 0.5            # Schur complement at `col` becomes -0.5
    L, _, fail = factor32(S)
    assert fail == col + 1
    assert np.isnan(L[col, col])
    if col:
        assert np.max(np.abs(L[:col, :col] - ref[:col, :col])) <= 1e-12


def trsm_blocked(A, L, W_diag):
    """k_trsm_blk: block forward substitution with refined diagonal solves (W_diag[j] = explicit inverse of L_jj)."""
    X = np.zeros_like(A)
    for j in range(4):
        J = slice(32 * j, 32 * j + 32)
        T = A[:, J].copy()
        for i in range(j):
            I = slice(32 * i, 32 * i + 32)
            T -= X[:, I] @ L[J, I].T
        X0 = T @ W_diag[j].T
        R = T - X0 @ L[J, J].T
        X[:, J] = X0 + R @ W_diag[j].T
    return X


def test_block_substitution_with_refinement_is_backward_stable():
    """An ill-conditioned lower factor (the Cholesky factor of a squared-exponential matrix with a 1e-10 jitter): the
    product with the inverse of the whole 128 x 128 tile leaves a residual ~ cond * u; block substitution with refined
    32 x 32 solves is at the level of LAPACK's trsm."""
    rng = np.random.default_rng(3)
    x = np.sort(rng.random(128))
    K = np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2 / 0.05 ** 2) + 1e-10 * np.eye(128)
    L = np.linalg.cholesky(K)
    A = rng.standard_normal((64, 128))
    W_full = np.linalg.inv(L)
    W_diag = [np.linalg.inv(L[32 * j:32 * j + 32, 32 * j:32 * j + 32]) for j in range(4)]
    X_inv = A @ W_full.T
    X_blk = trsm_blocked(A, L, W_diag)
    X_ref = scipy.linalg.solve_triangular(L, A.T, lower=True).T

    def backward(X):  # row-wise relative residual |X L^T - A| / (|X| |L^T| + |A|)
        return np.max(np.abs(X @ L.T - A) / (np.abs(X) @ np.abs(L.T) + np.abs(A)))

    assert backward(X_ref) <= 1e-14
    assert backward(X_blk) <= 5.0 * max(backward(X_ref), 2.0 ** -52)
    assert backward(X_inv) >= 100.0 * backward(X_blk)          # what the refinement and the small blocks buy


def tile_inverse(L, W_diag):
    """k_tile_inv: the inverse of a 128 x 128 lower factor from its four inverted diagonal 32-blocks -- first the two 64 x 64
    diagonal halves (W_ba = -W_b (L_ba W_a)), then W21 = -W22 (L21 W11)."""
    W = np.zeros((128, 128))
    for h in range(2):
        a, b = slice(64 * h, 64 * h + 32), slice(64 * h + 32, 64 * h + 64)
        W[a, a], W[b, b] = W_diag[2 * h], W_diag[2 * h + 1]
        W[b, a] = -W_diag[2 * h + 1] @ (L[b, a] @ W_diag[2 * h])
    lo, hi = slice(0, 64), slice(64, 128)
    T = L[hi, lo] @ W[lo, lo]
    W[hi, lo] = -W[hi, hi] @ T
    return W


def test_tile_inverse_from_diagonal_blocks():
    rng = np.random.default_rng(9)
    X = rng.standard_normal((128, 160))
    L = np.linalg.cholesky(X @ X.T / 128 + np.eye(128))
    W_diag = [np.linalg.inv(L[32 * j:32 * j + 32, 32 * j:32 * j + 32]) for j in range(4)]
    W = tile_inverse(L, W_diag)
    assert np.max(np.abs(np.triu(W, 1))) == 0.0
    assert np.max(np.abs(W @ L - np.eye(128))) <= 1e-13
