"""CPU test of the multi-rank schedule through its numpy model (tests/dist_model.py): for several world sizes, matrix
sizes (with and without padding) and both schedules (the shipped recursive one and the experimental panel look-ahead),
every rank must end with the replicated factor, the NLL, alpha = K^-1 y and its own rows of K^-1 -- and the flag
protocol must not dead-lock under a round-robin interleaving of the ranks."""
import numpy as np
import pytest

from dist_model import Model


def spd(n, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n + 3))
    return A @ A.T + n * np.eye(n), rng.standard_normal(n)


@pytest.mark.parametrize("schedule,panel", [("recursive", 0), ("panels", 1), ("panels", 2), ("panels", 3)])
@pytest.mark.parametrize("world", [1, 2, 3, 4])
@pytest.mark.parametrize("n", [4, 10, 24, 37])
def test_schedule_reproduces_the_dense_results(n, world, schedule, panel):
    tile = 4
    K, y = spd(n, seed=n + world)
    m = Model(K, y, world, tile=tile, schedule=schedule, panel=panel).run()
    T = -(-n // tile)
    npad = T * tile
    Kp = np.eye(npad)
    Kp[:n, :n] = K
    Lref = np.linalg.cholesky(Kp)
    Xref = np.linalg.inv(Kp)
    yp = np.zeros(npad)
    yp[:n] = y
    v = np.linalg.solve(Lref, yp)
    nll = 0.5 * v @ v + np.sum(np.log(np.diag(Lref)[:n])) + 0.5 * n * np.log(2 * np.pi)
    alpha = Xref @ yp
    covered = set()
    for rk in m.ranks:
        L = rk.L[:npad]
        for i in range(T):
            for j in range(i):          # below-diagonal tiles are replicated on every rank
                assert np.allclose(L[rk.rows(i), j * tile:(j + 1) * tile], Lref[rk.rows(i), j * tile:(j + 1) * tile], rtol=1e-10, atol=1e-12)
        assert np.allclose(np.diag(L), np.diag(Lref), rtol=1e-10)      # diagonal entries (log-det) on every rank
        assert rk.nll == pytest.approx(nll, rel=1e-11)
        assert np.allclose(rk.alpha, alpha, rtol=1e-8, atol=1e-10)
        for i, Xi in rk.X.items():
            assert np.allclose(Xi, Xref[rk.rows(i), :(i + 1) * tile], rtol=1e-8, atol=1e-10)
            covered.add(i)
    assert covered == set(range(T))      # every row tile of K^-1 is produced by exactly its owner


def test_missing_publication_is_detected_as_deadlock():
    K, y = spd(16, 0)
    m = Model(K, y, 2, tile=4)
    m.ranks[1].ops = [op for op in m.ranks[1].ops if op[0] != "push_panel"]   # rank 1 never publishes its panel rows
    with pytest.raises(RuntimeError, match="dead-lock"):
        m.run()


@pytest.mark.parametrize("schedule,panel", [("recursive", 0), ("panels", 2), ("panels", 3)])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_random_stream_interleavings_respect_all_dependencies(world, schedule, panel):
    """The stream / event / flag structure of the C++ (chain, side, publication and bulk streams per rank) under random
    interleavings: any missing dependency would give a wrong factor, NLL, alpha or K^-1 row for some seed."""
    from dist_model import AsyncModel

    n, tile = 29, 4
    K, y = spd(n, seed=7)
    T = -(-n // tile)
    npad = T * tile
    Kp = np.eye(npad)
    Kp[:n, :n] = K
    Xref = np.linalg.inv(Kp)
    yp = np.zeros(npad)
    yp[:n] = y
    Lref = np.linalg.cholesky(Kp)
    v = np.linalg.solve(Lref, yp)
    nll = 0.5 * v @ v + np.sum(np.log(np.diag(Lref)[:n])) + 0.5 * n * np.log(2 * np.pi)
    for seed in range(12):
        m = AsyncModel(K, y, world, tile=tile, schedule=schedule, panel=panel).run(seed)
        for rk in m.ranks:
            assert rk.nll == pytest.approx(nll, rel=1e-11), seed
            assert np.allclose(rk.alpha, Xref @ yp, rtol=1e-8, atol=1e-10), seed
            for i, Xi in rk.X.items():
                assert np.allclose(Xi, Xref[rk.rows(i), :(i + 1) * tile], rtol=1e-8, atol=1e-10), seed


def test_the_async_model_detects_a_missing_dependency():
    """Sensitivity check of the model itself: publishing a panel without waiting for the TRSM that produces it
    (dropping the ev_upd wait of the publication stream) must corrupt the result for some interleaving."""
    from dist_model import AsyncModel

    n, tile = 29, 4
    K, y = spd(n, seed=7)
    Kp = np.eye(32)
    Kp[:n, :n] = K
    Xref = np.linalg.inv(Kp)
    bad = 0
    for seed in range(12):
        m = AsyncModel(K, y, 2, tile=tile)
        for rk in m.ranks:
            for item in rk.q["C"]:
                if item["op"][0] == "push_panel":
                    item["wait"] = ()
        try:
            m.run(seed)
            ok = all(np.allclose(Xi, Xref[rk.rows(i), :(i + 1) * tile], rtol=1e-8, atol=1e-10)
                     for rk in m.ranks for i, Xi in rk.X.items())
        except np.linalg.LinAlgError:  # a diagonal tile factored from unpublished (NaN / stale) panel rows
            ok = False
        bad += not ok
    assert bad > 0
