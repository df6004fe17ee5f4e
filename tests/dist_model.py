"""numpy model of the block-cyclic multi-rank evaluation of stopro_b200/csrc/pigp_dist.cu (TEST INFRASTRUCTURE).

It mirrors the *host-side schedule* of the CUDA implementation -- ownership map, global-layout buffers, the merged
recursion (`leaf` / `rec`), the Y = L^-T formulation with triangular k-ranges, the private y tile, K^-1 = Y Y^T on own
row tiles, the flag protocol (DIAG / PANEL / YDONE) -- with dense numpy tiles standing in for the kernels, so that the
index arithmetic and the dependency structure of the N > 1 path can be checked on a CPU:

  * every rank's operation list is generated in the order the C++ issues it (one in-order queue per rank: the side and
    publication streams are folded into program order, which is one valid execution of them);
  * a round-robin scheduler runs the ranks against each other; an operation that waits for a flag that is not up makes
    its rank yield, and "no rank can advance" is reported as a dead-lock;
  * a read of data that a peer has not published yet shows up as a wrong result (buffers start as NaN).

`schedule="panels"` models the panel schedule with look-ahead (`chol_lookahead`, `pigp_set_lookahead`).  The tile size is
a parameter (4 in the tests instead of 128).
"""
import numpy as np


class Rank:
    def __init__(self, rank, world, n, tile, alloc=True):
        self.r, self.P, self.n, self.t = rank, world, n, tile
        self.T = -(-n // tile)
        self.npad = self.T * tile
        self.gy = self.first_own(self.T)
        if alloc:  # tools/schedule_sim.py builds the programs only (timing model, no numerics)
            self.L = np.full(((self.T + world) * tile, self.npad), np.nan)
            self.Y = np.full((self.npad, self.npad), np.nan)
            self.invd = [np.full((tile, tile), np.nan) for _ in range(self.T)]
        self.flags = {}
        self.ops, self.pc = [], 0

    # pigp_dsolver::first_own / count_own
    def first_own(self, a):
        return a + ((self.r - a) % self.P)

    def count_own(self, a, b):
        f = self.first_own(a)
        return (b - f + self.P - 1) // self.P if f < b else 0

    def own_tiles(self, a, b):
        f = self.first_own(a)
        return list(range(f, b, self.P))

    def rows(self, tile_index):
        return slice(tile_index * self.t, (tile_index + 1) * self.t)


class Model:
    def __init__(self, K, y, world, tile=4, schedule="recursive", panel=2, n=None):
        """K = None builds the operation lists only (for n points), without buffers."""
        n = len(y) if n is None else n
        self.world, self.tile, self.n = world, tile, n
        self.ranks = [Rank(r, world, n, tile, alloc=K is not None) for r in range(world)]
        self.T = self.ranks[0].T
        for rk in self.ranks:
            if K is not None:
                self._assemble(rk, K, y)
            self._program(rk, schedule, panel)

    # ---- pigp_dsolver_nll_grad prologue: own rows of K (lower), identity padding, y tile, zeroed own rows of Y
    def _assemble(self, rk, K, y):
        t, n, npad = rk.t, rk.n, rk.npad
        Kp = np.eye(npad)
        Kp[:n, :n] = K
        for i in rk.own_tiles(0, rk.T):
            rk.L[rk.rows(i), :] = np.tril(Kp)[rk.rows(i), :]     # lower part only; the strict upper part is never read
            rk.Y[rk.rows(i), :] = 0.0
        ytile = np.zeros((t, npad))
        ytile[0, :n] = y
        rk.L[rk.rows(rk.gy), :] = ytile

    # ---- operation lists in the C++ issue order
    def _program(self, rk, schedule, panel):
        ops = rk.ops

        def leaf(k):
            mine = k % rk.P == rk.r
            if mine:
                ops.append(("potf2", k))
                ops.append(("push_diag", k))
            else:
                ops.append(("wait", ("DIAG", k)))
            ops.append(("trsm", k))
            ops.append(("push_panel", k))
            if mine:
                ops.append(("place_diag", k))
            ops.append(("trtri_leaf", k))

        def rec(c0, nt):
            if nt == 1:
                return leaf(c0)
            n1 = nt // 2
            n2 = nt - n1
            rec(c0, n1)
            k1 = c0 + n1 - 1
            ops.append(("wait_panel", k1))
            ops.append(("update", c0, c0 + n1, c0 + n1, c0 + nt))        # K1 = [c0, c0+n1), J2 = [c0+n1, c0+nt)
            ops.append(("trtri_update", c0, c0 + n1, c0 + n1, c0 + nt))
            rec(c0 + n1, n2)

        T = rk.T
        if schedule == "recursive":
            rec(0, T)
        else:  # chol_lookahead
            j0 = 0
            while j0 < T:
                j1 = min(j0 + panel, T)
                rec(j0, j1 - j0)
                if j1 >= T:
                    break
                jn = min(j1 + panel, T)
                ops.append(("wait_panel", j1 - 1))
                ops.append(("update", j0, j1, j1, jn))       # panel p -> columns of panel p + 1
                ops.append(("update", j0, j1, jn, T))        # ... -> the rest
                ops.append(("trtri_update", j0, j1, j1, T))  # V_p
                j0 = j1
        for k in range(T):
            ops.append(("wait", ("DIAG", k)))
        ops.append(("nll",))
        ops.append(("push_y",))
        for src in range(rk.P):
            if src != rk.r:
                ops.append(("wait", ("YDONE", src)))
        ops.append(("alpha",))
        ops.append(("lauum",))

    # ---- scheduler
    def run(self):
        progress = True
        while progress:
            progress = False
            for rk in self.ranks:
                while rk.pc < len(rk.ops) and self._ready(rk, rk.ops[rk.pc]):
                    self._exec(rk, rk.ops[rk.pc])
                    rk.pc += 1
                    progress = True
        stuck = [(rk.r, rk.ops[rk.pc]) for rk in self.ranks if rk.pc < len(rk.ops)]
        if stuck:
            raise RuntimeError(f"dead-lock: {stuck}")
        return self

    def _ready(self, rk, op):
        if op[0] == "wait":
            return rk.flags.get(op[1], False)
        if op[0] == "wait_panel":
            return all(rk.flags.get(("PANEL", op[1], src), False) for src in range(rk.P) if src != rk.r)
        return True

    def _peers(self, rk):
        return [q for q in self.ranks if q is not rk]

    def _exec(self, rk, op):
        t, T, ld = rk.t, rk.T, rk.npad
        kind = op[0]
        if kind in ("wait", "wait_panel"):
            return
        if kind == "potf2":
            k = op[1]
            A = rk.L[rk.rows(k), k * t:(k + 1) * t]
            Lkk = np.linalg.cholesky(np.tril(A) + np.tril(A, -1).T)
            rk.L[rk.rows(k), k * t:(k + 1) * t] = Lkk
            rk.invd[k] = np.linalg.inv(Lkk)
        elif kind == "push_diag":                       # k_push_diag: inv(L_kk) and diag(L_kk) to every peer, DIAG flag everywhere
            k = op[1]
            for q in self._peers(rk):
                q.invd[k] = rk.invd[k].copy()
                idx = np.arange(k * t, (k + 1) * t)
                q.L[idx, idx] = rk.L[idx, idx]
            for q in self.ranks:
                q.flags[("DIAG", k)] = True
        elif kind == "trsm":                            # L_ik = A_ik inv(L_kk)^T, own row tiles in [k+1, gy]
            k = op[1]
            for i in rk.own_tiles(k + 1, rk.gy + 1):
                blk = rk.L[rk.rows(i), k * t:(k + 1) * t]
                rk.L[rk.rows(i), k * t:(k + 1) * t] = blk @ rk.invd[k].T
        elif kind == "push_panel":                      # k_push_panel: own rows of the matrix proper (not the y tile)
            k = op[1]
            for i in rk.own_tiles(k + 1, T):
                for q in self._peers(rk):
                    q.L[rk.rows(i), k * t:(k + 1) * t] = rk.L[rk.rows(i), k * t:(k + 1) * t]
            for q in self._peers(rk):
                q.flags[("PANEL", k, rk.r)] = True
        elif kind == "update":                          # C[i, J] -= L[i, Kr] L[J, Kr]^T, lower part, own rows >= jc0
            k0, k1, jc0, jc1 = op[1:]
            for i in rk.own_tiles(jc0, rk.gy + 1):
                for j in range(jc0, min(jc1, i + 1) if i < T else jc1):
                    rk.L[rk.rows(i), j * t:(j + 1) * t] -= rk.L[rk.rows(i), k0 * t:k1 * t] @ rk.L[rk.rows(j), k0 * t:k1 * t].T
        elif kind == "place_diag":
            k = op[1]
            rk.Y[rk.rows(k), k * t:(k + 1) * t] = rk.invd[k].T
        elif kind == "trtri_leaf":                      # Y[j, k] = R[j, k] inv(L_kk)^T, own rows j < k
            k = op[1]
            for j in rk.own_tiles(0, k):
                rk.Y[rk.rows(j), k * t:(k + 1) * t] = rk.Y[rk.rows(j), k * t:(k + 1) * t] @ rk.invd[k].T
        elif kind == "trtri_update":                    # Y[j, J] -= sum_{k in Kr, k >= j} Y[j, k] L[J, k]^T, own rows j < k1
            k0, k1, jc0, jc1 = op[1:]
            for j in rk.own_tiles(0, k1):
                kb = max(k0, j)                         # kmode 1: k tiles >= the row's own tile
                if kb >= k1:
                    continue
                rk.Y[rk.rows(j), jc0 * t:jc1 * t] -= rk.Y[rk.rows(j), kb * t:k1 * t] @ rk.L[jc0 * t:jc1 * t, kb * t:k1 * t].T
        elif kind == "nll":
            n = rk.n
            d = np.diag(rk.L[:n, :n])
            v = rk.L[rk.gy * t, :]
            rk.v = v.copy()
            rk.nll = 0.5 * float(v @ v) + float(np.sum(np.log(d))) + 0.5 * n * np.log(2.0 * np.pi)
        elif kind == "push_y":                          # k_push_rows: own rows, columns from the diagonal tile on
            for j in rk.own_tiles(0, T):
                for q in self._peers(rk):
                    q.Y[rk.rows(j), j * t:] = rk.Y[rk.rows(j), j * t:]
            for q in self._peers(rk):
                q.flags[("YDONE", rk.r)] = True
        elif kind == "alpha":                           # alpha[j] = sum_{k >= tile(j)} Y[j, k] v[k]
            a = np.zeros(ld)
            for j in range(T):
                a[rk.rows(j)] = rk.Y[rk.rows(j), j * t:] @ rk.v[j * t:]
            rk.alpha = a
        elif kind == "lauum":                           # X[i, j<=i] = sum_{k >= i} Y[i, k] Y[j, k], own row tiles
            rk.X = {}
            for i in rk.own_tiles(0, T):
                rk.X[i] = rk.Y[rk.rows(i), i * t:] @ rk.Y[:(i + 1) * t, i * t:].T
        else:
            raise ValueError(kind)


class AsyncModel(Model):
    """The same numerics with the C++'s stream structure made explicit: per rank the chain stream A, the side stream B
    (Y = L^-T products), the publication stream C and -- panel schedule only -- the bulk stream D, ordered by the events
    the C++ records (ev_diag, ev_upd, ev_pan, ev_next, ev_b, ev_c, ev_d) and by the peer flags.  `run(seed)` executes a
    random interleaving of all (rank, stream) queues that respects those dependencies; a dependency missing from the
    schedule shows up as a wrong result (or NaN) for some seed."""

    def _program(self, rk, schedule, panel):
        rk.q = {"A": [], "B": [], "C": [], "D": []}
        rk.events = set()
        rk.head = {s: 0 for s in rk.q}
        multi = rk.P > 1

        def put(stream, op, wait=(), rec=None):
            rk.q[stream].append(dict(op=op, wait=tuple(wait), rec=rec))

        def leaf(k):
            mine = k % rk.P == rk.r
            if mine:
                put("A", ("potf2", k))
            put("A", ("nop",), rec=("ev_diag", k))
            if mine and multi:
                put("C", ("push_diag", k), wait=[("ev_diag", k)])
            if not mine:
                put("A", ("wait", ("DIAG", k)))
            put("A", ("trsm", k), rec=("ev_upd", k))
            if multi:
                put("C", ("push_panel", k), wait=[("ev_upd", k)])
            if mine:
                put("B", ("place_diag", k), wait=[("ev_diag", k)])
            else:
                put("B", ("wait", ("DIAG", k)), wait=[("ev_diag", k)])
            put("B", ("trtri_leaf", k))

        def rec(c0, nt):
            if nt == 1:
                return leaf(c0)
            n1 = nt // 2
            rec(c0, n1)
            k1 = c0 + n1 - 1
            put("A", ("wait_panel", k1))
            put("A", ("update", c0, c0 + n1, c0 + n1, c0 + nt))
            put("B", ("wait_panel", k1), wait=[("ev_upd", k1)])
            put("B", ("trtri_update", c0, c0 + n1, c0 + n1, c0 + nt))
            rec(c0 + n1, nt - n1)

        T = rk.T
        if schedule == "recursive":
            rec(0, T)
        else:
            j0, p = 0, 0
            while j0 < T:
                j1 = min(j0 + panel, T)
                rec(j0, j1 - j0)
                if j1 >= T:
                    break
                jn = min(j1 + panel, T)
                put("A", ("nop",), rec=("ev_pan", p))
                put("D", ("wait_panel", j1 - 1), wait=[("ev_pan", p)])
                put("D", ("update", j0, j1, j1, jn), rec=("ev_next", p))
                put("D", ("update", j0, j1, jn, T))
                put("A", ("nop",), wait=[("ev_next", p)])
                put("B", ("wait_panel", j1 - 1), wait=[("ev_pan", p)])
                put("B", ("trtri_update", j0, j1, j1, T))
                j0, p = j1, p + 1
            put("D", ("nop",), rec=("ev_d",))
            put("A", ("nop",), wait=[("ev_d",)])
        if multi:  # the C++ joins the publication stream and waits for every DIAG flag only when there are peers
            for k in range(T):
                put("A", ("wait", ("DIAG", k)))
            put("C", ("nop",), rec=("ev_c",))
            put("A", ("nop",), wait=[("ev_c",)])
        put("A", ("nll",))
        put("B", ("nop",), rec=("ev_b",))
        put("A", ("push_y",), wait=[("ev_b",)])
        for src in range(rk.P):
            if src != rk.r:
                put("A", ("wait", ("YDONE", src)))
        put("A", ("alpha",))
        put("A", ("lauum",))

    def run(self, seed=0):
        rng = np.random.default_rng(seed)
        while True:
            ready = []
            for rk in self.ranks:
                for s, q in rk.q.items():
                    h = rk.head[s]
                    if h < len(q) and all(e in rk.events for e in q[h]["wait"]) and self._ready(rk, q[h]["op"]):
                        ready.append((rk, s))
            if not ready:
                break
            rk, s = ready[rng.integers(len(ready))]
            item = rk.q[s][rk.head[s]]
            if item["op"][0] != "nop":
                self._exec(rk, item["op"])
            if item["rec"] is not None:
                rk.events.add(item["rec"])
            rk.head[s] += 1
        stuck = [(rk.r, s, q[rk.head[s]]["op"]) for rk in self.ranks for s, q in rk.q.items() if rk.head[s] < len(q)]
        if stuck:
            raise RuntimeError(f"dead-lock: {stuck[:4]}")
        return self
