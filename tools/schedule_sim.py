"""Discrete-event timing model of the multi-rank schedules (planning aid, no GPU): predicts the duration of one
NLL + gradient evaluation for a given N, number of GPUs and schedule from the operation lists of tests/dist_model.py
(the mirror of csrc/pigp_dist.cu) and a per-operation cost table calibrated on the round-1 timelines.

    python tools/schedule_sim.py [N]

Model: per rank the streams A (chain), B (L^-T products), C (publication), D (bulk updates of the panel schedule)
execute their operations in order; an operation starts when its predecessor in the stream, the events it waits for and
the peer flags it waits for are all done.  Operations longer than 100 us ("bulk" GEMMs) of one rank additionally share
that rank's GPU: they are serialised on a per-rank bulk resource (small kernels are assumed to slip in between).
It ignores SM-slot starvation of small kernels behind long CTAs (findings), so it is optimistic for the
recursive schedule at large P -- use it to compare schedules, not to quote numbers.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from dist_model import AsyncModel  # noqa: E402

TILE = 128


def gemm_rate(k_elems):
    """measured TFLOP/s of k_gemm as a function of K (profiles / per-launch dumps of round 1)"""
    table = [(128, 11.0), (256, 20.0), (512, 27.0), (1024, 29.5), (2048, 32.5), (4096, 33.5), (1 << 30, 34.0)]
    for k, r in table:
        if k_elems <= k:
            return r
    return 34.0


def cost_us(rk, op, nvlink_gbs=500.0):
    kind = op[0]
    t3 = 2.0 * TILE ** 3  # flops of one tile product
    P, T = rk.P, rk.T
    if kind == "nop":
        return 0.0
    if kind in ("wait", "wait_panel"):
        return 6.0
    if kind == "potf2":
        return 55.0
    if kind == "push_diag":
        return 12.0
    if kind == "place_diag":
        return 5.0
    if kind == "trsm":
        cnt = rk.count_own(op[1] + 1, rk.gy + 1)
        return max(20.0, cnt * t3 / (gemm_rate(128) * 1e6) * 4.0) if cnt else 0.0
    if kind == "trtri_leaf":
        cnt = rk.count_own(0, op[1])
        return max(20.0, cnt * t3 / (gemm_rate(128) * 1e6) * 4.0) if cnt else 0.0
    if kind == "push_panel":
        cnt = rk.count_own(op[1] + 1, T)
        return 15.0 + cnt * TILE * TILE * 8.0 * (P - 1) / (nvlink_gbs * 1e3)
    if kind == "update":
        k0, k1, jc0, jc1 = op[1:]
        tiles = sum((min(jc1, i + 1) if i < T else jc1) - jc0 for i in rk.own_tiles(jc0, rk.gy + 1) if (i >= jc0))
        tiles = max(tiles, 0)
        return max(20.0, tiles * (k1 - k0) * t3 / (gemm_rate((k1 - k0) * TILE) * 1e6)) if tiles else 0.0
    if kind == "trtri_update":
        k0, k1, jc0, jc1 = op[1:]
        prod = sum((k1 - max(k0, j)) * (jc1 - jc0) for j in rk.own_tiles(0, k1) if max(k0, j) < k1)
        return max(20.0, prod * t3 / (gemm_rate((k1 - k0) * TILE) * 1e6)) if prod else 0.0
    if kind == "nll":
        return 60.0
    if kind == "push_y":
        own = rk.own_tiles(0, T)
        return 20.0 + sum((T - j) * TILE * TILE * 8.0 for j in own) * (P - 1) / (nvlink_gbs * 1e3) if P > 1 else 0.0
    if kind == "alpha":
        return 300.0
    if kind == "lauum":
        prod = sum((T - i) * (i + 1) for i in rk.own_tiles(0, T))
        return prod * t3 / (gemm_rate(1 << 20) * 1e6) + 2000.0 / P  # + fused gradient reduction and assembly share
    raise ValueError(kind)


def simulate(n, world, schedule="recursive", panel=8):
    m = AsyncModel(None, None, world, tile=TILE, schedule=schedule, panel=panel, n=n)
    for rk in m.ranks:
        rk.t_stream = {s: 0.0 for s in rk.q}
        rk.t_event = {}
        rk.t_bulk = 0.0
    t_flag = {}   # (rank, flag) -> time it is visible on that rank
    done = 0
    total = sum(len(q) for rk in m.ranks for q in rk.q.values())
    while done < total:
        advanced = False
        for rk in m.ranks:
            for s, q in rk.q.items():
                while rk.head[s] < len(q):
                    item = q[rk.head[s]]
                    op = item["op"]
                    if not all(e in rk.t_event for e in item["wait"]):
                        break
                    start = max([rk.t_stream[s]] + [rk.t_event[e] for e in item["wait"]])
                    if op[0] == "wait":
                        if (rk.r, op[1]) not in t_flag:
                            break
                        start = max(start, t_flag[(rk.r, op[1])])
                    elif op[0] == "wait_panel":
                        need = [(rk.r, ("PANEL", op[1], src)) for src in range(rk.P) if src != rk.r]
                        if not all(f in t_flag for f in need):
                            break
                        start = max([start] + [t_flag[f] for f in need])
                    dur = cost_us(rk, op)
                    if dur > 100.0:  # bulk GEMMs of one rank share its GPU
                        start = max(start, rk.t_bulk)
                        rk.t_bulk = start + dur
                    end = start + dur
                    rk.t_stream[s] = end
                    if item["rec"] is not None:
                        rk.t_event[item["rec"]] = end
                    if op[0] == "push_diag":
                        for q2 in m.ranks:
                            t_flag[(q2.r, ("DIAG", op[1]))] = end
                    elif op[0] == "push_panel":
                        for q2 in m.ranks:
                            if q2 is not rk:
                                t_flag[(q2.r, ("PANEL", op[1], rk.r))] = end
                    elif op[0] == "push_y":
                        for q2 in m.ranks:
                            if q2 is not rk:
                                t_flag[(q2.r, ("YDONE", rk.r))] = end
                    rk.head[s] += 1
                    done += 1
                    advanced = True
        if not advanced:
            raise RuntimeError("dead-lock in the timing model")
    return max(max(rk.t_stream.values()) for rk in m.ranks) * 1e-3  # ms


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    measured = {1: 254.8, 2: 137.9, 4: 81.2, 8: 56.7} if n == 20000 else {}
    print(f"N = {n}: predicted ms per NLL+grad evaluation (measured round-1 values in brackets)")
    print(f"{'GPUs':>5} {'recursive':>12} " + " ".join(f"{'panels W=' + str(w):>13}" for w in (4, 8, 16)))
    for world in (1, 2, 4, 8):
        row = [simulate(n, world, "recursive")] + [simulate(n, world, "panels", w) for w in (4, 8, 16)]
        meas = f" [{measured[world]:.1f}]" if world in measured else ""
        print(f"{world:>5} {row[0]:>12.1f}{meas} " + " ".join(f"{v:>13.1f}" for v in row[1:]))
