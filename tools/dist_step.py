"""Time the block-cyclic multi-rank NLL+grad evaluation with all ranks inside ONE process (one host thread, stream and
-- when the box has enough GPUs -- device per rank; slabs connected by raw pointers / cudaDeviceEnablePeerAccess).

    python tools/dist_step.py N WORLD [REPEATS] [--same-device]
"""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from stopro_b200 import _lib, synthetic
from stopro_b200.dist import DistSolver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
repeats = int(sys.argv[3]) if len(sys.argv) > 3 else 3
same = "--same-device" in sys.argv or torch.cuda.device_count() < world
prof = "--prof" in sys.argv

cfg = synthetic.stokes2d_scaling(n, n_test=16)
r, y, eps, th = cfg["r_train"], cfg["delta_y"], cfg["eps"], cfg["theta0"]
gps, solvers = [None] * world, [None] * world
barrier = threading.Barrier(world)
times = [[] for _ in range(world)]
res = [None] * world


def work(k):
    dev = 0 if same else k
    torch.cuda.set_device(dev)
    _lib.check(_lib.lib().pigp_set_device(dev))
    gp = synthetic.make_model(cfg)
    gp.set_constants(r, y, eps, only_training=True)
    gps[k] = gp
    solvers[k] = DistSolver(gp._training_plan(r), k, world)
    barrier.wait()
    solvers[k].connect_pointers([s.slab()[0] for s in solvers])
    barrier.wait()
    for it in range(repeats + 1):
        if prof and k == 0 and it == repeats:
            _lib.profile_start()
        barrier.wait()
        t0 = time.perf_counter()
        res[k] = solvers[k].nll_grad_host(th, y, eps)
        times[k].append(time.perf_counter() - t0)
        if prof and k == 0 and it == repeats:
            p = _lib.profile_stop()
            print("rank0 classes:", {c: (round(v["ms"], 2), v["launches"]) for c, v in p.items()})
    barrier.wait()


ts = [threading.Thread(target=work, args=(k,)) for k in range(world)]
for t in ts:
    t.start()
for t in ts:
    t.join()
best = min(max(times[k][i] for k in range(world)) for i in range(1, repeats + 1))
print(f"N={n} world={world} {'same device' if same else 'one device per rank'}: {best * 1e3:.1f} ms per NLL+grad "
      f"({1 / best:.2f} evals/s)  nll={res[0][0]:.10g} info={res[0][2]} |grad|={np.linalg.norm(res[0][1]):.6g}")
if world > 1:
    print("ranks agree:", all(res[k][0] == res[0][0] and np.array_equal(res[k][1], res[0][1]) for k in range(world)))
