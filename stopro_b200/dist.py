"""Multi-GPU evaluation: K dealt block-cyclically (128-row tiles) over the ranks of one NVLink/NVSwitch box.

New functionality (the reference is single device).  One process per GPU under torchrun;
``torch.distributed`` is used only to exchange the CUDA-IPC handles of the solver slabs (any backend: gloo or nccl)
and by the callers for barriers / timing.  The data path is inside libpigp.so: stores into peer memory over NVLink
from the producing kernels and epoch flags (csrc/pigp_dist.cu).

The pure layout helpers below (owner_of_tile, owned_tiles, shard_summary) are the host-side logic the CPU tests
cover with a world_size-2 gloo group.
"""
import ctypes as C

import numpy as np

from . import _lib

TILE = _lib.TILE


def n_tiles(n):
    return (int(n) + TILE - 1) // TILE


def owner_of_tile(t, world):
    """Rank that assembles, factors and inverts 128-row tile ``t``."""
    return int(t) % int(world)


def owned_tiles(rank, world, lo, hi):
    """Tiles t in [lo, hi) owned by ``rank`` (ascending)."""
    first = lo + ((rank - lo) % world)
    return list(range(first, hi, world))


def y_tile(rank, world, n):
    """Tile index of the rank's private y row tile: the first tile index >= n_tiles(n) it owns."""
    t = n_tiles(n)
    return t + ((rank - t) % world)


def shard_summary(n, world):
    """Per-rank share of the three O(N^3) phases, in tile-products (for load-balance checks and DESIGN.md).
    POTRF trailing update: row tile i does sum_{k<i} (i-k) ~ i^2/2 products; TRTRI row tile j: (T-j)^2/2;
    LAUUM row tile i: (i+1)(T-i)."""
    T = n_tiles(n)
    out = []
    for r in range(world):
        tiles = owned_tiles(r, world, 0, T)
        out.append(dict(rank=r, tiles=len(tiles),
                        potrf=sum(i * (i + 1) / 2 for i in tiles),
                        trtri=sum((T - j) * (T - j + 1) / 2 for j in tiles),
                        lauum=sum((i + 1) * (T - i) for i in tiles)))
    return out


def exchange_handles(handle_bytes, group=None):
    """all-gather of one bytes object per rank -> list indexed by rank (works on gloo and nccl groups)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, bytes(handle_bytes), group=group)
    return out


class DistSolver:
    """One rank's share of the sharded NLL + gradient evaluation (pigp_dsolver_* of include/pigp.h)."""

    def __init__(self, plan, rank=0, world=1):
        if not plan.symmetric:
            raise ValueError("DistSolver needs the symmetric training plan")
        self.plan, self.rank, self.world = plan, int(rank), int(world)
        h = C.c_void_p()
        _lib.check(_lib.lib().pigp_dsolver_create(plan.handle, self.rank, self.world, C.byref(h)))
        self.handle = h
        self._opened = []

    # -- wiring
    def slab(self):
        ptr, nbytes = C.c_void_p(), C.c_int64()
        _lib.check(_lib.lib().pigp_dsolver_slab(self.handle, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    def ipc_handle(self):
        buf = (C.c_char * _lib.IPC_HANDLE_BYTES)()
        _lib.check(_lib.lib().pigp_dsolver_ipc_handle(self.handle, buf))
        return bytes(buf)

    def connect_pointers(self, slabs):
        """Several ranks in one process: ``slabs[r]`` is rank r's slab address."""
        arr = (C.c_void_p * self.world)(*[C.c_void_p(p) for p in slabs])
        _lib.check(_lib.lib().pigp_dsolver_connect(self.handle, arr))

    def connect_ipc(self, group=None):
        """One process per GPU: exchange CUDA-IPC handles over torch.distributed and map every peer's slab."""
        if self.world == 1:
            return
        uuid = (C.c_char * 16)()
        _lib.check(_lib.lib().pigp_device_uuid(uuid))
        mine = bytes(uuid)
        both = exchange_handles(self.ipc_handle() + mine, group)
        handles = [b[:_lib.IPC_HANDLE_BYTES] for b in both]
        shared = any(b[_lib.IPC_HANDLE_BYTES:] == mine for r, b in enumerate(both) if r != self.rank)
        slabs = []
        for r, hb in enumerate(handles):
            if r == self.rank:
                slabs.append(self.slab()[0])
                continue
            ptr = C.c_void_p()
            _lib.check(_lib.lib().pigp_ipc_open(hb, C.byref(ptr)))
            self._opened.append(ptr.value)
            slabs.append(ptr.value)
        self.connect_pointers(slabs)
        _lib.check(_lib.lib().pigp_dsolver_set_shared_device(self.handle, int(shared)))

    # -- evaluation
    def nll_grad(self, theta_ptr, y_ptr, eps, nll_ptr, grad_ptr, info_ptr=None, stream=None):
        _lib.check(_lib.lib().pigp_dsolver_nll_grad(self.handle, theta_ptr, y_ptr, float(eps), nll_ptr, grad_ptr, info_ptr,
                                                    stream))

    def nll_grad_host(self, theta, y, eps, want_grad=True):
        p = self.plan
        th = p._theta(theta)
        yy = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
        if yy.size != p.rows:
            raise ValueError(f"delta_y has {yy.size} entries, the training set has {p.rows}")
        nll = C.c_double()
        grad = np.empty(p.theta_len, dtype=np.float64)
        info = C.c_int32()
        _lib.check(_lib.lib().pigp_dsolver_nll_grad_host(self.handle, th.ctypes.data, yy.ctypes.data, float(eps),
                                                         int(want_grad), C.addressof(nll), grad.ctypes.data,
                                                         C.addressof(info)))
        return nll.value, (grad if want_grad else None), info.value

    def reset(self):
        """Re-arm after a failed evaluation (a lost or late peer): call on every rank, then barrier on the host."""
        _lib.check(_lib.lib().pigp_dsolver_reset(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().pigp_dsolver_destroy(self.handle)
            self.handle = None
            for ptr in self._opened:
                _lib.lib().pigp_ipc_close(ptr)
            self._opened = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
