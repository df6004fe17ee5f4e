"""CPU tests of the host side of the multi-GPU path (stopro_b200/dist.py): the block-cyclic ownership map, the y-tile
placement, load balance, and the rendezvous (exchange of the per-rank handles) over a world_size-2 gloo group.
The device protocol itself is covered by tests/test_gpu_dist.py."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from stopro_b200 import dist as pd


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("n", [32, 128, 498, 1280, 20000])
def test_every_tile_has_exactly_one_owner(world, n):
    T = pd.n_tiles(n)
    owned = [pd.owned_tiles(r, world, 0, T) for r in range(world)]
    flat = sorted(t for tiles in owned for t in tiles)
    assert flat == list(range(T))
    for r, tiles in enumerate(owned):
        assert all(pd.owner_of_tile(t, world) == r for t in tiles)
        assert tiles == sorted(tiles)
    # sub-ranges (what the recursion asks for): tiles in [lo, hi)
    for lo, hi in [(0, T), (T // 3, T), (T // 2, T // 2 + 1), (T, T)]:
        got = sorted(t for r in range(world) for t in pd.owned_tiles(r, world, lo, hi))
        assert got == list(range(lo, hi))


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_y_tiles_are_private_and_owned(world):
    for n in (100, 128, 129, 20000):
        T = pd.n_tiles(n)
        ys = [pd.y_tile(r, world, n) for r in range(world)]
        assert len(set(ys)) == world                      # one private tile per rank
        assert all(T <= y < T + world for y in ys)        # inside the (T + world)-tile factorisation buffer
        assert all(pd.owner_of_tile(y, world) == r for r, y in enumerate(ys))


def test_load_balance_at_the_benchmark_size():
    for world in (2, 4, 8):
        s = pd.shard_summary(20000, world)
        for phase in ("potrf", "trtri", "lauum"):
            w = np.array([x[phase] for x in s], dtype=float)
            assert w.max() / w.mean() < 1.08, (world, phase, w)  # cyclic dealing keeps every phase within 8 % of the mean
        assert sum(x["tiles"] for x in s) == pd.n_tiles(20000)


def _free_port():
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]


def _rendezvous_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # what connect_ipc exchanges: 64-byte IPC handle + 16-byte device UUID per rank
        mine = bytes([rank]) * 64 + bytes([7]) * 16
        both = pd.exchange_handles(mine, None)
        assert len(both) == world
        for r, b in enumerate(both):
            assert b[:64] == bytes([r]) * 64
        shared = any(b[64:] == mine[64:] for r, b in enumerate(both) if r != rank)
        assert shared  # same fake UUID on both ranks -> the ranks would share a device
        # the ownership map is a pure function of (rank, world): all ranks derive the same partition
        T = pd.n_tiles(1500)
        gathered = [None] * world
        dist.all_gather_object(gathered, pd.owned_tiles(rank, world, 0, T))
        assert sorted(t for tiles in gathered for t in tiles) == list(range(T))
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_rendezvous_over_gloo_world_size_2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_rendezvous_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
