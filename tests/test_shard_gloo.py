"""world_size-2 gloo test of the N>1 host logic (parameter broadcast, timing reduction, stream sharding)."""
import ctypes as C
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lvm_b200 import capi
from lvm_b200.shard import broadcast_params, max_over_ranks, shard_streams, sum_over_ranks


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = None
    if rank == 0:
        p = capi.McParams()
        capi.lib().mc_params_from_ui(C.byref(p), capi.MODE_LAPLACE, 20, 50.0, 0.4, 3.0, 25, 6, 30.0)
    got = broadcast_params(p, dist)
    tmax = max_over_ranks(1.0 + rank, dist)
    tsum = sum_over_ranks(10.0 * (rank + 1), dist)
    q.put((rank, got.levels, got.coLow, got.coHigh, got.chromAttenuation, got.coWavelength, tmax, tsum,
           shard_streams(7, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_broadcast_and_reduce(built):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = capi.McParams()
    capi.lib().mc_params_from_ui(C.byref(ref), capi.MODE_LAPLACE, 20, 50.0, 0.4, 3.0, 25, 6, 30.0)
    for r in res:
        assert r[1:6] == (6, ref.coLow, ref.coHigh, ref.chromAttenuation, ref.coWavelength)
        assert r[6] == 2.0 and r[7] == 30.0
    assert res[0][8] == [0, 1, 2, 3] and res[1][8] == [4, 5, 6]


def test_shard_streams_partition():
    for total in (1, 7, 8, 16, 129):
        for world in (1, 2, 4, 8):
            parts = [shard_streams(total, r, world) for r in range(world)]
            flat = [s for p in parts for s in p]
            assert flat == list(range(total))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
