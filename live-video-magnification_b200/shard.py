"""Multi-GPU plumbing for the replica-parallel path: one process per GPU, every rank
serves its own independent streams.  The data path has NO per-frame collective; the only exchange is
a one-time broadcast of the POD parameter block from rank 0, plus a MAX-reduction of the timings for
reporting.  Works with any torch.distributed backend (NCCL on B200s, gloo in the CPU tests)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

from .capi import McParams


def shard_streams(total_streams: int, rank: int, world: int) -> List[int]:
    """Contiguous, balanced partition of stream ids 0..total-1 across ranks (weak scaling uses
    total = lanes_per_gpu * world so every rank gets exactly lanes_per_gpu)."""
    base, rem = divmod(total_streams, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def params_to_bytes(p: McParams) -> bytes:
    return bytes(p)


def params_from_bytes(b: bytes) -> McParams:
    p = McParams()
    assert len(b) == C.sizeof(p)
    C.memmove(C.byref(p), b, C.sizeof(p))
    return p


def broadcast_params(p: Optional[McParams], dist=None, device="cpu", src: int = 0) -> McParams:
    """Rank `src` supplies p; everyone returns the same block (one NCCL / gloo broadcast)."""
    import torch
    n = C.sizeof(McParams)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        assert p is not None
        return p
    buf = torch.zeros(n, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        buf = torch.frombuffer(bytearray(params_to_bytes(p)), dtype=torch.uint8).to(device)
    dist.broadcast(buf, src=src)
    return params_from_bytes(bytes(buf.cpu().numpy().tobytes()))


def max_over_ranks(x: float, dist=None, device="cpu") -> float:
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, dist=None, device="cpu") -> float:
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
