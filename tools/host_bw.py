#!/usr/bin/env python
"""Host <-> device copy ceiling of one box, with NO kernels: every rank (one per GPU, torchrun) streams the bench's
per-step byte counts (199 MB in, 199 MB out for 32 lanes of 1080p BGR) between pinned host memory and its GPU on two
streams, concurrently with all other ranks.  The aggregate GB/s is the ceiling the end-to-end (`e2e`) bench line can
reach at that N, whatever the kernels do: it shows whether e2e scaling is bound by the GPUs or by the host's memory /
PCIe root complexes.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/host_bw.py
Prints one JSON line (rank 0)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from lvm_b200.shard import bind_to_gpu_numa_node
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bind = os.environ.get("HOST_BW_BIND", "1") == "1"
    numa = bind_to_gpu_numa_node(local)[1] if bind else {"numa_node": None, "reason": "binding off"}
    nbytes = 32 * 1920 * 1080 * 3
    depth, iters = 3, int(os.environ.get("HOST_BW_ITERS", "60"))
    h_in = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(depth)]
    h_out = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(depth)]
    for t in h_in + h_out:
        t.fill_(1)                                   # first touch on the bound node
    d_in = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(depth)]
    d_out = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(depth)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def run(n):
        for i in range(n):
            with torch.cuda.stream(s_in):
                d_in[i % depth].copy_(h_in[i % depth], non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out[i % depth].copy_(d_out[i % depth], non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run(6)
    barrier()
    t0 = time.perf_counter()
    run(iters)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    node = torch.tensor([float(numa.get("numa_node") if numa.get("numa_node") is not None else -1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        nodes = [torch.zeros_like(node) for _ in range(world)]
        dist.all_gather(nodes, node)
    else:
        nodes = [node]
    if rank == 0:
        per_dir = world * iters * nbytes / dt.item() / 1e9
        print(json.dumps({"n_gpus": world, "numa_bound": bind, "gpu_numa_nodes": [int(n.item()) for n in nodes],
                          "h2d_GBps": per_dir, "d2h_GBps": per_dir, "both_directions_GBps": 2 * per_dir,
                          "equivalent_e2e_frames_per_s": world * iters * 32 / dt.item(), "bytes_per_direction_per_step": nbytes}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
