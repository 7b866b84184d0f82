#!/usr/bin/env python
"""Host <-> device copy rates with every rank copying AT THE SAME TIME (one process per GPU, torchrun), with and
without binding each rank (CPU affinity + first touch of its pinned buffers) to its GPU's NUMA node:

    torchrun --nproc-per-node N scripts/host_path_probe.py [--mib 2048]

Prints one JSON line per mode from rank 0: per-rank GB/s for H2D alone, D2H alone and both directions at once,
the GPU's NUMA node, the CPUs the rank ran on and the node its pinned pages live on.  This is the measurement
behind bench.py's e2e leg at N > 1 (the e2e pipeline moves 2 x 8 B x cells per step and rank across PCIe)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import bind_to_gpu_numa_node, gpu_numa_node, pages_numa_node  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.mib * (1 << 20) // 8

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for mode in ("unbound", "numa-bound"):
        info = {"gpu_numa": gpu_numa_node(local)}
        if mode == "numa-bound":
            info["bound"] = bind_to_gpu_numa_node(local)
        info["cpus"] = len(os.sched_getaffinity(0))
        h_in = torch.empty(n, dtype=torch.float64, pin_memory=True)
        h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
        h_in.fill_(0.5)
        h_out.fill_(0.0)
        info["pinned_pages_numa"] = pages_numa_node(h_in)
        d_in = torch.empty(n, dtype=torch.float64, device="cuda")
        d_out = torch.ones(n, dtype=torch.float64, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def h2d():
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)

        def d2h():
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)

        def timed(fn):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.reps):
                fn()
            barrier()
            return (time.perf_counter() - t0) / args.reps

        gb = 8 * n / 1e9
        res = {"h2d": gb / timed(h2d), "d2h": gb / timed(d2h), "both_each": gb / timed(lambda: (h2d(), d2h()))}
        rows = [None] * world
        payload = dict(info, **{k: round(v, 1) for k, v in res.items()})
        if world > 1:
            dist.all_gather_object(rows, payload)
        else:
            rows = [payload]
        if rank == 0:
            print(json.dumps({"mode": mode, "ranks": world, "mib_per_copy": args.mib,
                              "sum_both_each_gbs": round(sum(r["both_each"] for r in rows), 1), "per_rank": rows}), flush=True)
        del h_in, h_out, d_in, d_out
        barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
