#!/usr/bin/env python
"""Pinned host -> device copy bandwidth per rank and in aggregate, with the box's NUMA / PCIe topology.

    python tools/h2d_bw.py                                        # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_bw.py [--bind]

Every rank copies a pinned 1 GiB buffer to its GPU, alone (ranks take turns) and then all ranks at once; rank 0 prints
one JSON object: per-rank solo GB/s, per-rank concurrent GB/s, aggregate concurrent GB/s, each GPU's NUMA node and CPU
affinity as NVML reports them.  `--bind` first pins the rank (and hence its pinned allocation: first touch) to the CPUs
NVML lists as local to its GPU -- the NUMA-local configuration of the end-to-end scaling question:
bench.py's e2e number is bounded by exactly this aggregate."""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def gpu_affinity(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, v in enumerate(words) for b in range(64) if (v >> b) & 1]
        try:
            node = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:
            node = None
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        return cpus, node, bus if isinstance(bus, str) else bus.decode()
    except Exception as exc:
        return [], None, repr(exc)


def copy_gbs(x, y, reps=5):
    for _ in range(2):
        y.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        y.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    return reps * x.numel() * 4 / (time.perf_counter() - t) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", action="store_true", help="bind the rank to its GPU's local CPUs before allocating")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cpus, node, bus = gpu_affinity(local)
    bound = False
    if args.bind and cpus:
        try:
            os.sched_setaffinity(0, set(cpus) & os.sched_getaffinity(0) or os.sched_getaffinity(0))
            bound = True
        except OSError:
            pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    x = torch.empty(1 << 28, dtype=torch.float32, pin_memory=True)       # 1 GiB, first-touched by this rank
    x.fill_(1.0)
    y = torch.empty_like(x, device="cuda")
    solo = 0.0
    for r in range(world):                                               # ranks take turns
        if world > 1:
            dist.barrier()
        if r == rank:
            solo = copy_gbs(x, y)
    if world > 1:
        dist.barrier()
    conc = copy_gbs(x, y, reps=10)                                      # all ranks at once
    rec = {"rank": rank, "solo_gbs": round(solo, 1), "concurrent_gbs": round(conc, 1), "gpu_numa_node": node,
           "gpu_bus": bus, "gpu_local_cpus": "%d cpus (%s..%s)" % (len(cpus), cpus[0] if cpus else None, cpus[-1] if cpus else None),
           "rank_affinity_cpus": len(os.sched_getaffinity(0)), "bound_to_gpu_cpus": bound}
    if world > 1:
        recs = [None] * world
        dist.all_gather_object(recs, rec)
        dist.destroy_process_group()
    else:
        recs = [rec]
    if rank == 0:
        numa = None
        try:
            numa = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        except OSError:
            pass
        print("H2D_JSON " + json.dumps({"n_gpus": world, "bind": args.bind, "host_numa_nodes": numa, "host_cpus": os.cpu_count(),
                          "aggregate_concurrent_gbs": round(sum(r["concurrent_gbs"] for r in recs), 1),
                          "ranks": recs}), flush=True)


if __name__ == "__main__":
    main()
