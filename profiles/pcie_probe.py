"""Aggregate pinned host->device copy bandwidth of the box versus the number of ranks copying at once.

    python profiles/pcie_probe.py                                   (1 rank)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29530 \
        profiles/pcie_probe.py [--affinity]                 (N ranks)

Every rank copies a 203 MB pinned buffer (the FIC batch of bench.py) to its GPU 20 times, all ranks at once
(barrier before, max over ranks after) -- the host side of `e2e`, without any kernel.  Variants: --affinity pins
each rank to the CPUs of its GPU's NUMA node before it allocates (first-touch places the pinned pages there).
Also measures D2H alone and H2D + D2H
together.  Prints one JSON line on rank 0; profiles/pcie_r02.json collects them by N.
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def numa_cpus_of_gpu(index):
    try:
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        bus = nv.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()[-12:]  # 0000:xx:yy.z
        node = open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip()
        if int(node) < 0:
            return None, node
        cpus = open("/sys/devices/system/node/node%s/cpulist" % node).read().strip()
        out = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            out.update(range(int(a), int(b or a) + 1))
        return out, node
    except Exception as e:  # noqa: BLE001
        return None, repr(e)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--affinity", action="store_true")
    ap.add_argument("--mb", type=int, default=203)
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    node = None
    if args.affinity:
        cpus, node = numa_cpus_of_gpu(local)
        if cpus:
            allowed = os.sched_getaffinity(0) & cpus
            if allowed:
                os.sched_setaffinity(0, allowed)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = args.mb * 1000 * 1000
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.fill_(rank + 1)  # first touch on this rank's CPUs
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d2 = torch.full((n,), 7, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        barrier()
        return n * reps * world / dt / 1e9

    def h2d():
        with torch.cuda.stream(s_in):
            d.copy_(h, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s_out):
            h2.copy_(d2, non_blocking=True)

    def both():
        h2d()
        d2h()

    res = {"ranks": world, "affinity": bool(args.affinity), "numa_node_rank0": node, "buffer_mb": args.mb,
           "h2d_gbs_aggregate": timed(h2d), "d2h_gbs_aggregate": timed(d2h), "duplex_gbs_each_direction": timed(both)}
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
