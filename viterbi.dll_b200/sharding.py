"""Multi-GPU partitioning for the FEC hot path.

Frames and superframes are independent (the reference keeps no state between calls:
deconvolve.cpp:116-132 re-initialises the metrics, rschecksf.cpp:72 uses stack scratch), so the
batch is cut into contiguous index ranges, one per rank, with no collective on the data path.
The only collective is the optional gather of the result bitstreams.
"""
from __future__ import annotations


def shard_bounds(n: int, world: int, rank: int, align: int = 1) -> tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank`; boundaries fall on multiples of `align`
    (64 = one warp group of the Viterbi kernel, 5 = the logical frames of one DAB+ superframe)
    except the very last one."""
    if world < 1 or not (0 <= rank < world) or align < 1:
        raise ValueError("bad shard request")
    units = (n + align - 1) // align
    lo_u = units * rank // world
    hi_u = units * (rank + 1) // world
    return min(lo_u * align, n), min(hi_u * align, n)


def all_shards(n: int, world: int, align: int = 1):
    return [shard_bounds(n, world, r, align) for r in range(world)]


def gather_to_all(local, n_total: int, world: int, rank: int, align: int = 1, group=None):
    """all-gather row-sharded results (torch tensors, CPU/gloo or CUDA/NCCL) into [n_total, ...]."""
    import torch
    import torch.distributed as dist

    bounds = all_shards(n_total, world, align)
    lo, hi = bounds[rank]
    assert local.shape[0] == hi - lo
    width = max(h - l for l, h in bounds)
    padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: hi - lo] = local
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    if out.is_cuda:
        # The collective is complete on return, not merely enqueued: a rank that goes on to tear its communicator
        # down while peers are still inside the all-gather leaves them in the NCCL watchdog (seen once at N = 8).
        torch.cuda.current_stream(out.device).synchronize()
    parts = [out[r * width : r * width + (h - l)] for r, (l, h) in enumerate(bounds)]
    return torch.cat(parts, dim=0)

