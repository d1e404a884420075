"""CPU tests of the multi-GPU host logic (world_size 2, gloo): the batch is cut into contiguous
per-rank ranges, every rank decodes its own shard (here with the CPU oracle standing in for the
device), and the gathered bitstreams equal the single-process result.  No collective on the data
path; the only collective is the result gather (viterbi.dll_b200/sharding.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_align():
    from viterbi_dll_b200 import sharding

    for n in (0, 1, 63, 64, 65, 1000, 65536, 262144 + 5):
        for world in (1, 2, 3, 4, 8):
            for align in (1, 5, 64):
                b = sharding.all_shards(n, world, align)
                assert b[0][0] == 0 and b[-1][1] == n
                assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
                assert all(lo % align == 0 for lo, _ in b if lo < n)
                sizes = [hi - lo for lo, hi in b]
                assert max(sizes) - min(sizes) <= align or n < world * align
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, framebits, s_rs, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib
    from viterbi_dll_b200 import dabgen, sharding

    port_lib = oracle_lib.port()
    # every rank regenerates the same global batch from the seed and keeps only its shard
    sym, _ = dabgen.make_frames(n, framebits, 3.0, seed=11)
    lo, hi = sharding.shard_bounds(n, world, rank, align=64)
    local = port_lib.deconvolve_batch(framebits, sym[lo:hi], nthreads=2)
    allout = sharding.gather_to_all(torch.from_numpy(local), n, world, rank, align=64)
    # RS: superframes sharded the same way; return values gathered too
    rx, _, _ = dabgen.make_superframes(n // 8, s_rs, seed=12)
    lo2, hi2 = sharding.shard_bounds(n // 8, world, rank)
    o, r = port_lib.rs_batch(rx[lo2:hi2], s_rs, fill=0xEE, nthreads=2)
    allo = sharding.gather_to_all(torch.from_numpy(o), n // 8, world, rank)
    allr = sharding.gather_to_all(torch.from_numpy(r), n // 8, world, rank)
    if rank == 0:
        np.save(os.path.join(tmpdir, "vit.npy"), allout.numpy())
        np.save(os.path.join(tmpdir, "rs_out.npy"), allo.numpy())
        np.save(os.path.join(tmpdir, "rs_ret.npy"), allr.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_partition_and_gather_matches_single_process(tmp_path):
    import oracle_lib
    from viterbi_dll_b200 import dabgen

    n, framebits, s_rs, world = 200, 768, 3, 2  # 200 frames: shards of 128 + 72 (64-aligned, ragged tail)
    mp.spawn(_worker, args=(world, _free_port(), n, framebits, s_rs, str(tmp_path)), nprocs=world, join=True)
    port_lib = oracle_lib.port()
    sym, _ = dabgen.make_frames(n, framebits, 3.0, seed=11)
    assert np.array_equal(np.load(tmp_path / "vit.npy"), port_lib.deconvolve_batch(framebits, sym))
    rx, _, _ = dabgen.make_superframes(n // 8, s_rs, seed=12)
    o, r = port_lib.rs_batch(rx, s_rs, fill=0xEE)
    assert np.array_equal(np.load(tmp_path / "rs_out.npy"), o)
    assert np.array_equal(np.load(tmp_path / "rs_ret.npy"), r)


def test_configs4_round_sizes_cover_the_share_in_whole_waves():
    """bench.py's configs[4] schedule: rounds sum to the rank's share, respect the resident-symbol cap, are whole waves
    except the last, and shrink towards the end (the last gather is the only exposed one)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    wave = 148 * 16 * 64 // 5
    cap = ((1 << 21) // 5) // wave * wave
    for world in (1, 2, 3, 4, 8):
        my = -(-((1 << 24) // 5) // world)
        sizes = bench.configs4_round_sizes(my, wave, cap)
        assert sum(sizes) == my and max(sizes) <= cap and min(sizes) > 0
        assert all(sz % wave == 0 for sz in sizes[:-1])
        assert wave // 2 <= sizes[-1] <= wave + wave // 2
        tail = sizes[-4:]
        assert tail == sorted(tail, reverse=True)
    for my in (1, 100, wave, wave + wave // 2 + 1, 2 * wave + 5, 14 * wave + 1):
        sizes = bench.configs4_round_sizes(my, wave, cap)
        assert sum(sizes) == my and max(sizes) <= cap and min(sizes) > 0
