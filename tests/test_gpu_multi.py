"""GPU tests of the multi-device paths: they use min(2, device_count) -- or all -- devices
of the box, so they run on one GPU and exercise the real thing on two or more.

* the *_multi C-ABI calls: one host batch, sharded inside the library over the selected devices, one process;
* a native host program (tests/host/multi_device_check.cpp, no torch) doing the same plus the NCCL gather of
  device-resident shards (fec_allgather_device);
* the one-process-per-GPU flavour used by bench.py: torch.distributed NCCL ranks, contiguous shards, gather.
Everything is compared bit for bit with the CPU checker."""
import json
import os
import shutil
import socket
import subprocess
import sys

import numpy as np
import pytest

from viterbi_dll_b200 import dabgen

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _device(vb):
    assert vb.lib.fec_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    assert vb.initialize()
    yield
    vb.set_devices(None)


def _device_sets(vb):
    n = vb.lib.fec_device_count()
    sets = [[0], list(range(n))]
    if n >= 2:
        sets += [[n - 1], [1, 0]]
    return sets


def test_device_list_api(vb):
    n = vb.lib.fec_device_count()
    vb.set_devices(None)
    assert vb.get_devices() == list(range(n))
    vb.set_devices([n - 1])
    assert vb.get_devices() == [n - 1]
    import ctypes

    bad = (ctypes.c_int * 2)(0, 0)
    assert vb.lib.fec_set_devices(bad, 2) == vb.FEC_ERR_ARG  # duplicate
    bad = (ctypes.c_int * 1)(n)
    assert vb.lib.fec_set_devices(bad, 1) == vb.FEC_ERR_ARG  # out of range
    assert vb.get_devices() == [n - 1]  # unchanged by the failed calls
    vb.set_devices(None)


@pytest.mark.parametrize("framebits,n", [(768, 20000), (3072, 5000), (100, 333), (768, 1)])
def test_viterbi_multi_matches_checker(vb, checker, framebits, n):
    sym, _ = dabgen.make_frames(n, framebits, 2.5, seed=framebits + n)
    want = checker.deconvolve_batch(framebits, sym)
    pin = vb.host_array(sym.shape)
    pin[:] = sym
    for devs in _device_sets(vb):
        vb.set_devices(devs)
        assert np.array_equal(vb.deconvolve_batch_multi(framebits, sym), want), devs  # pageable buffers
        out = vb.host_array(want.shape)
        out[:] = 0x77
        vb.deconvolve_batch_multi(framebits, pin, out=out)  # pinned buffers
        assert np.array_equal(out, want), devs


def test_rs_and_dabplus_multi_match_checker(vb, checker):
    s = 7
    rx, _, _ = dabgen.make_superframes(9001, s, seed=3)
    want_out, want_ret = checker.rs_batch(rx, s, fill=0x3C)
    syms, _, _ = dabgen.make_superframe_frames(700, 768, 2.0, seed=8, max_err=5)
    dec = checker.deconvolve_batch(768, syms)
    dab_out, dab_ret = checker.rs_batch(dec.reshape(700, 120 * 4), 4, fill=0xEE)
    for devs in _device_sets(vb):
        vb.set_devices(devs)
        out, ret = vb.rs_check_superframe_batch_multi(rx, s, fill=0x3C)
        assert np.array_equal(ret, want_ret) and np.array_equal(out, want_out), devs
        pin_out = vb.host_array(want_out.shape)
        pin_out[:] = 0x3C
        out, ret = vb.rs_check_superframe_batch_multi(rx, s, out=pin_out)  # pinned outVector: zero-copy originals
        assert np.array_equal(ret, want_ret) and np.array_equal(pin_out, want_out), devs
        out, ret = vb.dabplus_decode_superframes_multi(768, syms, fill=0xEE)
        assert np.array_equal(ret, dab_ret) and np.array_equal(out, dab_out), devs
    assert (want_ret < 0).any() and (want_ret >= 0).any()


def test_allgather_device_from_one_process(vb):
    import torch

    n = vb.lib.fec_device_count()
    vb.set_devices(None)
    shards = [torch.full((4096, 96), 10 + i, dtype=torch.uint8, device="cuda:%d" % i) for i in range(n)]
    for i, t in enumerate(shards):
        t[:, 0] = i
    outs = [torch.zeros((n * 4096, 96), dtype=torch.uint8, device="cuda:%d" % i) for i in range(n)]
    vb.allgather_device(shards, outs)
    for i in range(n):
        torch.cuda.synchronize(i)
    want = torch.cat([t.cpu() for t in shards], dim=0)
    for o in outs:
        assert torch.equal(o.cpu(), want)


def test_copy_engine_gather_between_devices(vb):
    """fec_memcpy_d2d_async: every device pushes its shard into every device's array (the gather by the copy engines,
    no kernel and no collective); on one GPU the same call is a plain device-to-device copy."""
    import torch

    n = vb.lib.fec_device_count()
    vb.set_devices(None)
    if n > 1:
        assert vb.lib.fec_enable_peer_access() == 0
    rows = 4099
    shards = [torch.randint(0, 256, (rows, 96), dtype=torch.uint8, device="cuda:%d" % i) for i in range(n)]
    outs = [torch.zeros((n * rows, 96), dtype=torch.uint8, device="cuda:%d" % i) for i in range(n)]
    for i in range(n):
        torch.cuda.synchronize(i)
    for i, src in enumerate(shards):
        with torch.cuda.device(i):
            st = torch.cuda.Stream(device=i)
            for o in outs:
                vb.memcpy_d2d_async(o[i * rows:(i + 1) * rows], src, src.numel(), st)
            st.synchronize()
    want = torch.cat([t.cpu() for t in shards], dim=0)
    for o in outs:
        assert torch.equal(o.cpu(), want)
    vb.memcpy_d2d_async(outs[0], shards[0], 0)  # zero bytes: a no-op
    assert vb.lib.fec_memcpy_d2d_async(None, shards[0].data_ptr(), 16, None) == vb.FEC_ERR_ARG


def test_bcast_stores_results_into_every_copy(vb, checker):
    """The *_bcast calls: the RS kernel stores its results into extra buffers as well -- on this device and, with
    peer access, on the others (the gather fused into the producing kernel)."""
    import torch

    ndev = vb.lib.fec_device_count()
    vb.set_devices(None)
    if ndev > 1:
        assert vb.lib.fec_enable_peer_access() == 0, vb.lib.fec_last_error()
    torch.cuda.set_device(0)
    for s, n in ((1, 5001), (5, 3000), (16, 700)):
        rx, _, _ = dabgen.make_superframes(n, s, seed=70 + s)
        want_out, want_ret = checker.rs_batch(rx, s, fill=0xEE)
        d_rx = torch.from_numpy(rx).cuda()
        out = torch.full((n, 110 * s), 0xEE, dtype=torch.uint8, device="cuda:0")
        ret = torch.empty((n,), dtype=torch.int32, device="cuda:0")
        devs = ["cuda:0", "cuda:%d" % (ndev - 1), "cuda:%d" % (1 % ndev)]
        outs = [torch.full((n, 110 * s), 0xEE, dtype=torch.uint8, device=d) for d in devs]
        rets = [torch.full((n,), -9, dtype=torch.int32, device=d) for d in devs]
        vb.rs_check_superframe_batch_device_bcast(d_rx, s, out, ret, outs, rets)
        torch.cuda.synchronize()
        for o, r in zip([out] + outs, [ret] + rets):
            assert np.array_equal(r.cpu().numpy(), want_ret) and np.array_equal(o.cpu().numpy(), want_out), (s, o.device)
    # the Viterbi -> RS pipeline with copies
    syms, _, _ = dabgen.make_superframe_frames(900, 768, 2.0, seed=4, max_err=5)
    dec = checker.deconvolve_batch(768, syms)
    want_out, want_ret = checker.rs_batch(dec.reshape(900, 480), 4, fill=0xEE)
    d = torch.from_numpy(syms).cuda()
    out = torch.full((900, 440), 0xEE, dtype=torch.uint8, device="cuda:0")
    ret = torch.empty((900,), dtype=torch.int32, device="cuda:0")
    o2 = torch.full((900, 440), 0xEE, dtype=torch.uint8, device="cuda:%d" % (ndev - 1))
    r2 = torch.full((900,), -9, dtype=torch.int32, device="cuda:%d" % (ndev - 1))
    vb.dabplus_decode_superframes_device_bcast(768, d, out, ret, [o2], [r2])
    torch.cuda.synchronize()
    for o, r in ((out, ret), (o2, r2)):
        assert np.array_equal(r.cpu().numpy(), want_ret) and np.array_equal(o.cpu().numpy(), want_out)
    # argument checks: misaligned copy, too many copies
    bad = torch.zeros(900 * 440 + 8, dtype=torch.uint8, device="cuda:0")[1:]
    assert vb.lib.rs_check_superframe_batch_device_bcast(d.data_ptr(), 4, 1, out.data_ptr(), ret.data_ptr(),
                                                         vb._ptr_array([bad.data_ptr()]), vb._ptr_array([r2.data_ptr()]), 1, None) == vb.FEC_ERR_ARG
    assert vb.lib.rs_check_superframe_batch_device_bcast(d.data_ptr(), 4, 1, out.data_ptr(), ret.data_ptr(), None, None, 16, None) == vb.FEC_ERR_ARG


@pytest.mark.timeout(900)
@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
@pytest.mark.parametrize("gather", ["dma", "nccl"])
def test_native_host_decodes_one_batch_on_all_devices(vb, tmp_path, gather):
    """tests/host/multi_device_check.cpp: a C++ host without torch, all GPUs of the box, bit-exact vs the checker;
    once per transport of fec_allgather_device (copy engines, the default, and ncclAllGather)."""
    import oracle_lib

    chk = oracle_lib.checker()
    exe = tmp_path / "multi_device_check"
    subprocess.run(["g++", "-std=c++17", "-O2", "-pthread", "-o", str(exe),
                    os.path.join(ROOT, "tests", "host", "multi_device_check.cpp"), "-ldl"], check=True)
    run = subprocess.run([str(exe), vb.LIB_PATH, chk.lib._name, "16384"], capture_output=True, text=True,
                         env=dict(os.environ, VITERBI_B200_GATHER=gather))
    assert run.returncode == 0, run.stdout + run.stderr
    rep = json.loads(run.stdout.strip().splitlines()[-1])
    assert rep["ok"] and rep["devices"] == vb.lib.fec_device_count()
    assert rep["viterbi_mismatched_frames"] == 0 and rep["rs_mismatch"] == 0 and rep["dabplus_mismatch"] == 0
    assert rep["allgather_mismatch"] == 0
    assert 0 < rep["rs_failed"] < rep["rs_superframes"]  # both RS write paths were exercised


# ---- one process per GPU (the bench.py flavour): NCCL ranks, contiguous shards, result gather ----------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, n, framebits, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import viterbi_dll_b200 as vb
    from viterbi_dll_b200 import dabgen, sharding

    sym, _ = dabgen.make_frames(n, framebits, 3.0, seed=21)  # every rank regenerates the batch, keeps its shard
    lo, hi = sharding.shard_bounds(n, world, rank, align=64)
    local = vb.deconvolve_batch_device(framebits, torch.from_numpy(sym[lo:hi]).cuda())
    allout = sharding.gather_to_all(local, n, world, rank, align=64)  # complete on return
    rx, _, _ = dabgen.make_superframes(n // 4, 6, seed=22)
    lo2, hi2 = sharding.shard_bounds(n // 4, world, rank)
    o = torch.full((hi2 - lo2, 660), 0xEE, dtype=torch.uint8, device="cuda")
    o, r = vb.rs_check_superframe_batch_device(torch.from_numpy(rx[lo2:hi2]).cuda(), 6, o)
    allo = sharding.gather_to_all(o, n // 4, world, rank)
    allr = sharding.gather_to_all(r, n // 4, world, rank)
    # the same RS results again, gathered by the kernel itself: every rank maps every rank's result arrays (CUDA IPC)
    # and its RS kernel stores its shard into all of them over NVLink
    nsf = n // 4
    buf_o = vb.PeerBuffer(nsf * 660, world, rank, rank)
    buf_r = vb.PeerBuffer(nsf * 4, world, rank, rank)
    full_o = buf_o.local.view(nsf, 660)
    full_r = buf_r.local.view(torch.int32)
    full_o.fill_(0xEE)
    full_r.fill_(-9)
    torch.cuda.synchronize()
    dist.barrier()  # every rank's arrays are pre-filled before anybody stores into them
    others = [r_ for r_ in range(world) if r_ != rank]
    vb.rs_check_superframe_batch_device_bcast(torch.from_numpy(rx[lo2:hi2]).cuda(), 6, full_o[lo2:hi2], full_r[lo2:hi2],
                                              [buf_o.peer_ptr(r_, full_o[lo2:hi2]) for r_ in others],
                                              [buf_r.peer_ptr(r_, full_r[lo2:hi2]) for r_ in others])
    torch.cuda.synchronize()
    dist.barrier()  # every rank's kernel has finished: all shards have landed everywhere
    np.save(os.path.join(tmpdir, "bco%d.npy" % rank), full_o.cpu().numpy())
    np.save(os.path.join(tmpdir, "bcr%d.npy" % rank), full_r.cpu().numpy())
    del full_o, full_r
    buf_o.close()
    buf_r.close()
    np.save(os.path.join(tmpdir, "vit%d.npy" % rank), allout.cpu().numpy())
    np.save(os.path.join(tmpdir, "rso%d.npy" % rank), allo.cpu().numpy())
    np.save(os.path.join(tmpdir, "rsr%d.npy" % rank), allr.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_one_process_per_gpu_shards_and_nccl_gather(vb, checker, tmp_path):
    import torch.multiprocessing as mp

    world = min(2, vb.lib.fec_device_count())
    n, framebits = 9000, 768
    mp.spawn(_rank_main, args=(world, _free_port(), n, framebits, str(tmp_path)), nprocs=world, join=True)
    sym, _ = dabgen.make_frames(n, framebits, 3.0, seed=21)
    want = checker.deconvolve_batch(framebits, sym)
    rx, _, _ = dabgen.make_superframes(n // 4, 6, seed=22)
    want_o, want_r = checker.rs_batch(rx, 6, fill=0xEE)
    for rank in range(world):  # every rank holds the whole gathered result
        assert np.array_equal(np.load(tmp_path / ("vit%d.npy" % rank)), want)
        assert np.array_equal(np.load(tmp_path / ("rso%d.npy" % rank)), want_o)
        assert np.array_equal(np.load(tmp_path / ("rsr%d.npy" % rank)), want_r)
        assert np.array_equal(np.load(tmp_path / ("bco%d.npy" % rank)), want_o)  # gathered by the kernel's own peer stores
        assert np.array_equal(np.load(tmp_path / ("bcr%d.npy" % rank)), want_r)
