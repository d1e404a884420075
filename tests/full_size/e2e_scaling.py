"""BASELINE configs[4]: sharded end-to-end DAB+ decode -- 2^24 MSC frames (F=3072), Viterbi + superframe RS
check on the device, 1/2/4/8 B200, next to the all-core host CPU reference.

    python tests/full_size/e2e_scaling.py                       (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/full_size/e2e_scaling.py                          (N GPUs)

Strong scaling: the job is --total-frames frames whatever N is.  Each rank owns total/N frames (whole
superframes: 5 consecutive frames).  2^24 frames of soft symbols are 206 GB, more than one GPU holds, so a rank
keeps at most --resident-frames frames (default 2^21 = 25.8 GB) in HBM and decodes its share as repeated passes
over them -- every pass streams 25.8 GB of symbols from HBM, far beyond the 126 MB L2, so a repeated pass costs
what a fresh one does.  At N = 8 each rank decodes its 2^21 frames exactly once.

Traffic: random payload -> RS(120,110) encode, s = 16 codewords per superframe, up to --max-err byte errors
per codeword injected before the convolutional encoder -> K=7 rate-1/4 encode -> AWGN at --ebn0 dB -> u8.
Checked: every superframe the RS stage accepts (ret >= 0) must equal the transmitted payload
(size-independent round-trip property); a slice is also compared bit for bit with the CPU reference chain.
NCCL is used once, after the timed region, to gather the result arrays.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-frames", type=int, default=1 << 24)
    ap.add_argument("--resident-frames", type=int, default=1 << 21)
    ap.add_argument("--framebits", type=int, default=3072)
    ap.add_argument("--ebn0", type=float, default=4.0)
    ap.add_argument("--max-err", type=int, default=3)
    ap.add_argument("--cpu-superframes", type=int, default=4096, help="slice decoded by the CPU reference (0 = skip)")
    ap.add_argument("--no-gather", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import viterbi_dll_b200 as vb
    from viterbi_dll_b200 import dabgen, sharding

    f = args.framebits
    s = f // 192
    total_sf = args.total_frames // 5
    lo, hi = sharding.shard_bounds(total_sf, world, rank)  # superframes of this rank
    my_sf = hi - lo
    res_sf = min(my_sf, -(-args.resident_frames // 5))
    passes = (my_sf + res_sf - 1) // res_sf
    last_sf = my_sf - (passes - 1) * res_sf  # superframes of the final (possibly shorter) pass

    t0 = time.perf_counter()
    syms, payload = dabgen.make_superframe_frames_torch(res_sf, f, args.ebn0, seed=5000 + 17 * rank, device=dev,
                                                       max_err=args.max_err)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    out = torch.full((res_sf, 110 * s), 0xEE, dtype=torch.uint8, device=dev)
    ret = torch.empty((res_sf,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass(nsf):
        vb.dabplus_decode_superframes_device(f, syms[: nsf * 5], out[:nsf], ret[:nsf], stream)

    one_pass(min(res_sf, 4096))  # warm-up: allocations, module load
    one_pass(res_sf)
    barrier()
    l0 = vb.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for p in range(passes):
        one_pass(res_sf if p + 1 < passes else last_sf)
    e1.record(stream)
    torch.cuda.synchronize()
    launches = vb.kernel_launches() - l0
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())

    # ---- checks on the resident shard (one full pass was decoded above) ---------------------------------
    one_pass(res_sf)
    torch.cuda.synchronize()
    ok = ret >= 0
    accepted = int(ok.sum().item())
    wrong = int((out[ok] != payload[ok]).any(dim=1).sum().item())
    corrected = int(ret[ok].sum().item())
    stats = torch.tensor([accepted, wrong, corrected, res_sf], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(stats)
    accepted, wrong, corrected, checked = (int(x) for x in stats.tolist())

    # ---- the one collective: gather the result arrays of all ranks (outside the timed region) -----------
    gather_ms = None
    if world > 1 and not args.no_gather:
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        n_all = res_sf * world
        all_ret = sharding.gather_to_all(ret, n_all, world, rank)
        all_out = sharding.gather_to_all(out, n_all, world, rank)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
        assert all_ret.shape[0] == n_all and all_out.shape == (n_all, 110 * s)
        assert torch.equal(all_ret[rank * res_sf:(rank + 1) * res_sf], ret)
        del all_ret, all_out
        # Every rank must have left the collective before any rank tears its communicator down: in the first
        # 8-GPU run ranks that were done exited while three others were still inside the 5.9 GB all-gather,
        # which then sat in the NCCL watchdog for its full 10 minutes.
        barrier()

    # ---- CPU reference chain on a slice: bit-exact check + all-core timing (rank 0) ----------------------
    cpu = None
    if rank == 0 and args.cpu_superframes > 0:
        import oracle_lib

        chk = oracle_lib.checker()
        cores = oracle_lib.ncores()
        nsl = min(args.cpu_superframes, res_sf)
        h_syms = syms[: nsl * 5].cpu().numpy()
        tc = time.perf_counter()
        if chk.kind == "reference":
            dec = chk.deconvolve_batch_u32(f, h_syms.astype(np.uint32), cores)
        else:
            dec = chk.deconvolve_batch(f, h_syms, cores)
        c_out, c_ret = chk.rs_batch(dec.reshape(nsl, 120 * s), s, fill=0xEE, nthreads=cores)
        cpu_s = time.perf_counter() - tc
        same = bool(np.array_equal(c_ret, ret[:nsl].cpu().numpy()) and np.array_equal(c_out, out[:nsl].cpu().numpy()))
        if not same:
            raise RuntimeError("GPU pipeline differs from the CPU reference chain on the checked slice")
        cpu = {"superframes_per_s": nsl / cpu_s, "frames_per_s": nsl * 5 / cpu_s, "cores": cores, "kind": chk.kind,
               "sample": "%d superframes (%d frames) of rank 0's shard, Viterbi + RS on all %d host threads, "
                         "u32 conversion inside the timed region" % (nsl, nsl * 5, cores),
               "bit_exact_vs_gpu": same}

    if rank == 0:
        frames = total_sf * 5
        line = {
            "config": "BASELINE configs[4]: %d MSC frames (F=%d) -> %d DAB+ superframes (s=%d), Viterbi + RS check "
                      "on device, Eb/N0=%.1f dB, 0-%d byte errors per codeword before the convolutional encoder"
                      % (frames, f, total_sf, s, args.ebn0, args.max_err),
            "n_gpus": world, "scaling": "strong", "ms_total": ms, "passes_per_gpu": passes,
            "resident_frames_per_gpu": res_sf * 5,
            "frames_per_s": frames / (ms * 1e-3), "superframes_per_s": total_sf / (ms * 1e-3),
            "viterbi_gbit_per_s": frames * f / (ms * 1e-3) / 1e9, "gpu_launches_per_rank": launches,
            "checked_superframes": checked, "rs_accepted": accepted, "rs_accepted_but_wrong": wrong,
            "rs_corrected_bytes": corrected, "gather_ms": gather_ms, "generate_s": gen_s, "cpu_reference": cpu,
        }
        print(json.dumps(line))
        if wrong:
            raise SystemExit("payload mismatch on accepted superframes")
    if world > 1:
        barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
