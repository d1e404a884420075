"""BASELINE configs[2]: batched MSC decode, 262,144 CIF-sized frames (F=3072) per Eb/N0 point, 0..6 dB --
BER/FER of the B200 decoder next to the reference CPU decoder on the SAME soft symbols.

Every frame is decoded by both; outputs are compared bit for bit (so the two BER/FER columns are identical by
construction -- the script asserts it and counts mismatching frames).  Prints one JSON line.

    python tests/full_size/ber_fer_sweep.py [--frames 262144] [--slice 32768]
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
    ap.add_argument("--frames", type=int, default=262144)
    ap.add_argument("--framebits", type=int, default=3072)
    ap.add_argument("--slice", type=int, default=32768, help="frames handed to the CPU reference at a time")
    ap.add_argument("--points", default="0,1,2,3,4,5,6")
    args = ap.parse_args()

    import torch

    import oracle_lib
    import viterbi_dll_b200 as vb
    from viterbi_dll_b200 import dabgen

    assert torch.cuda.is_available()
    dev = torch.device("cuda", 0)
    chk = oracle_lib.checker()
    cores = oracle_lib.ncores()
    n, f = args.frames, args.framebits
    rows = []
    for eb in (float(x) for x in args.points.split(",")):
        syms, bits = dabgen.make_frames_torch(n, f, eb, seed=31000 + int(eb * 10), device=dev, want_bits=True)
        out = torch.empty((n, f // 8), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        vb.deconvolve_batch_device(f, syms, out)  # warm-up
        e0.record()
        vb.deconvolve_batch_device(f, syms, out)
        e1.record()
        torch.cuda.synchronize()
        gpu_ms = e0.elapsed_time(e1)
        err = (out ^ bits)
        # popcount per frame through a 256-entry table
        table = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int32, device=dev)
        bit_errors = int(table[err.long()].sum().item())
        frame_errors = int((err != 0).any(dim=1).sum().item())
        # the reference decoder on the same symbols
        mismatching, cpu_s, cpu_bit_errors, cpu_frame_errors = 0, 0.0, 0, 0
        h_out, h_bits = out.cpu().numpy(), bits.cpu().numpy()
        for lo in range(0, n, args.slice):
            hi = min(n, lo + args.slice)
            h_syms = syms[lo:hi].cpu().numpy()
            if chk.kind == "reference":
                s32 = h_syms.astype(np.uint32)
                t0 = time.perf_counter()
                want = chk.deconvolve_batch_u32(f, s32, cores)
            else:
                t0 = time.perf_counter()
                want = chk.deconvolve_batch(f, h_syms, cores)
            cpu_s += time.perf_counter() - t0
            mismatching += int((want != h_out[lo:hi]).any(axis=1).sum())
            d = want ^ h_bits[lo:hi]
            cpu_bit_errors += int(np.unpackbits(d).sum())
            cpu_frame_errors += int((d != 0).any(axis=1).sum())
        rows.append({"ebn0_db": eb, "frames": n, "ber_b200": bit_errors / (n * f), "fer_b200": frame_errors / n,
                     "ber_reference": cpu_bit_errors / (n * f), "fer_reference": cpu_frame_errors / n,
                     "frames_differing": mismatching, "b200_gbit_per_s": n * f / (gpu_ms * 1e-3) / 1e9,
                     "reference_gbit_per_s": n * f / cpu_s / 1e9})
        assert mismatching == 0, "decoders disagree at %.1f dB" % eb
        del syms, bits, out
    print(json.dumps({"config": "BASELINE configs[2]: %d frames x F=%d per point, AWGN, 8-bit soft symbols" % (n, f),
                      "reference": "%s, %d host threads" % (chk.kind, cores), "points": rows}))


if __name__ == "__main__":
    main()
