"""Measure both Viterbi kernels at several batch sizes (device-resident) and the single-frame drop-in
latency.  Used to place kVitWarpKernelMaxFrames (csrc/fec_internal.h).  Run on a B200:
    python profiles/kernel_crossover.py > gpurun_out/kernel_crossover.jsonl
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viterbi_dll_b200 as vb  # noqa: E402
from viterbi_dll_b200 import dabgen  # noqa: E402

assert vb.initialize()
for f in (768, 3072):
    for n in (1, 64, 296, 512, 1024, 2048, 4096, 6144, 8192, 12288, 16384, 32768, 65536):
        sym, _ = dabgen.make_frames_torch(n, f, 3.0, seed=1, device="cuda")
        out = torch.empty((n, f // 8), dtype=torch.uint8, device="cuda")
        row = {"framebits": f, "frames": n}
        for name, mode in (("pair", vb.VITERBI_PAIR), ("warp", vb.VITERBI_WARP)):
            vb.set_viterbi_kernel(mode)
            for _ in range(3):
                vb.deconvolve_batch_device(f, sym, out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                vb.deconvolve_batch_device(f, sym, out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            row[name + "_ms"] = round(ms, 4)
            row[name + "_gbps"] = round(n * f / ms / 1e6, 2)
        print(json.dumps(row), flush=True)
    # single-frame drop-in latency (host buffers, u32 symbols, synchronous)
    sym1, _ = dabgen.make_frames(1, f, 3.0, seed=2)
    s32 = sym1[0].astype(np.uint32)
    for name, mode in (("pair", vb.VITERBI_PAIR), ("warp", vb.VITERBI_WARP)):
        vb.set_viterbi_kernel(mode)
        for _ in range(5):
            vb.deconvolve(f, s32)
        t0 = time.perf_counter()
        for _ in range(50):
            rc, _ = vb.deconvolve(f, s32)
            assert rc == 0
        print(json.dumps({"framebits": f, "dropin_deconvolve_us": round((time.perf_counter() - t0) / 50 * 1e6, 1), "kernel": name}), flush=True)
vb.set_viterbi_kernel(vb.VITERBI_AUTO)

# concurrent drop-in callers (README.md:56: QIRX >= 4.0 calls deconvolve from several threads): each thread has
# its own streams and staging buffers, so single-frame calls overlap on the device.  ctypes releases the GIL.
import threading  # noqa: E402

for f in (768, 3072):
    sym1, _ = dabgen.make_frames(1, f, 3.0, seed=2)
    for nthreads in (1, 2, 4, 8):
        calls = 200

        def worker():
            s32 = sym1[0].astype(np.uint32)
            o = np.zeros(f // 8, np.uint8)
            for _ in range(calls):
                assert vb.lib.deconvolve(f, s32.ctypes.data, 0, o.ctypes.data) == 0

        worker_threads = [threading.Thread(target=worker) for _ in range(nthreads)]
        t0 = time.perf_counter()
        for t in worker_threads:
            t.start()
        for t in worker_threads:
            t.join()
        dt = time.perf_counter() - t0
        print(json.dumps({"framebits": f, "threads": nthreads, "dropin_calls_per_s": round(nthreads * calls / dt, 1),
                          "us_per_call_per_thread": round(dt / calls * 1e6, 1)}), flush=True)
