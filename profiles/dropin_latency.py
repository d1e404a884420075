"""Latency of the single-frame drop-in deconvolve() and its throughput from 8 concurrent threads."""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import viterbi_dll_b200 as vb  # noqa: E402
from viterbi_dll_b200 import dabgen  # noqa: E402

assert vb.initialize()
for f in (768, 3072):
    sym1, _ = dabgen.make_frames(1, f, 3.0, seed=2)
    s32 = sym1[0].astype(np.uint32)
    o = np.zeros(f // 8, np.uint8)
    for _ in range(20):
        vb.lib.deconvolve(f, s32.ctypes.data, 0, o.ctypes.data)
    t0 = time.perf_counter()
    for _ in range(300):
        assert vb.lib.deconvolve(f, s32.ctypes.data, 0, o.ctypes.data) == 0
    one = (time.perf_counter() - t0) / 300 * 1e6

    def worker():
        a = sym1[0].astype(np.uint32)
        b = np.zeros(f // 8, np.uint8)
        for _ in range(300):
            assert vb.lib.deconvolve(f, a.ctypes.data, 0, b.ctypes.data) == 0

    ts = [threading.Thread(target=worker) for _ in range(8)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    print(json.dumps({"framebits": f, "dropin_deconvolve_us": round(one, 1), "calls_per_s_8_threads": round(8 * 300 / dt, 1)}))
