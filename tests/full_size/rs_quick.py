"""RS kernel A/B: for the library named by VITERBI_B200_LIB check 8 x 4000 superframes (s = 1..8, 0-7 errors per
codeword) bit for bit against the CPU checker, then time the bench mix (8 x 125,000 superframes, device-resident)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import oracle_lib  # noqa: E402
import viterbi_dll_b200 as vb  # noqa: E402
from viterbi_dll_b200 import dabgen  # noqa: E402

assert vb.initialize()
chk = oracle_lib.checker()
per_s = 125000
sets, bad = [], 0
for s in range(1, 9):
    rx, _ = dabgen.make_superframes_torch(per_s, s, seed=900 + s, device="cuda")
    o = torch.full((per_s, 110 * s), 0xEE, dtype=torch.uint8, device="cuda")
    r = torch.empty((per_s,), dtype=torch.int32, device="cuda")
    sets.append((s, rx, o, r))
    vb.rs_check_superframe_batch_device(rx, s, o, r)
    torch.cuda.synchronize()
    want_o, want_r = chk.rs_batch(rx[:4000].cpu().numpy(), s, fill=0xEE)
    bad += int(not (np.array_equal(want_r, r[:4000].cpu().numpy()) and np.array_equal(want_o, o[:4000].cpu().numpy())))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    for s, rx, o, r in sets:
        vb.rs_check_superframe_batch_device(rx, s, o, r)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    for s, rx, o, r in sets:
        vb.rs_check_superframe_batch_device(rx, s, o, r)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("%s: %s, %.1f M superframes/s (%.3f ms per 10^6)" % (os.path.basename(os.environ.get("VITERBI_B200_LIB", "default")),
                                                          "PARITY OK" if bad == 0 else "PARITY FAILED (%d sets)" % bad, 8 * per_s / ms / 1e3, ms))
