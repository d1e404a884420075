"""End-to-end RS rate (pinned host buffers, 8 x 125,000 superframes, s = 1..8) for the chunk size given by
VITERBI_B200_RS_CHUNK_MB (read once per process)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import viterbi_dll_b200 as vb  # noqa: E402
from viterbi_dll_b200 import dabgen  # noqa: E402

assert vb.initialize()
per_s, host = 125000, []
for s in range(1, 9):
    rx, _ = dabgen.make_superframes_torch(per_s, s, seed=900 + s, device="cuda")
    h_rx = torch.empty(rx.shape, dtype=torch.uint8, pin_memory=True)
    h_rx.copy_(rx)
    host.append((s, h_rx, torch.full((per_s, 110 * s), 0xEE, dtype=torch.uint8, pin_memory=True),
                 torch.empty((per_s,), dtype=torch.int32, pin_memory=True)))
torch.cuda.synchronize()


def step():
    for s, h_rx, h_o, h_r in host:
        assert vb.lib.rs_check_superframe_batch(h_rx.data_ptr(), s, per_s, h_o.data_ptr(), h_r.data_ptr()) == 0


step()
t0 = time.perf_counter()
for _ in range(4):
    step()
dt = (time.perf_counter() - t0) / 4
print("RS chunk %s MB: %.1f M superframes/s end to end (%.2f ms per 10^6)" % (os.environ.get("VITERBI_B200_RS_CHUNK_MB", "32 (default)"), 8 * per_s / dt / 1e6, dt * 1e3))
