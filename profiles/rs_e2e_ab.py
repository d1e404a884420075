"""RS end-to-end rate per RSDims, pinned buffers, through rs_check_superframe_batch -- run once per variant:
    python profiles/rs_e2e_ab.py                 (kernel reads the caller's outVector bytes of failing superframes itself)
    VITERBI_B200_RS_UPLOAD=1 python profiles/rs_e2e_ab.py     (outVector uploaded first, the round-1 path)"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import viterbi_dll_b200 as vb  # noqa: E402
from viterbi_dll_b200 import dabgen  # noqa: E402

assert vb.initialize()
per_s = 125000
res, tot_t = {}, 0.0
for max_err in (7, 3):
    tot_t = 0.0
    for s in range(1, 9):
        rx, _ = dabgen.make_superframes_torch(per_s, s, seed=900 + s, device="cuda", max_err=max_err)
        h_rx = torch.empty(rx.shape, dtype=torch.uint8, pin_memory=True)
        h_rx.copy_(rx)
        h_o = torch.full((per_s, 110 * s), 0xEE, dtype=torch.uint8, pin_memory=True)
        h_r = torch.empty((per_s,), dtype=torch.int32, pin_memory=True)
        torch.cuda.synchronize()
        for _ in range(2):
            assert vb.lib.rs_check_superframe_batch(h_rx.data_ptr(), s, per_s, h_o.data_ptr(), h_r.data_ptr()) == 0
        t0 = time.perf_counter()
        for _ in range(5):
            vb.lib.rs_check_superframe_batch(h_rx.data_ptr(), s, per_s, h_o.data_ptr(), h_r.data_ptr())
        dt = (time.perf_counter() - t0) / 5
        tot_t += dt
        res["err%d_s%d" % (max_err, s)] = {"ms": dt * 1e3, "M_sf_per_s": per_s / dt / 1e6, "failed_frac": float((h_r < 0).float().mean()),
                                          "in_out_GBps": per_s * 230 * s / dt / 1e9}
    res["err%d_mix_M_sf_per_s" % max_err] = 8 * per_s / tot_t / 1e6
res["upload_forced"] = os.environ.get("VITERBI_B200_RS_UPLOAD") == "1"
print(json.dumps(res))
