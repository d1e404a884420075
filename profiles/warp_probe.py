"""A short program for ncu: the warp-per-frame kernel on one frame (the drop-in shape) and on a 2,048-frame batch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
import viterbi_dll_b200 as vb  # noqa: E402
from viterbi_dll_b200 import dabgen  # noqa: E402

assert vb.initialize()
chk = oracle_lib.checker()
vb.set_viterbi_kernel(vb.VITERBI_WARP)
for f, n in ((3072, 1), (768, 2048)):
    sym, _ = dabgen.make_frames(n, f, 3.0, seed=3)
    d = torch.from_numpy(sym).cuda()
    for _ in range(3):
        out = vb.deconvolve_batch_device(f, d)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), chk.deconvolve_batch(f, sym)), (f, n)
print("ok")
