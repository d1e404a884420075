"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / initcheck are 10-100x slower than
native, so the sizes are tiny but cover: pair kernel with full and ragged groups, F % 32 != 0, the warp kernel,
u32 compaction, depuncturing, RS with s = 1..17 and failing columns, the DAB+ chain).  Outputs are compared with
the CPU checker as well, so a sanitizer-clean run is also a correct one.

    compute-sanitizer --tool memcheck  python tests/full_size/kernels_smoke.py
    compute-sanitizer --tool racecheck python tests/full_size/kernels_smoke.py

(compute-sanitizer is closed on the round-1 GPU pool, so only the native run was done there: it passes.)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import oracle_lib  # noqa: E402
import viterbi_dll_b200 as vb  # noqa: E402
from viterbi_dll_b200 import dabgen  # noqa: E402

assert vb.initialize()
chk = oracle_lib.checker()
for mode in (vb.VITERBI_PAIR, vb.VITERBI_WARP):
    vb.set_viterbi_kernel(mode)
    for f, n in ((768, 130), (3072, 65), (100, 70), (34, 3), (2, 5), (9216, 2)):
        if mode == vb.VITERBI_WARP and n > 20:
            n = 20
        sym, _ = dabgen.make_frames(n, f, 3.0, seed=f + n)
        want = chk.deconvolve_batch(f, sym)
        got = vb.deconvolve_batch_device(f, torch.from_numpy(sym).cuda())
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy(), want), (mode, f, n)
        assert np.array_equal(vb.deconvolve_batch(f, sym.astype(np.uint32)), want)
vb.set_viterbi_kernel(vb.VITERBI_AUTO)
sym, _ = dabgen.make_frames(70, 768, 5.0, seed=1)
keep = dabgen.fic_puncture_pattern()
rx = dabgen.puncture(sym, keep)
assert np.array_equal(vb.deconvolve_batch_punctured(768, rx, keep), chk.deconvolve_batch(768, dabgen.depuncture(rx, keep)))
rc, one = vb.deconvolve(768, sym[3].astype(np.uint32))
assert rc == 0
for s in (1, 2, 3, 5, 8, 16, 17):
    rx, _, _ = dabgen.make_superframes(70, s, seed=s)
    want_out, want_ret = chk.rs_batch(rx, s, fill=0xEE)
    out, ret = vb.rs_check_superframe_batch(rx, s, fill=0xEE)
    assert np.array_equal(ret, want_ret) and np.array_equal(out, want_out), s
syms, payload, _ = dabgen.make_superframe_frames(6, 384, 6.0, seed=4, max_err=3)
out, ret = vb.dabplus_decode_superframes(384, syms, fill=0xEE)
assert (ret >= 0).all() and np.array_equal(out, payload)
print("sanitizer smoke ok, kernels launched:", vb.kernel_launches())
