"""CPU test of the built device code: the instruction budget of the two Viterbi kernels, read from the SASS of
the library the tests load.  The throughput kernel depends on ptxas turning the inline-PTX `min.u16x2 + setp`
pattern into ONE VIMNMX.U16x2 with two predicate outputs and on predicated adds for the decision bits
(viterbi_pair_core.h: min_decide); a toolkit that stops doing that still produces correct code, 15 % slower, and
nothing else would notice.  (Needs cuobjdump, which ships with nvcc.)"""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "profiles"))


@pytest.fixture(scope="module")
def sass(vb):  # vb: importing the binding (re)builds the library when stale
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    import sass_report

    return sass_report.summary(sass_report.functions(vb.LIB_PATH))[0]


def test_pair_kernel_acs_loop_budget(sass):
    k = sass["viterbi_pair_kernel"]
    h = k["histogram"]
    # 32 butterflies x 2 outputs x 2 steps: one saturating add-min and one packed min each
    assert h["VIADDMNMX.U16x2"] == 128 and h["VIMNMX.U16x2"] == 128
    assert k["vimnmx_u16x2_with_two_predicate_outputs"] == 128  # the fused min + two predicates
    # decision bits: one predicated add per frame and new state (a few become predicated IMADs)
    assert h.get("@P VIADD", 0) + h.get("@P IMAD.IADD.U32", 0) + h.get("@P IADD3", 0) >= 240  # 256 less the first bit of each word, which is a select
    # the un-fused fallback shows up as compares and byte gathers
    assert h.get("ISETP.NE.U32.AND", 0) + h.get("ISETP.EQ.U32.AND", 0) + h.get("ISETP.NE.AND", 0) < 8
    assert h.get("PRMT", 0) <= 16
    assert h["VIADDMNMX.S16x2.RELU"] == 64  # the folded renormalisation, even step only
    assert k["loop_instructions"] <= 850, k["loop_instructions"]  # 839 when this was written
    assert h.get("STL", 0) + h.get("STL.64", 0) + h.get("STL.128", 0) == 0  # no spill stores inside the loop


def test_warp_kernel_forward_loop_budget(sass):
    k = sass["viterbi_warp_kernel"]
    h = k["histogram"]
    assert h["SHFL.BFLY"] == 10  # ONE exchange per trellis step
    assert h["VIADDMNMX.U16x2"] == 10 and h["VIMNMX.U16x2"] == 10  # packed butterfly: both outputs in one instruction pair
    assert h["VOTE.ANY"] == 20 and h.get("STS.64", 0) == 10  # two ballots per step, one 8-byte decision store
    assert h.get("SHFL.IDX", 0) == 5  # renormalisation test on odd steps only
    assert k["loop_instructions"] <= 260, k["loop_instructions"]  # 238 when this was written
