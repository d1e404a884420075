"""CPU tests: the oracle (oracle/fec_oracle.c) against the golden vectors recorded from the
reference itself (tests/golden, made by make_golden.py from oracle/_ref) and, when the compiled
reference is present, against it directly.  No GPU involved."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib
from viterbi_dll_b200 import dabgen


@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as f:
        return json.load(f)


def _kat_symbols(e):
    if e.get("symbols_hex"):
        return np.frombuffer(bytes.fromhex(e["symbols_hex"]), dtype="<u4")
    return dabgen.lcg_symbols(e["lcg_seed"], 4 * (e["framebits"] + 6)).astype(np.uint32)


def test_viterbi_known_answers(kat, port):
    assert len(kat["viterbi"]) >= 10
    for e in kat["viterbi"]:
        out = port.deconvolve(e["framebits"], _kat_symbols(e))
        assert hashlib.sha256(out.tobytes()).hexdigest() == e["out_sha256"], e["name"]
        if e.get("out_hex"):
            assert out.tobytes().hex() == e["out_hex"], e["name"]


def test_rs_known_answers(kat, port):
    names = [e["name"] for e in kat["rs"]]
    assert names == ["R0", "R1", "R2", "R3", "R3b", "R4", "R4b"]
    for e in kat["rs"]:
        p = np.frombuffer(bytes.fromhex(e["in_hex"]), dtype=np.uint8)
        out = np.full(110 * e["s"], e["out_prefill"], dtype=np.uint8)
        ret = port.rs_check_superframe(p, e["s"], out)
        assert ret == e["ret"], e["name"]
        assert out.tobytes().hex() == e["out_hex"], e["name"]


def test_rs_generator_and_tables(kat, port):
    assert dabgen.rs_generator_poly().tobytes().hex() == kat["rs_generator_low_to_high_hex"]
    cw = dabgen.rs_encode(np.arange(1, 111, dtype=np.uint8)[None, :])[0]
    assert cw.tobytes().hex() == kat["rs_codeword_hex"]
    ato, iof = port.rs_tables()
    assert iof[0] == 255 and ato[255] == 1 and ato[0] == 1 and ato[1] == 2 and ato[8] == 0x1D
    assert all(ato[iof[v]] == v for v in range(1, 256))
    assert np.array_equal(ato[:255], ato[255:510]) and np.array_equal(ato[:255], ato[510:765])


def test_viterbi_fixture(golden_dir, port):
    fx = np.load(os.path.join(golden_dir, "viterbi_fixture.npz"))
    keys = sorted(k[:-4] for k in fx.files if k.endswith("_sym"))
    assert len(keys) >= 8
    for k in keys:
        f = int(k.split("_F")[1])
        out = port.deconvolve_batch(f, fx[k + "_sym"])
        assert np.array_equal(out, fx[k + "_out"]), k


def test_rs_fixture(golden_dir, port):
    fx = np.load(os.path.join(golden_dir, "rs_fixture.npz"))
    for s in (1, 2, 3, 4, 5, 6, 7, 8, 16, 24):
        out, ret = port.rs_batch(fx["s%d_in" % s], s, fill=0xEE)
        assert np.array_equal(ret, fx["s%d_ret" % s]), s
        assert np.array_equal(out, fx["s%d_out" % s]), s


def test_rs_semantics_on_fixture(golden_dir):
    """What the golden outputs say about the algorithm: <=5 errors are always corrected, the
    first failing column freezes the rest of the superframe."""
    fx = np.load(os.path.join(golden_dir, "rs_fixture.npz"))
    for s in (1, 4, 8):
        ret, nerr, out = fx["s%d_ret" % s], fx["s%d_nerr" % s], fx["s%d_out" % s]
        ok = (nerr <= 5).all(axis=1)
        assert (ret[ok] == nerr[ok].sum(axis=1)).all()
        failed = ret == -1
        assert failed.any()
        # last column of a failed superframe is never written (prefill 0xEE survives)
        assert (out[failed].reshape(failed.sum(), 110, s)[:, :, s - 1] == 0xEE).all()


def test_port_rejects_undefined_framebits(port):
    s = np.zeros(4 * (9218 + 6), dtype=np.uint32)
    out = np.zeros(2000, dtype=np.uint8)
    from ctypes import c_void_p

    assert port.lib.oracle_deconvolve(7, s.ctypes.data_as(c_void_p), 0, out.ctypes.data_as(c_void_p)) == -2
    assert port.lib.oracle_deconvolve(9218, s.ctypes.data_as(c_void_p), 0, out.ctypes.data_as(c_void_p)) == -2


def test_noiseless_roundtrip_and_ber(port):
    syms, bits = dabgen.make_frames(64, 768, 30.0, seed=1)
    assert np.array_equal(port.deconvolve_batch(768, syms), bits)
    syms, bits = dabgen.make_frames(400, 768, 3.0, seed=2)
    ber = np.unpackbits(port.deconvolve_batch(768, syms) ^ bits).mean()
    assert 1e-5 < ber < 2e-3  # survey probe: 1.65e-4 at 3 dB


@pytest.mark.skipif(oracle_lib.ref() is None, reason="oracle/_ref not built (needs /root/reference at build time)")
class TestPortAgainstCompiledReference:
    def test_viterbi_random(self, port):
        ref = oracle_lib.ref()
        rng = np.random.default_rng(5)
        for f, eb, n in ((768, 3.0, 300), (3072, 0.0, 100), (3072, 6.0, 100), (1536, 2.0, 100), (2304, 4.0, 60), (10, 1.0, 50)):
            syms, _ = dabgen.make_frames(n, f, eb, int(rng.integers(1 << 30)))
            assert np.array_equal(port.deconvolve_batch(f, syms), ref.deconvolve_batch(f, syms)), (f, eb)
        for f in (768, 3072):
            syms = rng.integers(0, 256, size=(100, 4 * (f + 6)), dtype=np.uint8)
            assert np.array_equal(port.deconvolve_batch(f, syms), ref.deconvolve_batch(f, syms))
            syms = rng.integers(0, 2, size=(100, 4 * (f + 6)), dtype=np.uint8) * 255
            assert np.array_equal(port.deconvolve_batch(f, syms), ref.deconvolve_batch(f, syms))

    def test_all_reference_flavours_agree(self):
        ref = oracle_lib.ref()
        syms, _ = dabgen.make_frames(40, 3072, 2.0, 99)
        outs = []
        for which in range(5):  # sse2_lut32, ssse3, avx, avx2, avx5
            ref.select(which)
            outs.append(ref.deconvolve_batch(3072, syms))
        ref.select(4 if ref.isa == "avx512" else 3)
        for o in outs[1:]:
            assert np.array_equal(o, outs[0])

    def test_upper_bytes_ignored(self):
        ref = oracle_lib.ref()
        rng = np.random.default_rng(3)
        syms, _ = dabgen.make_frames(1, 768, 3.0, 4)
        w = syms[0].astype(np.uint32)
        dirty = w | (rng.integers(0, 1 << 24, size=w.shape, dtype=np.uint32) << 8)
        assert np.array_equal(ref.deconvolve(768, w), ref.deconvolve(768, dirty))
        assert np.array_equal(oracle_lib.port().deconvolve(768, dirty), ref.deconvolve(768, w))

    def test_rs_random(self, port):
        ref = oracle_lib.ref()
        for s in (1, 2, 3, 5, 8, 13, 24):
            rx, _, _ = dabgen.make_superframes(400, s, 300 + s)
            oa, ra = port.rs_batch(rx, s)
            ob, rb = ref.rs_batch(rx, s)
            assert np.array_equal(ra, rb) and np.array_equal(oa, ob), s
        # heavy damage (random bytes): exercises deg(lambda) up to 10 and the -1 paths
        rng = np.random.default_rng(8)
        rx = rng.integers(0, 256, size=(3000, 120), dtype=np.uint8)
        oa, ra = port.rs_batch(rx, 1)
        ob, rb = ref.rs_batch(rx, 1)
        assert np.array_equal(ra, rb) and np.array_equal(oa, ob)
