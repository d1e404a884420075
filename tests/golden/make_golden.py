#!/usr/bin/env python
"""Regenerate tests/golden/* from the reference itself.

Runs ONLY where /root/reference exists (the build container): it drives
oracle/_ref -- the reference's deconvolve.cpp / rschecksf.cpp compiled
unmodified -- and records inputs + outputs.  The reference repository holds no
golden vectors of its own, so these are the pins:

  kat.json              known-answer vectors V1-V4 / R0-R4b of SURVEY.md section 8(c)
  viterbi_fixture.npz   noisy + adversarial frames (u8 symbols) and the decoded bytes
  rs_fixture.npz        superframes s=1..8 (+16, 24) with 0..7 errors/codeword, outputs, return values

Usage:  python tests/golden/make_golden.py
"""
import hashlib
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402

spec = importlib.util.spec_from_file_location("dabgen", os.path.join(ROOT, "viterbi.dll_b200", "dabgen.py"))
dabgen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(dabgen)


def main():
    oracle_lib.build()
    ref = oracle_lib.ref()
    assert ref is not None, "oracle/_ref not built (needs /root/reference)"
    kat = {"source": "oracle/_ref (reference compiled unmodified, g++ -O3, isa %s)" % ref.isa, "viterbi": [], "rs": []}

    def vit(name, framebits, syms32, note=""):
        out = ref.deconvolve(framebits, syms32)
        kat["viterbi"].append({"name": name, "framebits": framebits, "note": note,
                               "symbols_hex": np.asarray(syms32, dtype=np.uint32).astype("<u4").tobytes().hex()
                               if len(syms32) <= 1200 else None,
                               "out_hex": out.tobytes().hex() if framebits <= 64 else None,
                               "out_sha256": hashlib.sha256(out.tobytes()).hexdigest()})
        return out

    msg = np.unpackbits(np.frombuffer(bytes.fromhex("A53C0F81"), dtype=np.uint8))[None, :]
    code = dabgen.conv_encode(msg)[0].astype(np.uint32)
    assert vit("V1", 32, code * 255, "noiseless 0/255").tobytes().hex() == "a53c0f81"
    assert vit("V1b", 32, code * 144 + 56, "noiseless 56/200").tobytes().hex() == "a53c0f81"
    assert vit("V1c", 32, (code * 255) | 0xABCDEF00, "garbage in the upper 24 bits").tobytes().hex() == "a53c0f81"
    assert vit("V2", 64, np.full(280, 128, np.uint32)).tobytes().hex() == "fc0fc0fc0fc0fc3f"
    assert vit("V2b", 64, np.zeros(280, np.uint32)).tobytes().hex() == "00" * 8
    assert vit("V2c", 64, np.full(280, 255, np.uint32)).tobytes().hex() == "de606f1d93b276f3"
    assert vit("V3", 64, dabgen.lcg_symbols(12345, 280).astype(np.uint32)).tobytes().hex() == "4d1166302e40cccb"
    for f in (768, 3072, 9216):
        e = {"name": "V4_%d" % f, "framebits": f, "lcg_seed": 2024, "note": "symbols = dabgen.lcg_symbols(2024, 4*(F+6))",
             "out_sha256": hashlib.sha256(ref.deconvolve(f, dabgen.lcg_symbols(2024, 4 * (f + 6)).astype(np.uint32)).tobytes()).hexdigest()}
        kat["viterbi"].append(e)

    # --- RS known answers -------------------------------------------------
    gen = dabgen.rs_generator_poly()
    assert gen.tobytes().hex() == "c19d715f5ec76f9fc2d801"
    m = np.arange(1, 111, dtype=np.uint8)[None, :]
    cw = dabgen.rs_encode(m)[0]
    assert cw[110:].tobytes().hex() == "4f5bfa4fd93095e62f7b"
    kat["rs_generator_low_to_high_hex"] = gen.tobytes().hex()
    kat["rs_codeword_hex"] = cw.tobytes().hex()

    def rs(name, cols, note=""):
        """cols: list of per-column XOR patterns [(value, position), ...] applied to the codeword."""
        s = len(cols)
        rx = np.empty((s, 120), dtype=np.uint8)
        for j, pat in enumerate(cols):
            rx[j] = cw
            for v, pos in pat:
                rx[j, pos] ^= v
        p = dabgen.rs_interleave(rx, s)[0]
        out = np.full(110 * s, 0xEE, dtype=np.uint8)
        ret = ref.rs_check_superframe(p, s, out)
        kat["rs"].append({"name": name, "s": s, "note": note, "patterns": cols, "in_hex": p.tobytes().hex(),
                          "ret": int(ret), "out_hex": out.tobytes().hex(), "out_prefill": 0xEE})
        return ret, out.reshape(110, s)

    r, o = rs("R0", [[]], "clean codeword")
    assert r == 0 and np.array_equal(o[:, 0], m[0])
    r, o = rs("R1", [[(0x01, 0), (0x80, 57), (0xFF, 119)]])
    assert r == 3 and np.array_equal(o[:, 0], m[0])
    r2 = [(0x11, 3), (0x22, 20), (0x33, 41), (0x44, 66), (0x55, 90), (0x66, 118)]
    r, o = rs("R2", [r2], "6 errors: uncorrectable, output untouched")
    assert r == -1 and (o == 0xEE).all()
    r, o = rs("R3", [[(0xB9, 8), (0x42, 18), (0x98, 19), (0xAE, 58), (0x4A, 61), (0x58, 103)]],
              "ret 5 with roots in the virtual padding: only two bytes changed")
    assert r == 5 and not np.array_equal(o[:, 0], m[0])
    r, o = rs("R3b", [[(0x45, 27), (0x8B, 57), (0x70, 70), (0xFF, 85), (0x6F, 95), (0x4C, 112), (0x8C, 116)]],
              "7 errors: silent miscorrection")
    assert r == 5
    two = [(0xAA, 5), (0x01, 100)]
    r, o = rs("R4", [two, r2], "col0 corrected and written, col1 fails -> untouched")
    assert r == -1 and np.array_equal(o[:, 0], m[0]) and (o[:, 1] == 0xEE).all()
    r, o = rs("R4b", [r2, two], "col0 fails first -> nothing written")
    assert r == -1 and (o == 0xEE).all()

    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    # --- bulk fixtures ----------------------------------------------------
    rng = np.random.default_rng(20261018)
    vf = {}
    for tag, f, n, eb in (("fic", 768, 24, 3.0), ("msc_lo", 3072, 6, 0.0), ("msc", 3072, 6, 3.0), ("msc_hi", 3072, 4, 6.0),
                          ("tiny", 2, 16, 1.0), ("odd8", 100, 16, 2.0), ("max", 9216, 2, 3.0)):
        syms, _ = dabgen.make_frames(n, f, eb, int(rng.integers(1 << 31)))
        vf["%s_F%d_sym" % (tag, f)] = syms
        vf["%s_F%d_out" % (tag, f)] = ref.deconvolve_batch(f, syms)
    adv = np.stack([np.zeros(3096, np.uint8), np.full(3096, 128, np.uint8), np.full(3096, 255, np.uint8),
                    rng.integers(0, 2, 3096, dtype=np.uint8) * 255, rng.integers(0, 256, 3096, dtype=np.uint8),
                    rng.integers(120, 136, 3096, dtype=np.uint8)])
    vf["adv_F768_sym"] = adv
    vf["adv_F768_out"] = ref.deconvolve_batch(768, adv)
    np.savez_compressed(os.path.join(HERE, "viterbi_fixture.npz"), **vf)

    rf = {}
    for s in (1, 2, 3, 4, 5, 6, 7, 8, 16, 24):
        rx, _, nerr = dabgen.make_superframes(48 if s <= 8 else 12, s, 7000 + s)
        out, ret = ref.rs_batch(rx, s)
        rf["s%d_in" % s], rf["s%d_out" % s], rf["s%d_ret" % s], rf["s%d_nerr" % s] = rx, out, ret, nerr.astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "rs_fixture.npz"), **rf)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
