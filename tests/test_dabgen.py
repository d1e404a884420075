"""CPU tests of the synthetic-traffic generators (viterbi.dll_b200/dabgen.py) against the oracle: the inputs the
GPU parity tests and the benchmark feed to the decoders must themselves be right (the reference ships no
encoder for RS and only a Windows-only one for the convolutional code, viterbi-benchmark.cpp:304-311)."""
import numpy as np
import pytest

from viterbi_dll_b200 import dabgen


def test_convolutional_encoder_matches_known_answer(port):
    """KAT-V1 of SURVEY.md 8(c): message A5 3C 0F 81, noiseless symbols 0/255 -> the decoder returns the message."""
    msg = np.frombuffer(bytes.fromhex("A53C0F81"), dtype=np.uint8)
    bits = np.unpackbits(msg)[None, :]
    code = dabgen.conv_encode(bits)
    assert code.shape == (1, 4 * (32 + 6))
    assert np.array_equal(port.deconvolve_batch(32, (code * 255).astype(np.uint8)), msg[None, :])
    assert np.array_equal(port.deconvolve_batch(32, np.where(code, 200, 56).astype(np.uint8)), msg[None, :])


def test_rs_encoder_generator_and_roundtrip(port):
    g = dabgen.rs_generator_poly()
    assert bytes(g).hex().upper() == "C19D715F5EC76F9FC2D801"  # KAT-R0
    rng = np.random.default_rng(3)
    for s in (1, 4, 16):
        rx, payload, nerr = dabgen.make_superframes(50, s, seed=s, max_err=5)
        out, ret = port.rs_batch(rx, s, fill=0xEE)
        assert np.array_equal(out, payload)  # up to 5 errors per codeword are always corrected
        assert np.array_equal(ret, nerr.sum(axis=1))


def test_puncture_vectors_and_patterns():
    prev = np.zeros(32, np.uint8)
    for pi in range(1, 25):
        v = dabgen.puncture_vector(pi)
        assert v.sum() == 8 + pi and ((v - prev.astype(np.int16)) >= 0).all()  # each PI adds kept bits to PI - 1
        prev = v
    assert dabgen.puncture_vector(24).all() and dabgen.TAIL_VECTOR.sum() == 12
    keep = dabgen.fic_puncture_pattern()
    assert keep.size == 3096 and keep.sum() == 2304
    with pytest.raises(ValueError):
        dabgen.puncture_pattern(768, [(23, 16)])
    sym = np.arange(2 * 3096, dtype=np.uint32).reshape(2, 3096).astype(np.uint8)
    rx = dabgen.puncture(sym, keep)
    back = dabgen.depuncture(rx, keep, erasure=77)
    assert rx.shape == (2, 2304) and np.array_equal(back[:, keep == 1], rx) and (back[:, keep == 0] == 77).all()


def test_punctured_fic_still_decodes_on_a_clean_channel(port):
    sym, bits = dabgen.make_frames(20, 768, 12.0, seed=5)
    keep = dabgen.fic_puncture_pattern()
    out = port.deconvolve_batch(768, dabgen.depuncture(dabgen.puncture(sym, keep), keep))
    assert np.array_equal(out, bits)


def test_torch_generators_agree_with_the_decoders(port):
    """The device-side generators used by bench.py / tests/full_size/e2e_scaling.py, run here on the CPU."""
    sym, bits = dabgen.make_frames_torch(40, 96, 9.0, seed=1, device="cpu", want_bits=True)
    assert np.array_equal(port.deconvolve_batch(96, sym.numpy()), bits.numpy())
    rx, nerr = dabgen.make_superframes_torch(30, 3, seed=2, device="cpu", max_err=5)
    out, ret = port.rs_batch(rx.numpy(), 3, fill=0xEE)
    assert np.array_equal(ret, nerr.numpy().sum(axis=1))
    syms, payload = dabgen.make_superframe_frames_torch(6, 384, 9.0, seed=5, device="cpu", max_err=3)
    dec = port.deconvolve_batch(384, syms.numpy()).reshape(6, -1)
    out, ret = port.rs_batch(dec, 2, fill=0xEE)
    assert (ret >= 0).all() and np.array_equal(out, payload.numpy())


def test_energy_dispersal_prbs_known_prefix():
    """ETSI EN 300 401 clause 10: the PRBS of X^9 + X^5 + 1 from an all-ones register starts 0000 0111 1011 1110;
    period 511."""
    from viterbi_dll_b200 import dabgen

    p = dabgen.energy_dispersal_prbs(1100)
    assert "".join(map(str, p[:16])) == "0000011110111110"
    assert np.array_equal(p[:511], p[511:1022]) and not np.array_equal(p[:100], p[100:200])
