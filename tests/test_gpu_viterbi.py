"""GPU parity tests for the Viterbi path: libviterbi_b200.so (through its C ABI) against the
CPU checker -- the reference compiled unmodified (oracle/_ref) when it was built, else the C
port -- and against the golden vectors recorded from the reference.  Bit-exact or fail."""
import hashlib
import json
import os

import numpy as np
import pytest

from viterbi_dll_b200 import dabgen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device(vb):
    assert vb.lib.fec_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    assert vb.initialize()
    yield
    vb.set_viterbi_kernel(vb.VITERBI_AUTO)
    assert vb.lib.fec_in_save_mode() == 0


@pytest.fixture(params=["pair", "warp"])
def kernel(request, vb):
    """Run the test once per Viterbi kernel: the two-frames-per-thread throughput kernel and the
    warp-per-frame latency kernel (automatic selection would pick by batch size)."""
    vb.set_viterbi_kernel(vb.VITERBI_PAIR if request.param == "pair" else vb.VITERBI_WARP)
    yield request.param
    vb.set_viterbi_kernel(vb.VITERBI_AUTO)


def test_known_answers_through_dropin_deconvolve(vb, kernel, golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "kat.json")))
    for e in kat["viterbi"]:
        if e.get("symbols_hex"):
            sym = np.frombuffer(bytes.fromhex(e["symbols_hex"]), dtype="<u4")
        else:
            sym = dabgen.lcg_symbols(e["lcg_seed"], 4 * (e["framebits"] + 6)).astype(np.uint32)
        rc, out = vb.deconvolve(e["framebits"], sym)
        assert rc == 0
        assert hashlib.sha256(out.tobytes()).hexdigest() == e["out_sha256"], e["name"]


def test_golden_fixture_host_and_device(vb, kernel, golden_dir):
    import torch

    fx = np.load(os.path.join(golden_dir, "viterbi_fixture.npz"))
    for k in sorted(k[:-4] for k in fx.files if k.endswith("_sym")):
        f = int(k.split("_F")[1])
        sym, want = fx[k + "_sym"], fx[k + "_out"]
        assert np.array_equal(vb.deconvolve_batch(f, sym), want), k
        assert np.array_equal(vb.deconvolve_batch(f, sym.astype(np.uint32)), want), k + " (u32)"
        d = torch.from_numpy(sym).cuda()
        assert np.array_equal(vb.deconvolve_batch_device(f, d).cpu().numpy(), want), k + " (device)"
        d32 = torch.from_numpy(sym.astype(np.int32)).cuda()
        assert np.array_equal(vb.deconvolve_batch_device(f, d32).cpu().numpy(), want), k + " (device u32)"


@pytest.mark.parametrize("framebits,ebn0,n", [(768, 3.0, 4096), (768, 0.0, 1000), (3072, 0.0, 700), (3072, 3.0, 1500),
                                              (3072, 6.0, 700), (1536, 2.0, 513), (2304, 4.0, 257), (9216, 3.0, 130),
                                              (2, 1.0, 200), (10, 1.0, 65), (100, 2.0, 63), (770, 3.0, 129), (772, 3.0, 64)])
def test_random_frames_match_checker(vb, kernel, checker, framebits, ebn0, n):
    sym, _ = dabgen.make_frames(n, framebits, ebn0, seed=framebits * 7 + n)
    assert np.array_equal(vb.deconvolve_batch(framebits, sym), checker.deconvolve_batch(framebits, sym))


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 63, 64, 65, 127, 129, 1000])
def test_ragged_batch_sizes(vb, kernel, checker, n):
    sym, _ = dabgen.make_frames(n, 768, 2.0, seed=n)
    assert np.array_equal(vb.deconvolve_batch(768, sym), checker.deconvolve_batch(768, sym))


def test_adversarial_symbols(vb, kernel, checker):
    rng = np.random.default_rng(11)
    for f in (768, 3072):
        ns = 4 * (f + 6)
        rows = [np.zeros(ns, np.uint8), np.full(ns, 128, np.uint8), np.full(ns, 255, np.uint8), np.full(ns, 127, np.uint8)]
        rows += [rng.integers(0, 2, ns, dtype=np.uint8) * 255 for _ in range(60)]  # saturation / clamp heavy
        rows += [rng.integers(0, 256, ns, dtype=np.uint8) for _ in range(60)]
        rows += [rng.integers(120, 136, ns, dtype=np.uint8) for _ in range(30)]  # tie heavy
        rows += [np.tile(np.array([0, 255, 255, 0], np.uint8), f + 6), np.tile(np.array([255, 0, 0, 255], np.uint8), f + 6)]
        sym = np.stack(rows)
        assert np.array_equal(vb.deconvolve_batch(f, sym), checker.deconvolve_batch(f, sym)), f


def test_out_of_range_symbols_low_byte_only(vb, kernel, checker):
    """README.md:19 edge case: words above 255 -- only the low byte counts (deconvolve.cpp:219-228)."""
    rng = np.random.default_rng(12)
    sym, _ = dabgen.make_frames(70, 768, 3.0, seed=5)
    dirty = sym.astype(np.uint32) | (rng.integers(0, 1 << 24, size=sym.shape, dtype=np.uint32) << 8)
    want = checker.deconvolve_batch(768, sym)
    assert np.array_equal(vb.deconvolve_batch(768, dirty), want)
    rc, out = vb.deconvolve(768, dirty[3])
    assert rc == 0 and np.array_equal(out, want[3])


@pytest.mark.parametrize("framebits,n,ebn0", [(768, 700, 4.0), (768, 5000, 5.0), (3072, 300, 4.0), (96, 130, 6.0), (770, 1, 4.0),
                                              (768, 2, 4.0), (34, 65, 6.0)])
def test_punctured_input_equals_reference_on_expanded_symbols(vb, kernel, checker, framebits, n, ebn0):
    """Depuncturing front end: decoding the transmitted symbols + keep pattern must equal the
    reference decoder run on the host-expanded rate-1/4 stream (erasures = 128), host and device flavours."""
    import torch

    rng = np.random.default_rng(framebits + n)
    sym, _ = dabgen.make_frames(n, framebits, ebn0, seed=framebits + 3 * n)
    patterns = [dabgen.fic_puncture_pattern()] if framebits == 768 else []
    if framebits % 32 == 0:
        patterns.append(dabgen.puncture_pattern(framebits, [(framebits // 32, 8)]))      # rate 1/2 everywhere
    patterns.append((rng.random(4 * (framebits + 6)) < 0.6).astype(np.uint8))           # arbitrary pattern
    patterns.append(np.ones(4 * (framebits + 6), np.uint8))                             # nothing punctured
    sparse = (rng.random(4 * (framebits + 6)) < 0.12).astype(np.uint8)                  # whole steps without a symbol,
    sparse[-40:] = 0                                                                    # ... nothing at all in the tail
    sparse[5] = 1
    patterns.append(sparse)
    head = np.zeros(4 * (framebits + 6), np.uint8)
    head[:17] = 1                                                                       # 17 symbols, then erasures only
    patterns.append(head)
    for keep in patterns:
        rx = dabgen.puncture(sym, keep)
        for erasure in (128, 0):
            want = checker.deconvolve_batch(framebits, dabgen.depuncture(rx, keep, erasure))
            assert np.array_equal(vb.deconvolve_batch_punctured(framebits, rx, keep, erasure), want)
            got = vb.deconvolve_batch_punctured_device(framebits, torch.from_numpy(rx).cuda(), keep, erasure)
            assert np.array_equal(got.cpu().numpy(), want)


def test_punctured_fic_recovers_payload(vb):
    """FIC-shaped puncturing (3096 -> 2304 symbols) at a comfortable Eb/N0: the payload comes back."""
    sym, bits = dabgen.make_frames(4096, 768, 7.0, seed=9)
    keep = dabgen.fic_puncture_pattern()
    assert keep.sum() == 2304
    out = vb.deconvolve_batch_punctured(768, dabgen.puncture(sym, keep), keep)
    assert (out != bits).any(axis=1).mean() < 0.01


def test_punctured_argument_checks(vb):
    keep = dabgen.fic_puncture_pattern()
    rx = np.zeros((3, 2304), np.uint8)
    with pytest.raises(vb.FecError):
        vb.deconvolve_batch_punctured(768, rx[:, :-1], keep)  # pattern keeps 2304, rows hold 2303
    with pytest.raises(vb.FecError):
        vb.deconvolve_batch_punctured(768, rx, keep, erasure=256)
    assert vb.lib.fec_in_save_mode() == 0


def test_empty_and_zero_length(vb):
    assert vb.deconvolve_batch(768, np.zeros((0, 3096), np.uint8)).shape == (0, 96)
    assert vb.deconvolve_batch(0, np.zeros((5, 24), np.uint8)).shape == (5, 0)  # F = 0: nothing to write


def test_ber_fer_curve_identical_to_checker(vb, kernel, checker):
    """Eb/N0 sweep 0..6 dB (BASELINE config 3, reduced count): same decoded bits => same BER/FER."""
    for eb in range(0, 7):
        sym, bits = dabgen.make_frames(384, 3072, float(eb), seed=1000 + eb)
        got, want = vb.deconvolve_batch(3072, sym), checker.deconvolve_batch(3072, sym)
        assert np.array_equal(got, want), eb
        err = np.unpackbits(got ^ bits, axis=1).sum(axis=1)
        ber, fer = err.sum() / err.size / 3072, (err > 0).mean()
        if eb == 0:
            assert ber > 1e-2
        if eb == 6:
            assert ber < 1e-5 and fer < 0.02


def test_automatic_kernel_selection_agrees(vb, checker):
    """Below 4,096 frames per launch the library picks the warp kernel, above it the pair kernel."""
    vb.set_viterbi_kernel(vb.VITERBI_AUTO)
    for n in (100, 4095, 4096, 6000):
        sym, _ = dabgen.make_frames(n, 96, 2.0, seed=n)
        l0 = vb.kernel_launches()
        got = vb.deconvolve_batch(96, sym)
        assert vb.kernel_launches() > l0
        assert np.array_equal(got, checker.deconvolve_batch(96, sym)), n


def test_full_size_fic_roundtrip_on_device(vb):
    """BASELINE config 2 size (65,536 FIC blocks): encode -> noiseless channel -> decode == payload,
    and mild noise is fully corrected.  Size-independent property, no CPU reference needed."""
    import torch

    n, f = 65536, 768
    sym, bits = dabgen.make_frames_torch(n, f, 9.0, seed=3, device="cuda", want_bits=True)
    out = vb.deconvolve_batch_device(f, sym)
    torch.cuda.synchronize()
    assert torch.equal(out, bits)


def test_one_million_fic_frames_bit_exact(vb, checker):
    """>= 10^6 random frames, bit-exact against the checker (north-star acceptance)."""
    import torch

    f, total, chunk = 768, 1 << 20, 1 << 17
    if checker.kind != "reference":
        total = 1 << 16  # the scalar port is slower; keep the CPU side bounded
    bad = 0
    for i in range(total // chunk if total >= chunk else 1):
        m = min(chunk, total)
        sym, _ = dabgen.make_frames_torch(m, f, 1.0 + (i % 5), seed=77 + i, device="cuda")
        out = vb.deconvolve_batch_device(f, sym)
        torch.cuda.synchronize()
        want = checker.deconvolve_batch(f, sym.cpu().numpy())
        bad += int((out.cpu().numpy() != want).any(axis=1).sum())
    assert bad == 0


def test_save_mode_latch(vb):
    """exc_handler.cpp:204-214 convention: a fault makes deconvolve return 1 until initialize()."""
    import ctypes

    assert vb.lib.deconvolve(64, None, 0, None) == 1
    assert vb.lib.fec_in_save_mode() == 1
    rc, _ = vb.deconvolve(64, np.full(280, 128, np.uint32))
    assert rc == 1
    assert vb.initialize()
    rc, out = vb.deconvolve(64, np.full(280, 128, np.uint32))
    assert rc == 0 and out.tobytes().hex() == "fc0fc0fc0fc0fc3f"
    # an unsupported frame length is rejected without latching
    rc, _ = vb.deconvolve(7, np.zeros(52, np.uint32))
    assert rc == 1 and vb.lib.fec_in_save_mode() == 0


def test_concurrent_callers(vb, checker):
    """QIRX >= 4.0 calls deconvolve from several threads at once (README.md:56): concurrent drop-in and
    batched calls must not disturb one another."""
    import threading

    vb.set_viterbi_kernel(vb.VITERBI_AUTO)
    frames = {}
    for tid in range(6):
        f = (768, 3072, 1536)[tid % 3]
        sym, _ = dabgen.make_frames(40 + tid, f, 2.5, seed=500 + tid)
        frames[tid] = (f, sym, checker.deconvolve_batch(f, sym))
    errors = []

    def worker(tid):
        f, sym, want = frames[tid]
        try:
            for rep in range(4):
                if tid % 2:
                    got = vb.deconvolve_batch(f, sym)
                    if not np.array_equal(got, want):
                        errors.append((tid, rep, "batch"))
                else:
                    for i in range(0, sym.shape[0], 7):
                        rc, out = vb.deconvolve(f, sym[i].astype(np.uint32))
                        if rc != 0 or not np.array_equal(out, want[i]):
                            errors.append((tid, rep, i))
        except Exception as e:  # pragma: no cover
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in frames]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]


def test_concurrent_callers_with_large_frames(vb, checker):
    """Two threads decoding F = 9216 (111 KB of shared memory per warp) and F = 6144 (74 KB) at once, on the
    warp-per-frame kernel: the shared-memory opt-in is a per-device function attribute set once for the worst
    case, so neither caller can lower the other's limit (a per-thread cache of it used to)."""
    import threading

    vb.set_viterbi_kernel(vb.VITERBI_AUTO)
    jobs = {}
    for tid, f in enumerate((9216, 6144, 9216, 4092)):
        sym, _ = dabgen.make_frames(6, f, 3.0, seed=900 + tid)
        jobs[tid] = (f, sym, checker.deconvolve_batch(f, sym))
    errors = []

    def worker(tid):
        f, sym, want = jobs[tid]
        try:
            for rep in range(6):
                if rep % 2:
                    if not np.array_equal(vb.deconvolve_batch(f, sym), want):
                        errors.append((tid, rep, "batch"))
                else:
                    rc, out = vb.deconvolve(f, sym[rep % 6].astype(np.uint32))
                    if rc != 0 or not np.array_equal(out, want[rep % 6]):
                        errors.append((tid, rep, "dropin", rc))
        except Exception as e:  # pragma: no cover
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]
    assert vb.lib.fec_in_save_mode() == 0


def test_unaligned_device_output(vb, checker):
    """d_out needs no alignment: a misaligned output falls back to byte stores (F % 32 == 0 would otherwise use
    32-bit stores)."""
    import torch

    f, n = 768, 5000
    sym, _ = dabgen.make_frames(n, f, 3.0, seed=77)
    want = checker.deconvolve_batch(f, sym)
    d_sym = torch.from_numpy(sym).cuda()
    buf = torch.zeros(n * (f // 8) + 8, dtype=torch.uint8, device="cuda")
    for off in (1, 2, 3):
        out = buf[off : off + n * (f // 8)].view(n, f // 8)
        vb.set_viterbi_kernel(vb.VITERBI_PAIR)
        rc = vb.lib.viterbi_deconvolve_batch_device(f, d_sym.data_ptr(), n, out.data_ptr(), None)
        vb.set_viterbi_kernel(vb.VITERBI_AUTO)
        torch.cuda.synchronize()
        assert rc == 0 and np.array_equal(out.cpu().numpy(), want), off


def test_initialize_probes_and_keeps_working(vb, checker):
    """initialize() on a healthy device: context probed, staging state kept, decode still exact."""
    sym, _ = dabgen.make_frames(300, 768, 3.0, seed=5)
    want = checker.deconvolve_batch(768, sym)
    for _ in range(3):
        assert vb.initialize()
        assert np.array_equal(vb.deconvolve_batch(768, sym), want)
        rc, out = vb.deconvolve(768, sym[3].astype(np.uint32))
        assert rc == 0 and np.array_equal(out, want[3])
