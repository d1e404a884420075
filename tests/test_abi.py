"""CPU tests of the drop-in boundary: libviterbi_b200.so loads without a GPU and exports every
symbol include/viterbi_b200.h declares; without a device every entry point fails loudly (there is
no CPU decode path in the product)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "viterbi_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^[A-Za-z_][\w\s\*]*?\b(\w+)\s*\([^;{]*\)\s*;", text, flags=re.M)
    return sorted(set(names))


def test_header_declares_reference_exports():
    names = declared_functions()
    # viterbi.def:4-8 plus the spelling used by BASELINE.json
    for n in ("deconvolve", "initialize", "RScheckSuperframe", "RSCheckSuperframe", "GetCPUCaps", "WakeUpYMM"):
        assert n in names
    for n in ("viterbi_deconvolve_batch", "viterbi_deconvolve_batch_u32", "viterbi_deconvolve_batch_device",
              "viterbi_deconvolve_batch_u32_device", "rs_check_superframe_batch", "rs_check_superframe_batch_device"):
        assert n in names


def test_library_exports_every_declared_symbol(vb):
    out = subprocess.run(["nm", "-D", "--defined-only", vb.LIB_PATH], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in declared_functions() if n not in exported]
    assert not missing, missing
    # nothing but the C ABI leaks out (built with -fvisibility=hidden)
    extra = [n for n in exported if n not in declared_functions()]
    assert not extra, extra


def test_binding_covers_header(vb):
    assert sorted(vb._SIGNATURES) == declared_functions()


def test_library_does_not_link_the_oracle(vb):
    out = subprocess.run(["ldd", vb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "viterbi_ref" not in out
    for src in os.listdir(os.path.join(ROOT, "viterbi.dll_b200", "csrc")):
        text = open(os.path.join(ROOT, "viterbi.dll_b200", "csrc", src)).read()
        assert "oracle/" not in text and "fec_oracle" not in text, src


def _no_gpu(vb):
    return vb.lib.fec_device_count() == 0


def test_fails_loudly_without_device(vb):
    if not _no_gpu(vb):
        pytest.skip("a CUDA device is present")
    sym = np.full(4 * (64 + 6), 128, dtype=np.uint32)
    rc, _ = vb.deconvolve(64, sym)
    assert rc == 1 and vb.lib.fec_in_save_mode() == 1  # exc_handler.cpp:214 convention
    assert b"CUDA" in vb.lib.fec_last_error() or b"cuda" in vb.lib.fec_last_error()
    out = np.zeros(110, dtype=np.uint8)
    assert vb.RScheckSuperframe(np.zeros(120, dtype=np.uint8), 0, 1, out) == -1
    with pytest.raises(vb.FecError):
        vb.deconvolve_batch(64, np.zeros((2, 280), dtype=np.uint8))
    with pytest.raises(vb.FecError):
        vb.rs_check_superframe_batch(np.zeros((2, 120), dtype=np.uint8), 1)
    assert vb.initialize() is False  # no device to select
    assert vb.lib.fec_in_save_mode() == 0  # initialize() clears the latch (dllmain.cpp:157)


def test_argument_validation(vb):
    z = np.zeros((1, 4 * 13), dtype=np.uint8)
    assert vb.lib.viterbi_deconvolve_batch(7, z.ctypes.data_as(ctypes.c_void_p), 1, z.ctypes.data_as(ctypes.c_void_p)) == vb.FEC_ERR_ARG
    assert vb.lib.viterbi_deconvolve_batch(9218, None, 1, None) == vb.FEC_ERR_ARG
    assert vb.lib.viterbi_deconvolve_batch(768, None, 0, None) == vb.FEC_OK  # empty batch is a no-op
    assert vb.lib.rs_check_superframe_batch(None, 0, 1, None, None) == vb.FEC_ERR_ARG
    assert vb.lib.rs_check_superframe_batch(None, 4, 0, None, None) == vb.FEC_OK
    # punctured input: the keep pattern must account for exactly rx_per_frame symbols (checked before any device use)
    keep = np.ones(4 * 14, dtype=np.uint8)
    kp, zp = keep.ctypes.data_as(ctypes.c_void_p), z.ctypes.data_as(ctypes.c_void_p)
    assert vb.lib.viterbi_deconvolve_batch_punctured(8, zp, 55, kp, 128, 1, zp) == vb.FEC_ERR_ARG
    assert vb.lib.viterbi_deconvolve_batch_punctured(8, zp, 56, kp, 300, 1, zp) == vb.FEC_ERR_ARG
    assert vb.lib.viterbi_deconvolve_batch_punctured(8, zp, 56, None, 128, 1, zp) == vb.FEC_ERR_ARG
    assert vb.lib.viterbi_deconvolve_batch_punctured(8, None, 56, kp, 128, 0, None) == vb.FEC_OK
    assert vb.lib.GetCPUCaps() == 0
    vb.lib.WakeUpYMM()


def test_call_log_is_env_gated(vb, tmp_path):
    """VITERBI_B200_LOG=<file> (read by initialize()) logs one line per API call, like the reference's
    VIT_WRITE_LOGFILE build (deconvolve.cpp:602-621); off by default.  Works without a device: failed calls
    are logged with their return value."""
    import os

    path = tmp_path / "calls.log"
    os.environ["VITERBI_B200_LOG"] = str(path)
    try:
        vb.initialize()
        vb.deconvolve(7, np.zeros(52, np.uint32))  # unsupported framebits: rc 1, no save mode
        out = np.zeros(110, dtype=np.uint8)
        vb.RScheckSuperframe(np.zeros(120, dtype=np.uint8), 0, 1, out)
    finally:
        del os.environ["VITERBI_B200_LOG"]
        vb.initialize()
    lines = path.read_text().splitlines()
    assert len(lines) == 2
    assert " deco:" in lines[0] and "rc: 1" in lines[0] and "shape:    7" in lines[0]
    assert " rssf:" in lines[1] and "dT:" in lines[1] and "TID:" in lines[1] and "ReE: 0" in lines[1]
    n = len(lines)
    vb.deconvolve(7, np.zeros(52, np.uint32))  # logging is off again
    assert len(path.read_text().splitlines()) == n


def test_bench_reference_arm_runs_the_reference_code_only(tmp_path):
    """`bench.py --impl reference` (the arm the driver times beside the GPU arm): prints the contract's JSON line with
    "impl": "reference", the same metric / config as the GPU arm, a cpu_baseline that says what ran -- and never maps
    the product library."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--frames", "2048", "--rs-superframes", "800"],
                         capture_output=True, text=True, env=dict(os.environ, LD_DEBUG="files"), timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "viterbi_decoded_gbit_per_s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert "libviterbi_ref" in out.stderr or "libfec_oracle" in out.stderr, "LD_DEBUG=files did not list the checker library"
    assert "libviterbi_b200.so" not in out.stderr, "the reference arm loaded the product library"


def test_native_hosts_compile_against_the_header(tmp_path):
    """The C++ hosts that only run on a GPU box (tests/host/multi_device_check.cpp, profiles/microbench/latbench.cpp)
    must at least keep compiling here, so that a change of the C ABI cannot break them unnoticed."""
    import shutil
    import subprocess

    if shutil.which("g++") is None:
        import pytest

        pytest.skip("g++ not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for src in (os.path.join(root, "tests", "host", "multi_device_check.cpp"),
                os.path.join(root, "profiles", "microbench", "latbench.cpp")):
        out = subprocess.run(["g++", "-std=c++17", "-O0", "-pthread", "-c", "-o", str(tmp_path / "x.o"), src],
                             capture_output=True, text=True)
        assert out.returncode == 0, src + "\n" + out.stderr[-2000:]
