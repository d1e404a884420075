"""CPU tests of the device code.  The per-codeword decoder of the kernel (csrc/rs_decode.h) and its bit-sliced
Chien search (csrc/rs_chien_bitsliced.h) are written against a small policy class / as host-device templates, so
the very code the kernel runs is built here with g++ and checked: the Chien search against a plain Chien search
(rschecksf.cpp:296-320 restated) on random and fully splitting locator polynomials of every degree 1..10, the whole
decoder against the oracle on encoded codewords with 0..8 errors and on garbage.  The arithmetic of the Viterbi
throughput kernel (csrc/viterbi_pair_core.h) is checked the same way, with its packed DPX instructions emulated."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_bitsliced_chien_equals_plain_chien(tmp_path):
    exe = tmp_path / "chien_check"
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", str(exe), os.path.join(ROOT, "tests", "host", "chien_check.cpp")],
                   check=True)
    out = subprocess.run([str(exe), "4000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok:")


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_kernel_rs_decoder_equals_oracle_on_the_host(tmp_path):
    import oracle_lib

    oracle_lib.port()  # builds oracle/libfec_oracle.so if needed
    exe = tmp_path / "rs_decode_check"
    oracle_dir = os.path.join(ROOT, "oracle")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", str(exe), os.path.join(ROOT, "tests", "host", "rs_decode_check.cpp"),
                    "-L" + oracle_dir, "-lfec_oracle", "-Wl,-rpath," + oracle_dir], check=True, stderr=subprocess.DEVNULL)
    out = subprocess.run([str(exe), "300000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok:")


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_viterbi_pair_kernel_arithmetic_equals_oracle_on_the_host(tmp_path):
    """csrc/viterbi_pair_core.h (branch metrics, packed ACS step, renormalisation, decision layout, traceback step of
    the two-frames-per-thread kernel) with the DPX instructions emulated half by half, over whole frames of several
    sizes (F % 32 != 0 included) and adversarial symbol mixes."""
    import oracle_lib

    oracle_lib.port()
    exe = tmp_path / "viterbi_pair_check"
    oracle_dir = os.path.join(ROOT, "oracle")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", str(exe), os.path.join(ROOT, "tests", "host", "viterbi_pair_check.cpp"),
                    "-L" + oracle_dir, "-lfec_oracle", "-Wl,-rpath," + oracle_dir], check=True, stderr=subprocess.DEVNULL)
    out = subprocess.run([str(exe), "40"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok:")


def test_generated_tables_are_current():
    """rs_bitslice_tables.h is generated; it must match its generator."""
    gen = subprocess.run(["python", os.path.join(ROOT, "viterbi.dll_b200", "csrc", "gen_rs_bitslice_tables.py")],
                         capture_output=True, text=True, check=True).stdout
    assert gen == open(os.path.join(ROOT, "viterbi.dll_b200", "csrc", "rs_bitslice_tables.h")).read()
