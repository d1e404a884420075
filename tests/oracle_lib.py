"""ctypes loaders for the CPU checkers (oracle/ -- test infrastructure only).

* ``port()``  -> oracle/libfec_oracle.so, our plain-C restatement.
* ``ref()``   -> oracle/_ref/libviterbi_ref_{avx512,avx2}.so, the reference's own
  deconvolve.cpp / rschecksf.cpp compiled unmodified (None when not built).
Both are built by ``make -C oracle`` (called from __graft_entry__.build()).
"""
from __future__ import annotations

import ctypes
import functools
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")
_vp = ctypes.c_void_p


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


def cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def ncores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class Port:
    """Plain-C restatement (oracle/fec_oracle.c)."""

    kind = "port"

    def __init__(self, lib):
        self.lib = lib
        lib.oracle_deconvolve.restype = ctypes.c_int
        lib.oracle_deconvolve.argtypes = [ctypes.c_uint, _vp, ctypes.c_int, _vp]
        lib.oracle_deconvolve_batch_u8.restype = ctypes.c_int
        lib.oracle_deconvolve_batch_u8.argtypes = [ctypes.c_uint, _vp, ctypes.c_size_t, _vp, ctypes.c_int]
        lib.oracle_rs_check_superframe.restype = ctypes.c_int
        lib.oracle_rs_check_superframe.argtypes = [_vp, ctypes.c_int, ctypes.c_uint, _vp]
        lib.oracle_rs_check_superframe_batch.restype = ctypes.c_int
        lib.oracle_rs_check_superframe_batch.argtypes = [_vp, ctypes.c_uint, ctypes.c_size_t, _vp, _vp, ctypes.c_int]
        lib.oracle_rs_tables.argtypes = [_vp, _vp]

    def deconvolve(self, framebits: int, syms32: np.ndarray) -> np.ndarray:
        s = np.ascontiguousarray(syms32, dtype=np.uint32)
        out = np.zeros((framebits + 7) // 8, dtype=np.uint8)
        rc = self.lib.oracle_deconvolve(framebits, _ptr(s), 0, _ptr(out))
        assert rc == 0, rc
        return out

    def deconvolve_batch(self, framebits: int, syms: np.ndarray, nthreads: int | None = None) -> np.ndarray:
        s = np.ascontiguousarray(syms, dtype=np.uint8)
        n = s.shape[0]
        assert s.shape[1] == 4 * (framebits + 6)
        out = np.zeros((n, (framebits + 7) // 8), dtype=np.uint8)
        rc = self.lib.oracle_deconvolve_batch_u8(framebits, _ptr(s), n, _ptr(out), nthreads or ncores())
        assert rc == 0, rc
        return out

    def rs_check_superframe(self, p: np.ndarray, s: int, out: np.ndarray) -> int:
        p = np.ascontiguousarray(p, dtype=np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous
        return self.lib.oracle_rs_check_superframe(_ptr(p), 0, s, _ptr(out))

    def rs_batch(self, rx: np.ndarray, s: int, fill: int = 0xEE, nthreads: int | None = None, out: np.ndarray | None = None):
        rx = np.ascontiguousarray(rx, dtype=np.uint8)
        n = rx.shape[0]
        if out is None:  # else: the caller's outVector contents, updated in place (partial-write rule)
            out = np.full((n, 110 * s), fill, dtype=np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.shape == (n, 110 * s)
        ret = np.zeros(n, dtype=np.int32)
        self.lib.oracle_rs_check_superframe_batch(_ptr(rx), s, n, _ptr(out), _ptr(ret), nthreads or ncores())
        return out, ret

    def rs_tables(self):
        ato = np.zeros(768, dtype=np.uint8)
        iof = np.zeros(256, dtype=np.uint8)
        self.lib.oracle_rs_tables(_ptr(ato), _ptr(iof))
        return ato, iof


class Ref:
    """The reference's own code (oracle/_ref)."""

    kind = "reference"

    def __init__(self, lib, isa: str):
        self.lib = lib
        self.isa = isa
        lib.ref_deconvolve.restype = ctypes.c_int
        lib.ref_deconvolve.argtypes = [ctypes.c_uint, _vp, ctypes.c_int, _vp]
        lib.ref_deconvolve_batch_u8.argtypes = [ctypes.c_uint, _vp, ctypes.c_size_t, _vp, ctypes.c_int]
        lib.ref_deconvolve_batch_u32.argtypes = [ctypes.c_uint, _vp, ctypes.c_size_t, _vp, ctypes.c_int]
        lib.ref_rs_check_superframe.restype = ctypes.c_int
        lib.ref_rs_check_superframe.argtypes = [_vp, ctypes.c_int, ctypes.c_uint, _vp]
        lib.ref_rs_check_superframe_batch.argtypes = [_vp, ctypes.c_uint, ctypes.c_size_t, _vp, _vp, ctypes.c_int]
        lib.ref_select.argtypes = [ctypes.c_int]
        lib.ref_select(4 if isa == "avx512" else 3)  # decon_avx5 / decon_avx2 (same C source)

    def select(self, which: int):
        return self.lib.ref_select(which)

    def deconvolve(self, framebits: int, syms32: np.ndarray) -> np.ndarray:
        s = np.ascontiguousarray(syms32, dtype=np.uint32)
        out = np.zeros((framebits + 7) // 8, dtype=np.uint8)
        rc = self.lib.ref_deconvolve(framebits, _ptr(s), 0, _ptr(out))
        assert rc == 0
        return out

    def deconvolve_batch(self, framebits: int, syms: np.ndarray, nthreads: int | None = None) -> np.ndarray:
        s = np.ascontiguousarray(syms, dtype=np.uint8)
        n = s.shape[0]
        out = np.zeros((n, (framebits + 7) // 8), dtype=np.uint8)
        self.lib.ref_deconvolve_batch_u8(framebits, _ptr(s), n, _ptr(out), nthreads or ncores())
        return out

    def deconvolve_batch_u32(self, framebits: int, syms32: np.ndarray, nthreads: int) -> np.ndarray:
        s = np.ascontiguousarray(syms32, dtype=np.uint32)
        n = s.shape[0]
        out = np.zeros((n, (framebits + 7) // 8), dtype=np.uint8)
        self.lib.ref_deconvolve_batch_u32(framebits, _ptr(s), n, _ptr(out), nthreads)
        return out

    def rs_check_superframe(self, p: np.ndarray, s: int, out: np.ndarray) -> int:
        p = np.ascontiguousarray(p, dtype=np.uint8)
        return self.lib.ref_rs_check_superframe(_ptr(p), 0, s, _ptr(out))

    def rs_batch(self, rx: np.ndarray, s: int, fill: int = 0xEE, nthreads: int | None = None, out: np.ndarray | None = None):
        rx = np.ascontiguousarray(rx, dtype=np.uint8)
        n = rx.shape[0]
        if out is None:  # else: the caller's outVector contents, updated in place (partial-write rule)
            out = np.full((n, 110 * s), fill, dtype=np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.shape == (n, 110 * s)
        ret = np.zeros(n, dtype=np.int32)
        self.lib.ref_rs_check_superframe_batch(_ptr(rx), s, n, _ptr(out), _ptr(ret), nthreads or ncores())
        return out, ret


def build(quiet: bool = True) -> None:
    """make -C oracle (port always; _ref only where /root/reference exists)."""
    subprocess.run(["make", "-C", ORACLE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


@functools.lru_cache(maxsize=None)
def port() -> Port:
    path = os.path.join(ORACLE, "libfec_oracle.so")
    if not os.path.exists(path):
        build()
    return Port(ctypes.CDLL(path))


@functools.lru_cache(maxsize=None)
def ref() -> Ref | None:
    flags = cpu_flags()
    order = []
    if {"avx512f", "avx512bw", "avx512vl"} <= flags:
        order.append("avx512")
    if "avx2" in flags:
        order.append("avx2")
    for isa in order:
        path = os.path.join(ORACLE, "_ref", f"libviterbi_ref_{isa}.so")
        if os.path.exists(path):
            return Ref(ctypes.CDLL(path), isa)
    return None


def checker():
    """Best available checker: the compiled reference, else the port."""
    return ref() or port()
