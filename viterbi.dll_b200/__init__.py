"""viterbi.dll_b200 -- host-side mirror of viterbi.dll's FEC interface over libviterbi_b200.so.

The product is the C-ABI shared library built from ``csrc/`` (see include/viterbi_b200.h); this
module is the thin ctypes binding the tests and the benchmark use.  Function names, argument
meaning and return conventions follow the reference exports (viterbi.def:4-8):

    deconvolve(framebits, piData, inputLength, output)          deconvolve.cpp:551-554
    RScheckSuperframe(p, startIx, RSDims, outVector)            rschecksf.cpp:65-93
    initialize()                                                dllmain.cpp:156-160

plus the batched entry points.  There is NO CPU fallback: if the CUDA library cannot be built
or loaded, importing this module raises.

The directory name contains a dot, so it is imported through the loader shim
``viterbi_dll_b200.py`` at the repository root (``import viterbi_dll_b200 as vb``).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build
from . import dabgen  # noqa: F401  (synthetic traffic generator, re-exported)

_vp = ctypes.c_void_p
MAX_FRAMEBITS = 9216
FEC_OK, FEC_ERR_ARG, FEC_ERR_DEVICE = 0, 1, 2

# every symbol include/viterbi_b200.h declares: (restype, argtypes)
_SIGNATURES = {
    "deconvolve": (ctypes.c_int, [ctypes.c_uint, _vp, ctypes.c_int, _vp]),
    "RScheckSuperframe": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_uint, _vp]),
    "RSCheckSuperframe": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_uint, _vp]),
    "initialize": (ctypes.c_int, []),
    "GetCPUCaps": (ctypes.c_int, []),
    "WakeUpYMM": (None, []),
    "viterbi_deconvolve_batch": (ctypes.c_int, [ctypes.c_uint, _vp, ctypes.c_size_t, _vp]),
    "viterbi_deconvolve_batch_u32": (ctypes.c_int, [ctypes.c_uint, _vp, ctypes.c_size_t, _vp]),
    "viterbi_deconvolve_batch_device": (ctypes.c_int, [ctypes.c_uint, _vp, ctypes.c_size_t, _vp, _vp]),
    "viterbi_deconvolve_batch_u32_device": (ctypes.c_int, [ctypes.c_uint, _vp, ctypes.c_size_t, _vp, _vp]),
    "viterbi_deconvolve_batch_punctured": (ctypes.c_int, [ctypes.c_uint, _vp, ctypes.c_size_t, _vp, ctypes.c_uint,
                                                            ctypes.c_size_t, _vp]),
    "viterbi_deconvolve_batch_punctured_device": (ctypes.c_int, [ctypes.c_uint, _vp, ctypes.c_size_t, _vp, ctypes.c_uint,
                                                                   ctypes.c_size_t, _vp, _vp]),
    "rs_check_superframe_batch": (ctypes.c_int, [_vp, ctypes.c_uint, ctypes.c_size_t, _vp, _vp]),
    "rs_check_superframe_batch_device": (ctypes.c_int, [_vp, ctypes.c_uint, ctypes.c_size_t, _vp, _vp, _vp]),
    "dabplus_decode_superframes": (ctypes.c_int, [ctypes.c_uint, _vp, ctypes.c_size_t, _vp, _vp]),
    "dabplus_decode_superframes_device": (ctypes.c_int, [ctypes.c_uint, _vp, ctypes.c_size_t, _vp, _vp, _vp]),
    "viterbi_deconvolve_batch_multi": (ctypes.c_int, [ctypes.c_uint, _vp, ctypes.c_size_t, _vp]),
    "rs_check_superframe_batch_multi": (ctypes.c_int, [_vp, ctypes.c_uint, ctypes.c_size_t, _vp, _vp]),
    "dabplus_decode_superframes_multi": (ctypes.c_int, [ctypes.c_uint, _vp, ctypes.c_size_t, _vp, _vp]),
    "fec_set_devices": (ctypes.c_int, [_vp, ctypes.c_int]),
    "fec_get_devices": (ctypes.c_int, [_vp, ctypes.c_int]),
    "fec_allgather_device": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    "fec_set_thread_device": (ctypes.c_int, [ctypes.c_int]),
    "rs_check_superframe_batch_device_bcast": (ctypes.c_int, [_vp, ctypes.c_uint, ctypes.c_size_t, _vp, _vp, _vp, _vp, ctypes.c_int, _vp]),
    "dabplus_decode_superframes_device_bcast": (ctypes.c_int, [ctypes.c_uint, _vp, ctypes.c_size_t, _vp, _vp, _vp, _vp, ctypes.c_int, _vp]),
    "fec_enable_peer_access": (ctypes.c_int, []),
    "fec_ipc_export": (ctypes.c_int, [_vp, _vp]),
    "fec_ipc_import": (_vp, [_vp]),
    "fec_ipc_close": (ctypes.c_int, [_vp]),
    "fec_device_count": (ctypes.c_int, []),
    "fec_set_device": (ctypes.c_int, [ctypes.c_int]),
    "fec_get_device": (ctypes.c_int, []),
    "fec_in_save_mode": (ctypes.c_int, []),
    "fec_last_error": (ctypes.c_char_p, []),
    "fec_host_alloc": (_vp, [ctypes.c_size_t]),
    "fec_host_free": (None, [_vp]),
    "fec_device_alloc": (_vp, [ctypes.c_size_t]),
    "fec_device_free": (None, [_vp]),
    "fec_memcpy_h2d": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t]),
    "fec_memcpy_d2h": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t]),
    "fec_memcpy_d2d_async": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    "fec_device_synchronize": (ctypes.c_int, []),
    "fec_set_viterbi_kernel": (ctypes.c_int, [ctypes.c_int]),
    "fec_set_energy_dispersal": (ctypes.c_int, [ctypes.c_int]),
    "fec_kernel_launches": (ctypes.c_ulonglong, []),
}

LIB_PATH = _build.LIB


def _load() -> ctypes.CDLL:
    # VITERBI_B200_LIB: developer override to load an alternative build of the same C ABI
    path = os.environ.get("VITERBI_B200_LIB") or _build.build_library()  # rebuilds only when stale; raises without nvcc
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class FecError(RuntimeError):
    pass


def _check(rc: int, what: str) -> None:
    if rc != FEC_OK:
        raise FecError("%s failed (rc=%d): %s" % (what, rc, (lib.fec_last_error() or b"").decode()))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


# -------------------------------------------------------------------------------------------
# drop-in surface
# -------------------------------------------------------------------------------------------
def deconvolve(framebits: int, piData: np.ndarray, inputLength: int = 0, output: np.ndarray | None = None):
    """One frame, QIRX layout (uint32 per soft symbol).  Returns (rc, output bytes)."""
    sym = np.ascontiguousarray(piData, dtype=np.uint32)
    if output is None:
        output = np.zeros((framebits + 7) // 8, dtype=np.uint8)
    rc = lib.deconvolve(framebits, _ptr(sym), inputLength, _ptr(output))
    return rc, output


def RScheckSuperframe(p: np.ndarray, startIx: int, RSDims: int, outVector: np.ndarray) -> int:
    p = np.ascontiguousarray(p, dtype=np.uint8)
    assert outVector.dtype == np.uint8 and outVector.flags.c_contiguous
    return lib.RScheckSuperframe(_ptr(p), startIx, RSDims, _ptr(outVector))


def initialize() -> bool:
    return bool(lib.initialize())


# -------------------------------------------------------------------------------------------
# batched, host buffers
# -------------------------------------------------------------------------------------------
def deconvolve_batch(framebits: int, syms: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    """syms [n, 4*(F+6)] uint8 (or uint32, QIRX layout) -> [n, ceil(F/8)] uint8."""
    syms = np.ascontiguousarray(syms)
    n = syms.shape[0]
    if syms.shape[1] != 4 * (framebits + 6):
        raise ValueError("syms must be [n, 4*(framebits+6)]")
    if out is None:
        out = np.zeros((n, (framebits + 7) // 8), dtype=np.uint8)
    if syms.dtype == np.uint8:
        rc = lib.viterbi_deconvolve_batch(framebits, _ptr(syms), n, _ptr(out))
    elif syms.dtype == np.uint32:
        rc = lib.viterbi_deconvolve_batch_u32(framebits, _ptr(syms), n, _ptr(out))
    else:
        raise TypeError("syms must be uint8 or uint32")
    _check(rc, "viterbi_deconvolve_batch")
    return out


def deconvolve_batch_punctured(framebits: int, rx: np.ndarray, keep: np.ndarray, erasure: int = 128,
                               out: np.ndarray | None = None) -> np.ndarray:
    """rx [n, kept] uint8 = the transmitted soft symbols only; keep [4*(F+6)] non-zero where transmitted."""
    rx = np.ascontiguousarray(rx, dtype=np.uint8)
    keep = np.ascontiguousarray(keep, dtype=np.uint8)
    if keep.shape != (4 * (framebits + 6),):
        raise ValueError("keep must be [4*(framebits+6)]")
    n = rx.shape[0]
    if out is None:
        out = np.zeros((n, (framebits + 7) // 8), dtype=np.uint8)
    _check(lib.viterbi_deconvolve_batch_punctured(framebits, _ptr(rx), rx.shape[1], _ptr(keep), erasure, n, _ptr(out)),
           "viterbi_deconvolve_batch_punctured")
    return out


def rs_check_superframe_batch(rx: np.ndarray, RSDims: int, out: np.ndarray | None = None, fill: int = 0):
    """rx [n, 120*s] -> (out [n, 110*s], ret [n] int32).  `out` is updated in place when given."""
    rx = np.ascontiguousarray(rx, dtype=np.uint8)
    n = rx.shape[0]
    if out is None:
        out = np.full((n, 110 * RSDims), fill, dtype=np.uint8)
    ret = np.zeros(n, dtype=np.int32)
    _check(lib.rs_check_superframe_batch(_ptr(rx), RSDims, n, _ptr(out), _ptr(ret)), "rs_check_superframe_batch")
    return out, ret


def host_array(shape, dtype=np.uint8) -> np.ndarray:
    """Pinned host array from fec_host_alloc (freed with the array): host-pointer calls run at PCIe speed on
    pinned buffers (pageable ones go through the driver's bounce buffers)."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    ptr = lib.fec_host_alloc(max(nbytes, 1))
    if not ptr:
        raise FecError("fec_host_alloc failed: %s" % (lib.fec_last_error() or b"").decode())

    class _Owner:
        def __del__(self, _free=lib.fec_host_free, _p=ptr):
            _free(_p)

    buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(ptr)
    buf._owner = _Owner()  # the ctypes buffer keeps the allocation alive; numpy keeps the buffer
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


# -------------------------------------------------------------------------------------------
# batched, device buffers (torch tensors are only used as handles to HBM and streams)
# -------------------------------------------------------------------------------------------
def _stream_ptr(stream) -> int:
    if stream is None:
        import torch

        return torch.cuda.current_stream().cuda_stream
    return getattr(stream, "cuda_stream", stream)


def deconvolve_batch_device(framebits: int, syms, out=None, stream=None):
    """syms: CUDA uint8 tensor [n, 4*(F+6)] (or int32/uint32 words, QIRX layout).  Asynchronous on `stream`."""
    import torch

    n = syms.shape[0]
    if out is None:
        out = torch.empty((n, (framebits + 7) // 8), dtype=torch.uint8, device=syms.device)
    assert syms.is_contiguous() and out.is_contiguous() and syms.shape[1] == 4 * (framebits + 6)
    if syms.dtype == torch.uint8:
        rc = lib.viterbi_deconvolve_batch_device(framebits, syms.data_ptr(), n, out.data_ptr(), _stream_ptr(stream))
    else:
        assert syms.element_size() == 4
        rc = lib.viterbi_deconvolve_batch_u32_device(framebits, syms.data_ptr(), n, out.data_ptr(), _stream_ptr(stream))
    _check(rc, "viterbi_deconvolve_batch_device")
    return out


def deconvolve_batch_punctured_device(framebits: int, rx, keep: np.ndarray, erasure: int = 128, out=None, stream=None):
    """rx: CUDA uint8 tensor [n, kept]; keep: host array [4*(F+6)].  Asynchronous on `stream`."""
    import torch

    keep = np.ascontiguousarray(keep, dtype=np.uint8)
    n = rx.shape[0]
    if out is None:
        out = torch.empty((n, (framebits + 7) // 8), dtype=torch.uint8, device=rx.device)
    assert rx.is_contiguous() and out.is_contiguous() and keep.shape == (4 * (framebits + 6),)
    rc = lib.viterbi_deconvolve_batch_punctured_device(framebits, rx.data_ptr(), rx.shape[1], _ptr(keep), erasure, n,
                                                       out.data_ptr(), _stream_ptr(stream))
    _check(rc, "viterbi_deconvolve_batch_punctured_device")
    return out


def rs_check_superframe_batch_device(rx, RSDims: int, out, ret=None, stream=None):
    """rx: CUDA uint8 tensor [n, 120*s]; out [n, 110*s] is updated in place (partial-write rule)."""
    import torch

    n = rx.shape[0]
    if ret is None:
        ret = torch.empty((n,), dtype=torch.int32, device=rx.device)
    assert rx.is_contiguous() and out.is_contiguous() and ret.is_contiguous()
    rc = lib.rs_check_superframe_batch_device(rx.data_ptr(), RSDims, n, out.data_ptr(), ret.data_ptr(), _stream_ptr(stream))
    _check(rc, "rs_check_superframe_batch_device")
    return out, ret


def dabplus_decode_superframes(framebits: int, syms: np.ndarray, out: np.ndarray | None = None, fill: int = 0):
    """syms [nsf*5, 4*(F+6)] u8 -> Viterbi -> RS check -> (out [nsf, 110*s], ret [nsf]), s = F/192."""
    syms = np.ascontiguousarray(syms, dtype=np.uint8)
    nsf, s = syms.shape[0] // 5, framebits // 192
    assert syms.shape == (nsf * 5, 4 * (framebits + 6))
    if out is None:
        out = np.full((nsf, 110 * s), fill, dtype=np.uint8)
    ret = np.zeros(nsf, dtype=np.int32)
    _check(lib.dabplus_decode_superframes(framebits, _ptr(syms), nsf, _ptr(out), _ptr(ret)), "dabplus_decode_superframes")
    return out, ret


def dabplus_decode_superframes_device(framebits: int, syms, out, ret=None, stream=None):
    """Device-resident pipeline: syms CUDA u8 [nsf*5, 4*(F+6)], out CUDA u8 [nsf, 110*s] (updated in place)."""
    import torch

    nsf = syms.shape[0] // 5
    if ret is None:
        ret = torch.empty((nsf,), dtype=torch.int32, device=syms.device)
    assert syms.is_contiguous() and out.is_contiguous() and out.shape == (nsf, 110 * (framebits // 192))
    rc = lib.dabplus_decode_superframes_device(framebits, syms.data_ptr(), nsf, out.data_ptr(), ret.data_ptr(),
                                               _stream_ptr(stream))
    _check(rc, "dabplus_decode_superframes_device")
    return out, ret


# -------------------------------------------------------------------------------------------
# multi-device host calls (one process, all selected GPUs)
# -------------------------------------------------------------------------------------------
def set_devices(ordinals=None) -> None:
    """Devices of the *_multi calls; None / empty = all visible devices."""
    ordinals = list(ordinals or [])
    arr = (ctypes.c_int * max(len(ordinals), 1))(*ordinals)
    _check(lib.fec_set_devices(arr, len(ordinals)), "fec_set_devices")


def get_devices() -> list[int]:
    arr = (ctypes.c_int * 64)()
    n = lib.fec_get_devices(arr, 64)
    return [int(arr[i]) for i in range(min(n, 64))]


def deconvolve_batch_multi(framebits: int, syms: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    """Like deconvolve_batch (u8 layout), sharded over the selected devices inside the library."""
    syms = np.ascontiguousarray(syms, dtype=np.uint8)
    n = syms.shape[0]
    if syms.shape[1] != 4 * (framebits + 6):
        raise ValueError("syms must be [n, 4*(framebits+6)]")
    if out is None:
        out = np.zeros((n, (framebits + 7) // 8), dtype=np.uint8)
    _check(lib.viterbi_deconvolve_batch_multi(framebits, _ptr(syms), n, _ptr(out)), "viterbi_deconvolve_batch_multi")
    return out


def rs_check_superframe_batch_multi(rx: np.ndarray, RSDims: int, out: np.ndarray | None = None, fill: int = 0):
    rx = np.ascontiguousarray(rx, dtype=np.uint8)
    n = rx.shape[0]
    if out is None:
        out = np.full((n, 110 * RSDims), fill, dtype=np.uint8)
    ret = np.zeros(n, dtype=np.int32)
    _check(lib.rs_check_superframe_batch_multi(_ptr(rx), RSDims, n, _ptr(out), _ptr(ret)), "rs_check_superframe_batch_multi")
    return out, ret


def dabplus_decode_superframes_multi(framebits: int, syms: np.ndarray, out: np.ndarray | None = None, fill: int = 0):
    syms = np.ascontiguousarray(syms, dtype=np.uint8)
    nsf, s = syms.shape[0] // 5, framebits // 192
    assert syms.shape == (nsf * 5, 4 * (framebits + 6))
    if out is None:
        out = np.full((nsf, 110 * s), fill, dtype=np.uint8)
    ret = np.zeros(nsf, dtype=np.int32)
    _check(lib.dabplus_decode_superframes_multi(framebits, _ptr(syms), nsf, _ptr(out), _ptr(ret)),
           "dabplus_decode_superframes_multi")
    return out, ret


def allgather_device(shards, outs, streams=None) -> None:
    """ncclAllGather from one process: shards[i] / outs[i] are CUDA tensors on selected device i (equal byte sizes;
    outs[i] holds len(shards) shards).  Enqueued on each device's current stream unless `streams` is given."""
    import torch

    n = len(shards)
    nbytes = shards[0].numel() * shards[0].element_size()
    assert all(t.is_contiguous() and t.numel() * t.element_size() == nbytes for t in shards)
    assert all(o.is_contiguous() and o.numel() * o.element_size() == n * nbytes for o in outs)
    if streams is None:
        streams = [torch.cuda.current_stream(t.device).cuda_stream for t in shards]
    sp = (_vp * n)(*[t.data_ptr() for t in shards])
    op = (_vp * n)(*[t.data_ptr() for t in outs])
    st = (_vp * n)(*streams)
    _check(lib.fec_allgather_device(sp, op, nbytes, st), "fec_allgather_device")


class PeerBuffer:
    """A device buffer of this rank (fec_device_alloc) that every rank of the node has mapped (CUDA IPC): `local` is
    a torch uint8 view of this rank's buffer, `base[r]` the address of rank r's buffer in THIS process.  Used with the
    *_bcast calls: a kernel of this rank stores its results into every rank's buffer over NVLink.  Collective over
    the default process group: every rank constructs it with the same size, and calls close() together."""

    def __init__(self, nbytes: int, world: int, rank: int, device_index: int):
        import torch
        import torch.distributed as dist

        self.nbytes, self.world, self.rank = nbytes, world, rank
        self.ptr = lib.fec_device_alloc(nbytes)
        if not self.ptr:
            raise FecError("fec_device_alloc failed: %s" % (lib.fec_last_error() or b"").decode())
        handle = (ctypes.c_ubyte * 64)()
        _check(lib.fec_ipc_export(self.ptr, handle), "fec_ipc_export")
        mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda:%d" % device_index)
        allh = torch.empty((world, 64), dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(allh.view(-1), mine)
        allh = allh.cpu().numpy()
        self.base = []
        for r in range(world):
            if r == rank:
                self.base.append(self.ptr)
                continue
            p = lib.fec_ipc_import((ctypes.c_ubyte * 64)(*allh[r].tolist()))
            if not p:
                raise FecError("fec_ipc_import failed: %s" % (lib.fec_last_error() or b"").decode())
            self.base.append(p)

        class _Arr:  # zero-copy torch view of the raw allocation
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}

        self.local = torch.as_tensor(_Arr(), device="cuda:%d" % device_index)

    def peer_ptr(self, r: int, view) -> int:
        """Address, in rank r's buffer, of the bytes that `view` (a view of self.local) covers here."""
        return self.base[r] + (view.data_ptr() - self.ptr)

    def close(self):
        import torch.distributed as dist

        dist.barrier()  # nobody stores into a peer that is about to unmap / free
        for r, p in enumerate(self.base):
            if r != self.rank and p:
                lib.fec_ipc_close(p)
        dist.barrier()
        self.local = None
        lib.fec_device_free(self.ptr)
        self.base, self.ptr = [], None


def memcpy_d2d_async(dst, src, nbytes: int, stream=None):
    """Copy-engine copy between device buffers (tensors or raw addresses; peers' buffers included) on `stream`."""
    d = dst.data_ptr() if hasattr(dst, "data_ptr") else int(dst)
    s_ = src.data_ptr() if hasattr(src, "data_ptr") else int(src)
    _check(lib.fec_memcpy_d2d_async(d, s_, nbytes, _stream_ptr(stream)), "fec_memcpy_d2d_async")


def _ptr_array(ptrs):
    return (_vp * max(len(ptrs), 1))(*ptrs)


def rs_check_superframe_batch_device_bcast(rx, RSDims: int, out, ret, out_copies, ret_copies, stream=None):
    """Like rs_check_superframe_batch_device, and the results are also stored into `out_copies` / `ret_copies`: lists
    of CUDA tensors (same shapes as out / ret; typically peer GPUs' buffers) or raw device pointers (ints)."""
    n = rx.shape[0]
    oc = [getattr(t, "data_ptr", lambda t=t: t)() for t in out_copies]
    rc_ = [getattr(t, "data_ptr", lambda t=t: t)() for t in ret_copies]
    assert len(oc) == len(rc_) and rx.is_contiguous() and out.is_contiguous() and ret.is_contiguous()
    rc = lib.rs_check_superframe_batch_device_bcast(rx.data_ptr(), RSDims, n, out.data_ptr(), ret.data_ptr(), _ptr_array(oc),
                                                    _ptr_array(rc_), len(oc), _stream_ptr(stream))
    _check(rc, "rs_check_superframe_batch_device_bcast")
    return out, ret


def dabplus_decode_superframes_device_bcast(framebits: int, syms, out, ret, out_copies, ret_copies, stream=None):
    """Like dabplus_decode_superframes_device with extra destinations (see rs_check_superframe_batch_device_bcast)."""
    nsf = syms.shape[0] // 5
    oc = [getattr(t, "data_ptr", lambda t=t: t)() for t in out_copies]
    rc_ = [getattr(t, "data_ptr", lambda t=t: t)() for t in ret_copies]
    assert len(oc) == len(rc_) and syms.is_contiguous() and out.is_contiguous() and ret.is_contiguous()
    rc = lib.dabplus_decode_superframes_device_bcast(framebits, syms.data_ptr(), nsf, out.data_ptr(), ret.data_ptr(),
                                                     _ptr_array(oc), _ptr_array(rc_), len(oc), _stream_ptr(stream))
    _check(rc, "dabplus_decode_superframes_device_bcast")
    return out, ret


VITERBI_AUTO, VITERBI_PAIR, VITERBI_WARP = 0, 1, 2


def set_energy_dispersal(on: bool) -> None:
    """dabplus_* calls: remove the DAB energy dispersal (PRBS X^9 + X^5 + 1 per logical frame) between Viterbi and RS."""
    _check(lib.fec_set_energy_dispersal(1 if on else 0), "fec_set_energy_dispersal")


def set_viterbi_kernel(mode: int) -> None:
    """0 auto, 1 two-frames-per-thread throughput kernel, 2 warp-per-frame kernel (tests / measurements)."""
    _check(lib.fec_set_viterbi_kernel(mode), "fec_set_viterbi_kernel")


def kernel_launches() -> int:
    return int(lib.fec_kernel_launches())
