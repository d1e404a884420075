"""Synthetic DAB/DAB+ traffic for parity tests and benchmarks.

This is the portable counterpart of the generator inside the reference's only
test driver (viterbi-benchmark/viterbi-benchmark.cpp): random bits -> K=7 rate-1/4
convolutional encoder (polys 109,79,83,109; :64,304-311) -> 6 zero tail bits ->
AWGN -> 8-bit soft symbols ``(int)(127.5 + 32*(+-amp + n))`` clipped to 0..255
(:58-61,293-294,658-670).  MSVC rand() is not reproducible, so numpy / torch
generators seeded by the caller are used instead.  The RS(120,110) side
(systematic encoder, error injector, column interleaver) has no counterpart in the
reference, which never tests RScheckSuperframe functionally.

Nothing here is on the decode path; it only manufactures inputs.
"""
from __future__ import annotations

import math

import numpy as np

POLYS = (109, 79, 83, 109)  # viterbi-benchmark.cpp:64
K = 7
RATE = 4
GAIN = 32.0
OFFSET = 127.5


def nsym(framebits: int) -> int:
    """Soft symbols per frame: 4 per trellis step, F info + 6 tail steps."""
    return RATE * (framebits + K - 1)


def nout(framebits: int) -> int:
    return (framebits + 7) // 8


def noise_amp(ebn0_db: float) -> float:
    """Signal amplitude in noise-sigma units (viterbi-benchmark.cpp:293-294)."""
    esn0 = ebn0_db + 10.0 * math.log10(1.0 / RATE)
    return 1.0 / math.sqrt(0.5 / 10.0 ** (esn0 / 10.0))


def conv_encode(bits: np.ndarray) -> np.ndarray:
    """bits [n, F] in {0,1} -> code bits [n, 4*(F+6)] (tail-terminated)."""
    bits = np.asarray(bits, dtype=np.uint8)
    n, f = bits.shape
    padded = np.zeros((n, f + 2 * (K - 1)), dtype=np.uint8)  # 6 zeros before, 6 after
    padded[:, K - 1 : K - 1 + f] = bits
    steps = f + K - 1
    out = np.empty((n, steps, RATE), dtype=np.uint8)
    for j, poly in enumerate(POLYS):
        acc = np.zeros((n, steps), dtype=np.uint8)
        for d in range(K):  # bit d of the shift register = input d steps ago
            if (poly >> d) & 1:
                acc ^= padded[:, K - 1 - d : K - 1 - d + steps]
        out[:, :, j] = acc
    return out.reshape(n, steps * RATE)


def soft_symbols(code_bits: np.ndarray, ebn0_db: float, rng: np.random.Generator) -> np.ndarray:
    """AWGN channel + 8-bit quantiser of viterbi-benchmark.cpp:658-670."""
    amp = noise_amp(ebn0_db)
    x = rng.standard_normal(code_bits.shape, dtype=np.float32)
    x += (code_bits.astype(np.float32) * 2.0 - 1.0) * np.float32(amp)
    x = OFFSET + GAIN * x
    return np.clip(np.trunc(x), 0, 255).astype(np.uint8)


def pack_bits(bits: np.ndarray) -> np.ndarray:
    """MSB-first packing, the layout deconvolve() writes (deconvolve.cpp:416-435)."""
    return np.packbits(np.asarray(bits, dtype=np.uint8), axis=-1, bitorder="big")


def make_frames(n: int, framebits: int, ebn0_db: float, seed: int):
    """-> (symbols u8 [n, 4*(F+6)], packed info bits u8 [n, ceil(F/8)])."""
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, size=(n, framebits), dtype=np.uint8)
    return soft_symbols(conv_encode(bits), ebn0_db, rng), pack_bits(bits)


def lcg_symbols(seed: int, count: int) -> np.ndarray:
    """LCG byte stream used by the known-answer vectors V3/V4."""
    x = seed
    out = np.empty(count, dtype=np.uint8)
    for i in range(count):
        x = (1103515245 * x + 12345) & 0x7FFFFFFF
        out[i] = (x >> 16) & 0xFF
    return out


# ---------------------------------------------------------------------------
# Puncturing (the transmitter side of the depuncturing front end; not in viterbi.dll, which receives the
# already expanded rate-1/4 stream).  DAB punctures the mother code in blocks of 128 code bits = four
# repetitions of a 32-bit vector V_PI that keeps 8 + PI bits, PI = 1..24; the 24 tail code bits use a fixed
# vector that keeps 12.  The vectors below follow that structure (each PI adds one kept bit to PI - 1); the
# decoder API takes the pattern as data, so nothing on the decode path depends on this table.
# ---------------------------------------------------------------------------
_PI_ORDER = (0, 1, 4, 8, 12, 16, 20, 24, 28,      # PI = 1 keeps these 9 positions
             17, 9, 25, 5, 21, 13, 29,            # PI = 2..8: the second bit of each group of four
             2, 18, 10, 26, 6, 22, 14, 30,        # PI = 9..16: the third bit
             3, 19, 11, 27, 7, 23, 15, 31)        # PI = 17..24: the fourth bit


def puncture_vector(pi: int) -> np.ndarray:
    """32-entry keep vector with 8 + pi ones (pi = 1..24)."""
    if not 1 <= pi <= 24:
        raise ValueError("PI must be 1..24")
    v = np.zeros(32, dtype=np.uint8)
    v[list(_PI_ORDER[: 8 + pi])] = 1
    return v


TAIL_VECTOR = np.array([1, 1, 0, 0] * 6, dtype=np.uint8)  # 24 tail code bits -> 12


def puncture_pattern(framebits: int, segments) -> np.ndarray:
    """keep pattern [4*(F+6)] for a frame whose 4*F info-part code bits are cut into 128-bit blocks:
    segments = [(number_of_blocks, PI), ...] must cover 4*F/128 blocks; the 24 tail bits use TAIL_VECTOR."""
    if (4 * framebits) % 128:
        raise ValueError("framebits must be a multiple of 32")
    parts = [np.tile(puncture_vector(pi), 4 * nblk) for nblk, pi in segments]
    keep = np.concatenate(parts + [TAIL_VECTOR])
    if keep.size != 4 * (framebits + 6):
        raise ValueError("segments do not cover the frame")
    return keep


def fic_puncture_pattern() -> np.ndarray:
    """Transmission-mode-I FIC shape: 768 info bits, 21 blocks at PI = 16, 3 blocks at PI = 15, tail:
    3096 mother-code bits -> 2304 transmitted."""
    return puncture_pattern(768, [(21, 16), (3, 15)])


def puncture(syms: np.ndarray, keep: np.ndarray) -> np.ndarray:
    """[n, 4*(F+6)] -> the transmitted symbols only [n, keep.sum()]."""
    return np.ascontiguousarray(syms[:, np.asarray(keep, dtype=bool)])


def depuncture(rx: np.ndarray, keep: np.ndarray, erasure: int = 128) -> np.ndarray:
    """Host restatement of the expansion (for the checker side of the parity tests)."""
    keep = np.asarray(keep, dtype=bool)
    out = np.full((rx.shape[0], keep.size), erasure, dtype=np.uint8)
    out[:, keep] = rx
    return out


# ---------------------------------------------------------------------------
# torch generators (device-side synthetic input for the large benchmark configs)
# ---------------------------------------------------------------------------
def make_frames_torch(n: int, framebits: int, ebn0_db: float, seed: int, device, chunk: int = 8192,
                      want_bits: bool = False, payload=None):
    """Same channel model generated with torch on `device` -> u8 tensor [n, 4*(F+6)].

    Used only to manufacture benchmark inputs already resident in HBM; returns
    (symbols, packed_bits or None).  `payload` (u8 tensor [n, F/8], MSB-first) replaces the random info bits.
    """
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    steps = framebits + K - 1
    amp = noise_amp(ebn0_db)
    syms = torch.empty((n, steps * RATE), dtype=torch.uint8, device=device)
    packed = torch.empty((n, nout(framebits)), dtype=torch.uint8, device=device) if want_bits else None
    weights = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1], dtype=torch.uint8, device=device)
    shifts = torch.tensor([7, 6, 5, 4, 3, 2, 1, 0], dtype=torch.uint8, device=device)
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        if payload is None:
            bits = torch.randint(0, 2, (m, framebits), generator=g, device=device, dtype=torch.uint8)
        else:
            pb = payload[lo : lo + m].to(device)
            bits = ((pb[:, :, None] >> shifts) & 1).reshape(m, framebits)
        padded = torch.zeros((m, framebits + 2 * (K - 1)), dtype=torch.uint8, device=device)
        padded[:, K - 1 : K - 1 + framebits] = bits
        code = torch.empty((m, steps, RATE), dtype=torch.uint8, device=device)
        for j, poly in enumerate(POLYS):
            acc = torch.zeros((m, steps), dtype=torch.uint8, device=device)
            for d in range(K):
                if (poly >> d) & 1:
                    acc ^= padded[:, K - 1 - d : K - 1 - d + steps]
            code[:, :, j] = acc
        x = torch.randn((m, steps * RATE), generator=g, device=device, dtype=torch.float32)
        x += (code.reshape(m, -1).to(torch.float32) * 2.0 - 1.0) * amp
        x = OFFSET + GAIN * x
        syms[lo : lo + m] = torch.clamp(torch.trunc(x), 0, 255).to(torch.uint8)
        if want_bits:
            if framebits % 8:
                raise ValueError("want_bits needs framebits % 8 == 0")
            packed[lo : lo + m] = (bits.reshape(m, -1, 8) * weights).sum(dim=2, dtype=torch.int32).to(torch.uint8)
    return syms, packed


# ---------------------------------------------------------------------------
# Reed-Solomon RS(120,110) over GF(256)/0x11D, roots alpha^0..alpha^9
# ---------------------------------------------------------------------------
RS_N, RS_K, RS_T2 = 120, 110, 10


def _gf_tables():
    exp = np.zeros(512, dtype=np.int32)
    log = np.zeros(256, dtype=np.int32)
    sr = 1
    for i in range(255):
        exp[i] = sr
        log[sr] = i
        sr <<= 1
        if sr & 0x100:
            sr ^= 0x11D
    exp[255:510] = exp[0:255]
    return exp, log


_EXP, _LOG = _gf_tables()


def gf_mul(a, b):
    a = np.asarray(a, dtype=np.int32)
    b = np.asarray(b, dtype=np.int32)
    r = _EXP[(_LOG[a] + _LOG[b]) % 255]
    return np.where((a == 0) | (b == 0), 0, r).astype(np.uint8)


def rs_generator_poly() -> np.ndarray:
    """g(x) = prod_{i=0..9} (x - alpha^i), coefficients low -> high (11 entries)."""
    g = np.array([1], dtype=np.uint8)
    for i in range(RS_T2):
        root = np.uint8(_EXP[i])
        shifted = np.concatenate(([0], g)).astype(np.uint8)  # x * g
        scaled = np.concatenate((gf_mul(g, root), [0])).astype(np.uint8)  # root * g
        g = shifted ^ scaled
    return g


_GEN = rs_generator_poly()


def rs_encode(msg: np.ndarray) -> np.ndarray:
    """msg [n,110] -> systematic codewords [n,120]; byte 0 is the highest-degree coefficient."""
    msg = np.asarray(msg, dtype=np.uint8)
    n = msg.shape[0]
    reg = np.zeros((n, RS_T2), dtype=np.uint8)  # reg[:, 0] = highest-degree remainder coefficient
    ghi = _GEN[RS_T2 - 1 :: -1]  # g_9 .. g_0
    for k in range(RS_K):
        fb = msg[:, k] ^ reg[:, 0]
        reg[:, :-1] = reg[:, 1:]
        reg[:, -1] = 0
        reg ^= gf_mul(fb[:, None], ghi[None, :])
    return np.concatenate((msg, reg), axis=1)


def rs_inject_errors(cw: np.ndarray, nerr: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """XOR nerr[i] distinct non-zero byte errors into codeword i (cw [n,120])."""
    cw = np.array(cw, dtype=np.uint8, copy=True)
    n = cw.shape[0]
    nerr = np.broadcast_to(np.asarray(nerr), (n,))
    order = np.argsort(rng.random((n, RS_N)), axis=1)  # random distinct positions per row
    vals = rng.integers(1, 256, size=(n, RS_N), dtype=np.uint8)
    mask = np.arange(RS_N)[None, :] < nerr[:, None]
    rows = np.repeat(np.arange(n), RS_N).reshape(n, RS_N)
    cw[rows[mask], order[mask]] ^= vals[mask]
    return cw


def rs_interleave(cw: np.ndarray, s: int) -> np.ndarray:
    """cw [n*s,120] -> superframes [n,120*s] with codeword j byte k at j + k*s (rschecksf.cpp:75-76)."""
    n = cw.shape[0] // s
    return np.ascontiguousarray(cw.reshape(n, s, RS_N).transpose(0, 2, 1)).reshape(n, RS_N * s)


def make_superframes(n: int, s: int, seed: int, max_err: int = 7):
    """-> (received [n,120*s], clean payload [n,110*s], errors per codeword [n,s])."""
    rng = np.random.default_rng(seed)
    msg = rng.integers(0, 256, size=(n * s, RS_K), dtype=np.uint8)
    cw = rs_encode(msg)
    nerr = rng.integers(0, max_err + 1, size=n * s)
    rx = rs_inject_errors(cw, nerr, rng)
    payload = np.ascontiguousarray(msg.reshape(n, s, RS_K).transpose(0, 2, 1)).reshape(n, RS_K * s)
    return rs_interleave(rx, s), payload, nerr.reshape(n, s)


def make_superframes_torch(n: int, s: int, seed: int, device, max_err: int = 7, chunk: int = 1 << 18,
                           want_payload: bool = False):
    """Device-side version of make_superframes (benchmark inputs): -> (received [n,120*s] u8 tensor,
    errors per codeword [n,s]) and, with want_payload, the clean payload [n,110*s]."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    msgs = torch.empty((n * s, RS_K), dtype=torch.uint8, device=device) if want_payload else None
    exp = torch.from_numpy(_EXP.astype(np.int64)).to(device)
    log = torch.from_numpy(_LOG.astype(np.int64)).to(device)
    ghi_log = log[torch.from_numpy(_GEN[RS_T2 - 1 :: -1].astype(np.int64)).to(device)]  # g_9..g_0 (all non-zero)
    ncw = n * s
    rx = torch.empty((ncw, RS_N), dtype=torch.uint8, device=device)
    nerr_all = torch.empty((ncw,), dtype=torch.int64, device=device)
    for lo in range(0, ncw, chunk):
        m = min(chunk, ncw - lo)
        msg = torch.randint(0, 256, (m, RS_K), generator=g, device=device, dtype=torch.int64)
        reg = torch.zeros((m, RS_T2), dtype=torch.int64, device=device)
        for k in range(RS_K):
            fb = msg[:, k] ^ reg[:, 0]
            reg = torch.cat((reg[:, 1:], torch.zeros((m, 1), dtype=torch.int64, device=device)), dim=1)
            prod = exp[(log[fb][:, None] + ghi_log[None, :]) % 255]
            reg ^= torch.where(fb[:, None] == 0, torch.zeros_like(prod), prod)
        cw = torch.cat((msg, reg), dim=1)
        nerr = torch.randint(0, max_err + 1, (m,), generator=g, device=device)
        order = torch.argsort(torch.rand((m, RS_N), generator=g, device=device), dim=1)  # rank of each position
        vals = torch.randint(1, 256, (m, RS_N), generator=g, device=device, dtype=torch.int64)
        cw ^= torch.where(order < nerr[:, None], vals, torch.zeros_like(vals))
        rx[lo : lo + m] = cw.to(torch.uint8)
        nerr_all[lo : lo + m] = nerr
        if want_payload:
            msgs[lo : lo + m] = msg.to(torch.uint8)
    rx = rx.reshape(n, s, RS_N).transpose(1, 2).contiguous().reshape(n, RS_N * s)
    if want_payload:
        return rx, nerr_all.reshape(n, s), msgs.reshape(n, s, RS_K).transpose(1, 2).contiguous().reshape(n, RS_K * s)
    return rx, nerr_all.reshape(n, s)


# ---------------------------------------------------------------------------
# DAB+ pipeline traffic: RS-protected superframes carried in 5 convolutionally coded frames each
# ---------------------------------------------------------------------------
def energy_dispersal_prbs(nbits: int) -> np.ndarray:
    """DAB energy-dispersal sequence (ETSI EN 300 401 clause 10): PRBS of P(X) = X^9 + X^5 + 1, register
    initialised to all ones at the start of every logical frame.  First bits: 0000 0111 1011 1110 ..."""
    reg = [1] * 9
    out = np.empty(nbits, dtype=np.uint8)
    for i in range(nbits):
        b = reg[8] ^ reg[4]
        out[i] = b
        reg = [b] + reg[:8]
    return out


def make_superframe_frames(nsf: int, framebits: int, ebn0_db: float, seed: int, max_err: int = 0, scramble: bool = False):
    """-> (symbols u8 [nsf*5, 4*(F+6)], clean payload [nsf, 110*s], transmitted superframes [nsf, 120*s]).

    A superframe of s = F/192 interleaved RS(120,110) codewords is 5*F/8 bytes = the payload of five
    consecutive frames.  max_err > 0 additionally corrupts bytes BEFORE the convolutional encoder
    (errors the Viterbi decoder cannot remove), so the RS stage has work even on a clean channel.
    scramble: apply the DAB energy dispersal to every logical frame, as a real transmitter does."""
    if framebits % 192:
        raise ValueError("framebits must be a multiple of 192")
    s = framebits // 192
    rng = np.random.default_rng(seed)
    msg = rng.integers(0, 256, size=(nsf * s, RS_K), dtype=np.uint8)
    cw = rs_encode(msg)
    if max_err:
        cw = rs_inject_errors(cw, rng.integers(0, max_err + 1, size=nsf * s), rng)
    sf = rs_interleave(cw, s)  # [nsf, 120*s]
    bits = np.unpackbits(sf.reshape(nsf * 5, framebits // 8), axis=1, bitorder="big")
    if scramble:  # the transmitter's energy dispersal, per logical frame, in front of the convolutional encoder
        bits = bits ^ energy_dispersal_prbs(framebits)[None, :]
    syms = soft_symbols(conv_encode(bits), ebn0_db, rng)
    payload = np.ascontiguousarray(msg.reshape(nsf, s, RS_K).transpose(0, 2, 1)).reshape(nsf, RS_K * s)
    return syms, payload, sf


def make_superframe_frames_torch(nsf: int, framebits: int, ebn0_db: float, seed: int, device, max_err: int = 0):
    """Device-side version of make_superframe_frames (benchmark inputs for the end-to-end configuration):
    -> (symbols u8 [nsf*5, 4*(F+6)], clean payload u8 [nsf, 110*s])."""
    if framebits % 192:
        raise ValueError("framebits must be a multiple of 192")
    s = framebits // 192
    sf, _, payload = make_superframes_torch(nsf, s, seed, device, max_err=max_err, want_payload=True)
    syms, _ = make_frames_torch(nsf * 5, framebits, ebn0_db, seed + 1, device, payload=sf.reshape(nsf * 5, framebits // 8))
    return syms, payload
