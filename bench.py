#!/usr/bin/env python
"""bench.py -- throughput of the B200 FEC hot path (Viterbi `deconvolve` + `RScheckSuperframe`).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU code on the host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

A "step" is one pass of the hot path over one batch of synthetic input per GPU:
  * Viterbi: BASELINE.json configs[1] -- 65,536 FIC blocks (F = 768 info bits + 6 tail bits, 3,096
    eight-bit soft symbols each), random bits, AWGN at Eb/N0 = 3 dB (generator of
    viterbi-benchmark.cpp:293-311,658-670).  `value` = decoded info Gbit/s with inputs resident in HBM.
  * RS: BASELINE.json configs[3] -- 10^6 DAB+ superframes, s = 1..8 (125,000 each), 0-7 byte errors per
    codeword; reported under "rs" in the same JSON line (superframes/s).
Weak scaling: every rank decodes its own batch of that size; no collective on the data path
(frames are independent); the result bitstreams are gathered once with NCCL after the timed
region (reported as gather_ms).

The JSON line also carries: e2e (same metric through the C-ABI host-pointer call, pinned host
buffers, H2D + D2H inside the timed region), roofline (HBM view, contract shape) and
roofline_int_alu (the bound that actually applies to the ACS kernel), cpu_baseline (the
reference's own decoder from oracle/_ref on this box's host cores), clocks, gpu_launches.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # oracle_lib (CPU baseline legs only)

FIC_FRAMES, FIC_BITS = 65536, 768
RS_TOTAL = 1_000_000
W_VIT_OPS_PER_STEP = 320.0  # u8 integer ops per trellis step


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_int_peak():
    """Measured ALU-pipe lane-op rate (profiles/intbench_r01.jsonl), T lane-ops/s."""
    p = os.path.join(ROOT, "profiles", "intbench_r01.jsonl")
    best = None
    if os.path.exists(p):
        for line in open(p):
            try:
                d = json.loads(line)
            except ValueError:
                continue
            if d.get("op") in ("lop3", "viaddmnmx_u16x2", "iadd"):
                best = max(best or 0.0, float(d["tera_laneops_per_s"]))
    return (best, "measured (profiles/intbench_r01.jsonl)") if best else (18.6, "nominal 148 SM x 64 lanes x 1.965 GHz")


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while a timed region runs."""

    def __init__(self, torch_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            uuid = "GPU-" + str(torch.cuda.get_device_properties(torch_index).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_index)
            self.nv = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # no NVML: report that instead of inventing numbers
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable: " + self.err}
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


# ---------------------------------------------------------------------------------------------
# CPU legs (the only places that touch oracle/)
# ---------------------------------------------------------------------------------------------
def cpu_checker():
    import oracle_lib

    ref = oracle_lib.ref()
    if ref is not None:
        return ref, "reference", "oracle/_ref (deconvolve.cpp + rschecksf.cpp unmodified, g++ -O3, %s)" % ref.isa
    return oracle_lib.port(), "port", "oracle/fec_oracle.c (scalar C restatement)"


def time_cpu_viterbi(chk, kind, framebits, syms_u8: np.ndarray, threads: int, min_seconds: float):
    """Decode the given frames repeatedly until min_seconds elapsed; returns (Gbit/s, frames decoded)."""
    n = syms_u8.shape[0]
    if kind == "reference":
        s32 = syms_u8.astype(np.uint32)  # the reference's one-word-per-symbol layout; conversion untimed
        run = lambda: chk.deconvolve_batch_u32(framebits, s32, threads)  # noqa: E731
    else:
        run = lambda: chk.deconvolve_batch(framebits, syms_u8, threads)  # noqa: E731
    run()  # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    reps = 0
    while True:
        run()
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return n * reps * framebits / dt / 1e9, n * reps


def time_cpu_rs(chk, rs_sets, threads: int, min_seconds: float):
    chk.rs_batch(rs_sets[0][1][:256], rs_sets[0][0], nthreads=threads)
    t0 = time.perf_counter()
    done = 0
    while True:
        for s, rx in rs_sets:
            chk.rs_batch(rx, s, nthreads=threads)
            done += rx.shape[0]
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return done / dt, done


def run_reference(args):
    """--impl reference: the reference's CPU decoder on the host cores, same workload / metric."""
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    import oracle_lib
    from viterbi_dll_b200 import dabgen

    chk, kind, what = cpu_checker()
    cores = oracle_lib.ncores()
    n, f = args.frames, args.framebits
    syms, _ = dabgen.make_frames(n, f, args.ebn0, seed=1234)
    if kind == "reference":
        s32 = syms.astype(np.uint32)
        step = lambda: chk.deconvolve_batch_u32(f, s32, cores)  # noqa: E731
    else:
        step = lambda: chk.deconvolve_batch(f, syms, cores)  # noqa: E731
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = n * f * args.steps / dt / 1e9
    rs_val = None
    if not args.no_rs:
        per_s = 4000
        rs_sets = [(s, dabgen.make_superframes(per_s, s, seed=900 + s)[0]) for s in range(1, 9)]
        rs_val, _ = time_cpu_rs(chk, rs_sets, cores, 1.0)
    sample = "%d FIC frames (F=%d) per step, all %d host threads, u32 symbol layout" % (n, f, cores)
    line = {
        "impl": "reference", "metric": "viterbi_decoded_gbit_per_s", "value": value, "unit": "Gbit/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "batched FIC decode (BASELINE configs[1]): %d frames per GPU x F=%d info bits "
                               "(+6 tail), 8-bit soft symbols, AWGN Eb/N0=%.1f dB" % (n, f, args.ebn0),
                   "frames_per_gpu": n, "framebits": f, "implementation": what,
                   "sample": "one %d-frame batch per step on the host cores, whatever --gpus says" % n},
        "cpu_baseline": {"value": value, "unit": "Gbit/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rs": {"metric": "rs_superframes_per_s", "value": rs_val, "unit": "superframes/s",
               "sample": "8 x 4000 superframes, s=1..8, 0-7 errors/codeword, all host threads"} if rs_val else None,
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this implementation has no CPU path (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import viterbi_dll_b200 as vb
    from viterbi_dll_b200 import dabgen

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, f = args.frames, args.framebits
    steps_per_frame, nsym, nout = f + 6, 4 * (f + 6), (f + 7) // 8
    syms, bits = dabgen.make_frames_torch(n, f, args.ebn0, seed=1234 + rank, device=dev, want_bits=(f % 8 == 0))
    out = torch.zeros((n, nout), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def vit_step():
        vb.deconvolve_batch_device(f, syms, out, stream)

    # ---- device-resident throughput ("value") -----------------------------------------------
    for _ in range(max(args.warmup, 3)):
        vit_step()
    barrier()
    launches0 = vb.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        torch.cuda.nvtx.range_push("fec_timed_viterbi")  # ncu --nvtx --nvtx-include "fec_timed_viterbi/" = the timed region
        e0.record(stream)
        for _ in range(args.steps):
            vit_step()
        e1.record(stream)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
    launches = vb.kernel_launches() - launches0
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = n * world * f / (ms_step * 1e-3) / 1e9

    # sanity inside the bench: decoded bits equal the payload on (nearly) all frames at 3 dB
    fer = float((out != bits).any(dim=1).float().mean().item()) if bits is not None else None

    # ---- roofline of the dominant kernel -------------------------------------------------------
    hbm_peak, hbm_src = load_peaks()
    int_peak, int_src = load_int_peak()
    alg_bytes = n * (nsym + nout)  # 4(F+6) + F/8 bytes per frame
    kern_s = (e0.elapsed_time(e1) / args.steps) * 1e-3  # this rank's average launch duration
    ach_gbs = alg_bytes / kern_s / 1e9
    ach_tops = W_VIT_OPS_PER_STEP * n * steps_per_frame / kern_s / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("viterbi_pair_kernel_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                "traffic": traffic, "peak_source": hbm_src, "kernel": "viterbi_pair_kernel",
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "HBM is not the limiter of the ACS recurrence; see roofline_int_alu"}
    roofline_int = {"bound": "int_alu", "achieved": ach_tops, "peak": int_peak, "unit": "T u8-op/s vs T int32 lane-op/s",
                    "frac": ach_tops / int_peak, "peak_source": int_src,
                    "work": "320 u8 integer ops per trellis step x (F+6) steps x frames (SURVEY.md 8d); "
                            "the kernel packs two 16-bit metrics per 32-bit lane-op, so frac can exceed 1"}

    # ---- end to end through the C ABI with host buffers ------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_syms = torch.empty((n, nsym), dtype=torch.uint8, pin_memory=True)
        h_syms.copy_(syms)
        h_out = torch.empty((n, nout), dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        k_e2e = max(3, min(args.steps, 10))

        def e2e_step():
            rc = vb.lib.viterbi_deconvolve_batch(f, h_syms.data_ptr(), n, h_out.data_ptr())
            if rc != 0:
                raise RuntimeError("viterbi_deconvolve_batch rc=%d %s" % (rc, vb.lib.fec_last_error()))

        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        if not torch.equal(h_out.to(dev), out):
            raise RuntimeError("host-pointer path and device-pointer path disagree")
        e2e = {"value": n * world * f * k_e2e / dt / 1e9, "unit": "Gbit/s", "h2d_bytes_per_step": n * nsym,
               "d2h_bytes_per_step": n * nout, "steps": k_e2e, "ms_per_step": dt / k_e2e * 1e3,
               "api": "viterbi_deconvolve_batch (pinned host buffers)"}

    # ---- RS superframe check -----------------------------------------------------------------------
    rs = None
    if not args.no_rs:
        per_s = args.rs_superframes // 8
        sets = []
        for s in range(1, 9):
            rx, _ = dabgen.make_superframes_torch(per_s, s, seed=900 + s + 100 * rank, device=dev)
            sets.append((s, rx, torch.full((per_s, 110 * s), 0xEE, dtype=torch.uint8, device=dev),
                         torch.empty((per_s,), dtype=torch.int32, device=dev)))

        def rs_step():
            for s, rx, o, r in sets:
                vb.rs_check_superframe_batch_device(rx, s, o, r, stream)

        for _ in range(3):
            rs_step()
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = vb.kernel_launches()
        torch.cuda.nvtx.range_push("fec_timed_rs")
        r0.record(stream)
        for _ in range(args.steps):
            rs_step()
        r1.record(stream)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        rs_launches = vb.kernel_launches() - l0
        barrier()
        rs_ms = max_over_ranks(r0.elapsed_time(r1)) / args.steps
        rs_bytes = sum(per_s * (230 * s + 4) for s in range(1, 9))
        rs_gbs = rs_bytes / (r0.elapsed_time(r1) / args.steps * 1e-3) / 1e9
        rs = {"metric": "rs_superframes_per_s", "value": per_s * 8 * world / (rs_ms * 1e-3), "unit": "superframes/s",
              "ms_per_step": rs_ms, "gpu_launches": rs_launches,
              "config": {"workload": "%d DAB+ superframes per GPU, s=1..8 (%d each), 0-7 byte errors per codeword"
                                     % (per_s * 8, per_s)},
              "uncorrectable_frac": float(sum((r < 0).float().mean().item() for _, _, _, r in sets) / 8),
              "roofline": {"bound": "hbm", "achieved": rs_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": rs_gbs / hbm_peak,
                           "traffic": None, "algorithmic_bytes_per_launch_set": rs_bytes}}
        if not args.no_e2e:
            host = []
            for s, rx, o, _ in sets:
                h_rx = torch.empty(rx.shape, dtype=torch.uint8, pin_memory=True)
                h_rx.copy_(rx)
                h_o = torch.full(o.shape, 0xEE, dtype=torch.uint8, pin_memory=True)
                host.append((s, h_rx, h_o, torch.empty((per_s,), dtype=torch.int32, pin_memory=True)))
            torch.cuda.synchronize()

            def rs_e2e_step():
                for s, h_rx, h_o, h_r in host:
                    rc = vb.lib.rs_check_superframe_batch(h_rx.data_ptr(), s, per_s, h_o.data_ptr(), h_r.data_ptr())
                    if rc != 0:
                        raise RuntimeError("rs_check_superframe_batch rc=%d" % rc)

            rs_e2e_step()
            barrier()
            t0 = time.perf_counter()
            k_rs = 3
            for _ in range(k_rs):
                rs_e2e_step()
            dt = max_over_ranks(time.perf_counter() - t0)
            rs["e2e"] = {"value": per_s * 8 * world * k_rs / dt, "unit": "superframes/s",
                         "h2d_bytes_per_step": sum(per_s * 230 * s for s in range(1, 9)),
                         "d2h_bytes_per_step": sum(per_s * (110 * s + 4) for s in range(1, 9)),
                         "api": "rs_check_superframe_batch (pinned host buffers; outVector travels both ways)"}

    # ---- extras: MSC batch (BASELINE configs[2] shape) and the on-device DAB+ pipeline (configs[4] shape) ----
    extra = {}
    if not args.no_extra:
        def timed(fn, reps):
            for _ in range(2):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            barrier()
            return max_over_ranks(a.elapsed_time(b)) / reps

        mn, mf = args.msc_frames, 3072
        msym, mbits = dabgen.make_frames_torch(mn, mf, 3.0, seed=4321 + rank, device=dev, want_bits=True)
        mout = torch.zeros((mn, mf // 8), dtype=torch.uint8, device=dev)
        ms_msc = timed(lambda: vb.deconvolve_batch_device(mf, msym, mout, stream), 5)
        diff = (mout ^ mbits)
        nbad = int((diff != 0).any(dim=1).sum().item())
        extra["msc"] = {"workload": "batched MSC decode (BASELINE configs[2] shape): %d frames per GPU x F=3072, Eb/N0=3 dB" % mn,
                        "value": mn * world * mf / (ms_msc * 1e-3) / 1e9, "unit": "Gbit/s", "ms_per_step": ms_msc,
                        "roofline_int_alu_frac": W_VIT_OPS_PER_STEP * mn * (mf + 6) / (ms_msc * 1e-3) / 1e12 / int_peak,
                        "frame_error_rate": nbad / mn}
        # DAB+ pipeline on the first 5*k frames of the same symbols (payload is random, so most superframes
        # fail RS: the point is the cost of the fused chain, parity is covered by tests/test_gpu_rs.py)
        nsf = min(mn // 5, args.pipeline_superframes)
        psym = msym[: nsf * 5]
        pout = torch.full((nsf, 110 * 16), 0xEE, dtype=torch.uint8, device=dev)
        pret = torch.empty((nsf,), dtype=torch.int32, device=dev)
        ms_pipe = timed(lambda: vb.dabplus_decode_superframes_device(mf, psym, pout, pret, stream), 5)
        extra["dabplus_pipeline"] = {"workload": "%d MSC frames -> %d superframes (s=16) per GPU: Viterbi + RS check on device" % (nsf * 5, nsf),
                                     "superframes_per_s": nsf * world / (ms_pipe * 1e-3),
                                     "viterbi_gbit_per_s": nsf * 5 * world * mf / (ms_pipe * 1e-3) / 1e9, "ms_per_step": ms_pipe}
        del msym, mout, mbits, psym
        torch.cuda.empty_cache()
        # the same batch through the QIRX word-per-symbol layout (what the drop-in deconvolve() takes): 4x the bytes
        if not args.no_e2e:
            h_s32 = torch.empty((n, nsym), dtype=torch.int32, pin_memory=True)
            h_s32.copy_(syms.to(torch.int32))
            h_o32 = torch.empty((n, nout), dtype=torch.uint8, pin_memory=True)
            torch.cuda.synchronize()

            def u32_step():
                rc = vb.lib.viterbi_deconvolve_batch_u32(f, h_s32.data_ptr(), n, h_o32.data_ptr())
                if rc != 0:
                    raise RuntimeError("viterbi_deconvolve_batch_u32 rc=%d" % rc)

            for _ in range(2):
                u32_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(5):
                u32_step()
            dtu = max_over_ranks(time.perf_counter() - t0) / 5
            extra["e2e_u32_layout"] = {"workload": "same %d frames per GPU, one uint32 per soft symbol (QIRX layout), compacted on the device" % n,
                                       "value": n * world * f / dtu / 1e9, "unit": "Gbit/s", "ms_per_step": dtu * 1e3,
                                       "h2d_bytes_per_step": n * nsym * 4, "d2h_bytes_per_step": n * nout,
                                       "api": "viterbi_deconvolve_batch_u32 (pinned host buffers)"}
            del h_s32, h_o32
        # depuncturing front end: the same FIC batch sent as the 2,304 transmitted symbols per frame
        # (FIC-shaped puncturing) instead of the 3,096 expanded ones, host buffers, end to end.  The decoded
        # bits differ from the unpunctured run (erasures carry no information); parity of this path is in tests/.
        if f == 768 and not args.no_e2e:
            import ctypes

            keep = dabgen.fic_puncture_pattern()
            kidx = torch.from_numpy(np.flatnonzero(keep)).to(dev)
            h_rx = torch.empty((n, int(keep.sum())), dtype=torch.uint8, pin_memory=True)
            h_rx.copy_(syms.index_select(1, kidx))
            h_pout = torch.empty((n, nout), dtype=torch.uint8, pin_memory=True)
            torch.cuda.synchronize()

            def punct_step():
                rc = vb.lib.viterbi_deconvolve_batch_punctured(f, h_rx.data_ptr(), h_rx.shape[1],
                                                               keep.ctypes.data_as(ctypes.c_void_p), 128, n, h_pout.data_ptr())
                if rc != 0:
                    raise RuntimeError("viterbi_deconvolve_batch_punctured rc=%d" % rc)

            for _ in range(3):
                punct_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(10):
                punct_step()
            dtp = max_over_ranks(time.perf_counter() - t0) / 10
            extra["e2e_punctured_fic"] = {"workload": "same %d FIC frames per GPU as 2304 transmitted symbols each + keep pattern "
                                                      "(21 blocks PI=16, 3 blocks PI=15, tail)" % n,
                                          "value": n * world * f / dtp / 1e9, "unit": "Gbit/s", "ms_per_step": dtp * 1e3,
                                          "h2d_bytes_per_step": n * int(keep.sum()), "d2h_bytes_per_step": n * nout,
                                          "api": "viterbi_deconvolve_batch_punctured (pinned host buffers)"}

    # ---- gather of result bitstreams over NCCL (outside the timed region) -----------------------------
    gather_ms = None
    if world > 1:
        from viterbi_dll_b200 import sharding

        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        allout = sharding.gather_to_all(out, n * world, world, rank, align=64)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = max_over_ranks(g0.elapsed_time(g1))
        assert allout.shape[0] == n * world

    # ---- CPU baseline (rank 0, single-GPU run only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle_lib

        chk, kind, what = cpu_checker()
        cores = oracle_lib.ncores()
        host_syms = syms[: min(n, 32768)].cpu().numpy()
        want = chk.deconvolve_batch(f, host_syms[:4096])
        if not np.array_equal(want, out[:4096].cpu().numpy()):
            raise RuntimeError("GPU output differs from the CPU reference on the benchmark input")
        all_gbps, frames_done = time_cpu_viterbi(chk, kind, f, host_syms, cores, 4.0)
        one_gbps, _ = time_cpu_viterbi(chk, kind, f, host_syms[:4096], 1, 2.0)
        cpu = {"value": all_gbps, "unit": "Gbit/s", "cores": cores, "kind": kind,
               "sample": "%d-frame slice of the same FIC batch decoded repeatedly for >=4 s on all %d host threads "
                         "(%d frames in total), u32 symbol layout, distinct frames per call" % (host_syms.shape[0], cores, frames_done),
               "single_core": {"value": one_gbps, "unit": "Gbit/s", "cores": 1}, "implementation": what,
               "cpu_model": next((l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "?")}
        if rs is not None:
            rs_sets = [(s, rx[:4000].cpu().numpy()) for s, rx, _, _ in sets]
            rs_all, _ = time_cpu_rs(chk, rs_sets, cores, 2.0)
            rs_one, _ = time_cpu_rs(chk, rs_sets, 1, 1.0)
            rs["cpu_baseline"] = {"value": rs_all, "unit": "superframes/s", "cores": cores, "kind": kind,
                                  "sample": "8 x 4000 superframes of the same batch (s=1..8), >=2 s, all host threads",
                                  "single_core": {"value": rs_one, "unit": "superframes/s", "cores": 1}}

    if rank == 0:
        line = {
            "metric": "viterbi_decoded_gbit_per_s", "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "batched FIC decode (BASELINE configs[1]): %d frames per GPU x F=%d info bits "
                                   "(+6 tail), 8-bit soft symbols, AWGN Eb/N0=%.1f dB" % (n, f, args.ebn0),
                       "frames_per_gpu": n, "framebits": f, "parallelism": "independent frames, %d-way partition" % world,
                       "l2_policy": "input %d MB + decision scratch > 126 MB L2; no explicit flush" % (n * nsym // 1000000)},
            "e2e": e2e, "gpu_launches": launches, "clocks": clk.summary(), "roofline": roofline,
            "roofline_int_alu": roofline_int, "cpu_baseline": cpu, "rs": rs, "frame_error_rate": fer, "gather_ms": gather_ms,
            "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        barrier()  # no rank tears its communicator down while another is still inside a collective
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FIC_FRAMES)
    ap.add_argument("--framebits", type=int, default=FIC_BITS)
    ap.add_argument("--ebn0", type=float, default=3.0)
    ap.add_argument("--rs-superframes", type=int, default=RS_TOTAL)
    ap.add_argument("--no-rs", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the MSC and DAB+ pipeline side measurements")
    ap.add_argument("--msc-frames", type=int, default=262144)
    ap.add_argument("--pipeline-superframes", type=int, default=16384)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
