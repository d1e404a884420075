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
Weak scaling: every rank decodes its own batch of that size; no collective on the data path (frames are
independent).  Every rank checks a slice of its own results against the CPU checker (oracle/_ref) outside the
timed regions; the mismatch counts are all-reduced into `parity_mismatches`.

Other keys of the JSON line:
  e2e            same metric through the C-ABI host-pointer call, pinned host buffers, H2D + D2H inside the timed region
  roofline       the bound that applies to the ACS kernel: warp-instruction issue slots (592 sub-partitions x SM clock),
                 from the per-launch instruction count of a committed ncu capture; HBM and SURVEY-8(d) views inside it
  cpu_baseline   the reference's own decoder from oracle/_ref on this box's host cores (rank 0, N = 1)
  extra.msc      BASELINE configs[2] shape (262,144 MSC frames per GPU)
  extra.dropin   BASELINE configs[0]: single-frame deconvolve() latency next to the CPU reference per call
  extra.configs4 BASELINE configs[4]: 2^24 MSC frames -> superframes, Viterbi + RS on device, STRONG scaling over the
                 N ranks, with the NCCL all-gather of the results inside the timed region, overlapped with compute
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # oracle_lib (CPU checker legs only)

FIC_FRAMES, FIC_BITS = 65536, 768
RS_TOTAL = 1_000_000
W_VIT_OPS_PER_STEP = 320.0  # u8 integer ops per trellis step
SUBPARTITIONS = 148 * 4     # warp schedulers of a B200: one warp-instruction per clock each


def load_dabgen():
    """The traffic generator, loaded by path: importing the package would dlopen the product library, which the
    reference arm must not map."""
    spec = importlib.util.spec_from_file_location("fec_dabgen", os.path.join(ROOT, "viterbi.dll_b200", "dabgen.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload_config(args, world):
    """`config` of the JSON line -- identical for both arms."""
    n, f = args.frames, args.framebits
    return {"workload": "batched FIC decode (BASELINE configs[1]): %d frames per GPU x F=%d info bits (+6 tail), "
                        "8-bit soft symbols, AWGN Eb/N0=%.1f dB" % (n, f, args.ebn0),
            "frames_per_gpu": n, "framebits": f, "ebn0_db": args.ebn0,
            "parallelism": "independent frames, contiguous shards, no data-path collective",
            "l2_policy": "input %d MB + decision scratch > 126 MB L2; no explicit flush" % (n * 4 * (f + 6) // 1000000)}


def load_json(name):
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        try:
            with open(p) as fh:
                return json.load(fh)
        except ValueError:
            return None
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def time_cpu_single_calls(chk, kind, framebits, sym_u32: np.ndarray, calls: int):
    """viterbi-benchmark.cpp:332-348: `calls` single deconvolve() calls on one thread; returns us per call.
    Raw ctypes calls on prepared pointers, the same way the GPU drop-in is timed."""
    fn = chk.lib.ref_deconvolve if kind == "reference" else chk.lib.oracle_deconvolve
    out = np.zeros((framebits + 7) // 8, dtype=np.uint8)
    ptrs = [sym_u32[i].ctypes.data for i in range(sym_u32.shape[0])]
    optr, m = out.ctypes.data, len(ptrs)
    for i in range(min(calls, 50)):
        fn(framebits, ptrs[i % m], 0, optr)
    t0 = time.perf_counter()
    for i in range(calls):
        fn(framebits, ptrs[i % m], 0, optr)
    return (time.perf_counter() - t0) / calls * 1e6


def run_reference(args):
    """--impl reference: the reference's CPU decoder on the host cores, same workload / metric."""
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    import oracle_lib

    dabgen = load_dabgen()
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
    sample = "%d FIC frames (F=%d) per step, all %d host threads, u32 symbol layout, one %d-frame batch per step " \
             "whatever --gpus says" % (n, f, cores, n)
    line = {
        "impl": "reference", "metric": "viterbi_decoded_gbit_per_s", "value": value, "unit": "Gbit/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "Gbit/s", "cores": cores, "kind": kind, "sample": sample, "implementation": what},
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

    import oracle_lib
    import viterbi_dll_b200 as vb
    from viterbi_dll_b200 import dabgen

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x: float, op="max") -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())

    chk, chk_kind, chk_what = cpu_checker()
    cores = oracle_lib.ncores()
    chk_threads = max(1, cores // world)  # every rank checks its own slice at the same time
    parity = {"frames_checked": 0, "superframes_checked": 0, "mismatches": 0}

    if args.only_configs4:  # development aid: just the strong-scaling chain (A/B of gather modes at N = 2..8)
        c4 = run_configs4(args, vb, dabgen, chk, chk_threads, parity, dev, rank, world, barrier, reduce_ranks)
        bad = int(reduce_ranks(float(parity["mismatches"]), "sum"))
        if rank == 0:
            print(json.dumps({"only": "configs4", "n_gpus": world, "parity_mismatches": bad, "configs4": c4}))
        if world > 1:
            barrier()
            dist.destroy_process_group()
        return 1 if bad else 0

    n, f = args.frames, args.framebits
    steps_per_frame, nsym, nout = f + 6, 4 * (f + 6), (f + 7) // 8
    syms, bits = dabgen.make_frames_torch(n, f, args.ebn0, seed=1234 + rank, device=dev, want_bits=(f % 8 == 0))
    out = torch.zeros((n, nout), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def vit_step():
        vb.deconvolve_batch_device(f, syms, out, stream)

    # ---- device-resident throughput ("value") -----------------------------------------------
    warm = max(args.warmup, 3)
    for _ in range(warm):
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
    ms_total = reduce_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = n * world * f / (ms_step * 1e-3) / 1e9
    clocks = clk.summary()

    cpu_syms = syms[: min(n, 32768)].cpu().numpy() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    # sanity inside the bench: decoded bits equal the payload on (nearly) all frames at 3 dB
    fer = float((out != bits).any(dim=1).float().mean().item()) if bits is not None else None

    # ---- parity of this rank's results against the CPU checker (outside the timed region, every rank) ----------
    nchk = min(n, args.parity_frames)
    sel = torch.linspace(0, n - 1, nchk, device=dev).long()  # spread over the whole batch, not just its head
    want = chk.deconvolve_batch(f, syms.index_select(0, sel).cpu().numpy(), chk_threads)
    parity["frames_checked"] += nchk
    parity["mismatches"] += int((want != out.index_select(0, sel).cpu().numpy()).any(axis=1).sum())

    # ---- roofline of the dominant kernel: instruction issue ------------------------------------------------------
    hbm_peak, hbm_src = load_peaks()
    int_peaks = load_json("int_peaks.json") or {}
    int_peak = float(int_peaks.get("alu_pipe_tera_laneops_per_s", 18.56))
    counts = (load_json("inst_counts.json") or {}).get("viterbi_pair_kernel", {})
    kern_s = (e0.elapsed_time(e1) / args.steps) * 1e-3  # this rank's average launch duration
    groups = (n + 63) // 64
    sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965
    issue_peak = SUBPARTITIONS * sm_mhz * 1e6  # warp-instructions per second the device can issue
    inst_per_gs = counts.get("warp_inst_per_group_step")  # warp-instructions per 64-frame trellis step incl. traceback
    alu_per_gs = counts.get("alu_pipe_inst_per_group_step")
    alg_bytes = n * (nsym + nout)  # 4(F+6) + F/8 bytes per frame
    ach_gbs = alg_bytes / kern_s / 1e9
    ach_tops = W_VIT_OPS_PER_STEP * n * steps_per_frame / kern_s / 1e12
    traffic = counts.get("dram_bytes_per_fic_launch") if (n, f) == (FIC_FRAMES, FIC_BITS) else None
    if inst_per_gs:
        inst = inst_per_gs * groups * steps_per_frame
        ach_issue = inst / kern_s
        roofline = {"bound": "issue", "achieved": ach_issue / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-inst/s",
                    "frac": ach_issue / issue_peak, "traffic": traffic, "kernel": "viterbi_pair_kernel",
                    "warp_inst_per_launch": inst,
                    "alu_pipe_frac": (alu_per_gs * groups * steps_per_frame * 2 / kern_s / issue_peak) if alu_per_gs else None,
                    "peak_source": "592 SM sub-partitions x %d MHz (median SM clock of this run's timed region)" % sm_mhz,
                    "count_source": "profiles/inst_counts.json (%s)" % counts.get("source", "?"),
                    "note": "the ACS recurrence is bound by warp-instruction issue (ALU pipe: 2 issue cycles per "
                            "instruction); HBM and the SURVEY 8(d) op-count views follow"}
    else:
        roofline = {"bound": "issue", "achieved": None, "peak": issue_peak / 1e9, "unit": "G warp-inst/s", "frac": None,
                    "traffic": traffic, "note": "profiles/inst_counts.json missing"}
    roofline["hbm"] = {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                       "algorithmic_bytes_per_launch": alg_bytes, "peak_source": hbm_src}
    roofline["int_alu_survey_8d"] = {"achieved": ach_tops, "peak": int_peak, "unit": "T u8-op/s vs T int32 lane-op/s",
                                     "frac": ach_tops / int_peak,
                                     "peak_source": int_peaks.get("source", "nominal 148 SM x 64 lanes x 1.965 GHz"),
                                     "note": "320 u8 ops per trellis step (SURVEY 8d); two 16-bit metrics ride in each "
                                             "32-bit lane-op, so this scale does not bound the kernel"}

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
        dt = reduce_ranks(time.perf_counter() - t0)
        barrier()
        if not torch.equal(h_out.to(dev), out):
            parity["mismatches"] += 1
        e2e = {"value": n * world * f * k_e2e / dt / 1e9, "unit": "Gbit/s", "h2d_bytes_per_step": n * nsym,
               "d2h_bytes_per_step": n * nout, "steps": k_e2e, "ms_per_step": dt / k_e2e * 1e3,
               "api": "viterbi_deconvolve_batch (pinned host buffers)"}
        pcie = load_json("pcie_r02.json")
        if pcie:
            e2e["host_h2d_gbs_measured"] = pcie.get("h2d_gbs_by_ranks", {}).get(str(world))
            e2e["pcie_source"] = "profiles/pcie_r02.json (aggregate pinned H2D over %d ranks)" % world

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
        rs_ms = reduce_ranks(r0.elapsed_time(r1)) / args.steps
        rs_bytes = sum(per_s * (230 * s + 4) for s in range(1, 9))
        rs_s_local = r0.elapsed_time(r1) / args.steps * 1e-3
        rs_gbs = rs_bytes / rs_s_local / 1e9
        rs_counts = (load_json("inst_counts.json") or {}).get("rs_superframe_kernel", {})
        ncw = sum(per_s * s for s in range(1, 9))
        rs_roof = {"bound": "issue", "unit": "G warp-inst/s", "peak": issue_peak / 1e9, "achieved": None, "frac": None,
                   "traffic": rs_counts.get("dram_bytes_per_codeword", 0) * ncw or None,
                   "hbm": {"achieved": rs_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": rs_gbs / hbm_peak,
                           "algorithmic_bytes_per_launch_set": rs_bytes}}
        if rs_counts.get("warp_inst_per_codeword"):
            ach = rs_counts["warp_inst_per_codeword"] * ncw / rs_s_local
            rs_roof.update({"achieved": ach / 1e9, "frac": ach / issue_peak,
                            "count_source": "profiles/inst_counts.json (%s)" % rs_counts.get("source", "?"),
                            "note": "latency-bound table walks (LDS chains); issue slots are the nearest hard ceiling"})
        # parity: a slice of every s against the checker
        nsl = min(per_s, args.parity_superframes // 8)
        for s, rx, o, r in sets:
            w_out, w_ret = chk.rs_batch(rx[:nsl].cpu().numpy(), s, fill=0xEE, nthreads=chk_threads)
            parity["superframes_checked"] += nsl
            parity["mismatches"] += int((w_ret != r[:nsl].cpu().numpy()).sum()) + int((w_out != o[:nsl].cpu().numpy()).any(axis=1).sum())
        rs = {"metric": "rs_superframes_per_s", "value": per_s * 8 * world / (rs_ms * 1e-3), "unit": "superframes/s",
              "ms_per_step": rs_ms, "gpu_launches": rs_launches,
              "config": {"workload": "%d DAB+ superframes per GPU, s=1..8 (%d each), 0-7 byte errors per codeword"
                                     % (per_s * 8, per_s)},
              "uncorrectable_frac": float(sum((r < 0).float().mean().item() for _, _, _, r in sets) / 8),
              "roofline": rs_roof}
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
            dt = reduce_ranks(time.perf_counter() - t0)
            for (s, _, h_o, h_r), (_, _, o, r) in zip(host, sets):
                if not (torch.equal(h_o.to(dev), o) and torch.equal(h_r.to(dev), r)):
                    parity["mismatches"] += 1
            rs["e2e"] = {"value": per_s * 8 * world * k_rs / dt, "unit": "superframes/s",
                         "h2d_bytes_per_step": sum(per_s * 120 * s for s in range(1, 9)),
                         "d2h_bytes_per_step": sum(per_s * (110 * s + 4) for s in range(1, 9)),
                         "api": "rs_check_superframe_batch (pinned host buffers; the caller's outVector bytes of failing "
                                "superframes are read by the kernel through the pinned mapping, nothing is uploaded)"}
            del host

    # ---- extras ---------------------------------------------------------------------------------------------------
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
            return reduce_ranks(a.elapsed_time(b)) / reps, a.elapsed_time(b) / reps

        # MSC batch (BASELINE configs[2] shape, the largest single-GPU Viterbi config)
        mn, mf = args.msc_frames, 3072
        msym, mbits = dabgen.make_frames_torch(mn, mf, 3.0, seed=4321 + rank, device=dev, want_bits=True)
        mout = torch.zeros((mn, mf // 8), dtype=torch.uint8, device=dev)
        ms_msc, ms_msc_local = timed(lambda: vb.deconvolve_batch_device(mf, msym, mout, stream), 5)
        nbad = int(((mout ^ mbits) != 0).any(dim=1).sum().item())
        msel = torch.linspace(0, mn - 1, min(mn, args.parity_frames // 4), device=dev).long()
        mwant = chk.deconvolve_batch(mf, msym.index_select(0, msel).cpu().numpy(), chk_threads)
        parity["frames_checked"] += int(msel.numel())
        parity["mismatches"] += int((mwant != mout.index_select(0, msel).cpu().numpy()).any(axis=1).sum())
        extra["msc"] = {"workload": "batched MSC decode (BASELINE configs[2] shape): %d frames per GPU x F=3072, Eb/N0=3 dB" % mn,
                        "value": mn * world * mf / (ms_msc * 1e-3) / 1e9, "unit": "Gbit/s", "ms_per_step": ms_msc,
                        "roofline_issue_frac": (inst_per_gs * ((mn + 63) // 64) * (mf + 6) / (ms_msc_local * 1e-3) / issue_peak)
                        if inst_per_gs else None,
                        "frame_error_rate": nbad / mn}
        del msym, mout, mbits
        torch.cuda.empty_cache()

        # the same FIC batches issued round-robin on two streams: a 65,536-frame batch is 1,024 warps on 592 SM
        # sub-partitions (0.43 of the resident grid), so a single launch leaves issue slots idle -- at its tail and
        # while all of its warps read their decisions back at once.  Consecutive batches on two streams fill them.
        s2 = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        out2 = [out, torch.empty_like(out)]
        for st_ in s2:
            st_.wait_stream(stream)
        def fic_two_streams(reps):
            for i in range(reps):
                vb.deconvolve_batch_device(f, syms, out2[i & 1], s2[i & 1])
        fic_two_streams(4)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for st_ in s2:
            st_.wait_stream(stream)
        fic_two_streams(args.steps)
        for st_ in s2:
            stream.wait_stream(st_)
        p1.record(stream)
        torch.cuda.synchronize()
        barrier()
        ms_pipe = reduce_ranks(p0.elapsed_time(p1)) / args.steps
        if not torch.equal(out2[0], out2[1]):
            parity["mismatches"] += 1
        extra["fic_two_streams"] = {"workload": "the headline FIC batches, %d steps issued alternately on two streams (consecutive batches overlap on the device)" % args.steps,
                                    "value": n * world * f / (ms_pipe * 1e-3) / 1e9, "unit": "Gbit/s", "ms_per_step": ms_pipe,
                                    "roofline_issue_frac": (inst_per_gs * groups * steps_per_frame / (p0.elapsed_time(p1) / args.steps * 1e-3) / issue_peak)
                                    if inst_per_gs else None}
        del out2

        # single-frame drop-in latency (BASELINE configs[0]; viterbi-benchmark.cpp:332-348: repeated calls on one thread)
        if rank == 0:
            drop = {}
            for df in (768, 3072):
                dsym, _ = dabgen.make_frames(64, df, 3.0, seed=77 + df)
                d32 = np.ascontiguousarray(dsym.astype(np.uint32))
                dout = np.zeros(df // 8, dtype=np.uint8)
                dwant = chk.deconvolve_batch(df, dsym, 1)
                ptrs = [d32[i].ctypes.data for i in range(64)]
                optr = dout.ctypes.data
                bad = 0
                for i in range(64):  # warm-up (graph capture, pinned bounce buffer) + parity of the path
                    if vb.lib.deconvolve(df, ptrs[i], 0, optr) != 0 or not np.array_equal(dout, dwant[i]):
                        bad += 1
                parity["frames_checked"] += 64
                parity["mismatches"] += bad
                calls = 2000
                t0 = time.perf_counter()
                for i in range(calls):
                    vb.lib.deconvolve(df, ptrs[i & 63], 0, optr)
                us = (time.perf_counter() - t0) / calls * 1e6
                drop["F%d" % df] = {"us_per_call": us, "cpu_reference_us_per_call": time_cpu_single_calls(chk, chk_kind, df, d32, 500)}
            extra["dropin"] = {"workload": "single-frame deconvolve() calls from one host thread, QIRX u32 layout (BASELINE configs[0])",
                               "api": "deconvolve (pinned bounce buffer, single-kernel CUDA graph, decisions in shared memory)",
                               **drop}

        # the same FIC batch through the QIRX word-per-symbol layout and through the punctured entry point
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
            dtu = reduce_ranks(time.perf_counter() - t0) / 5
            if not torch.equal(h_o32.to(dev), out):
                parity["mismatches"] += 1
            extra["e2e_u32_layout"] = {"workload": "same %d frames per GPU, one uint32 per soft symbol (QIRX layout), compacted on the device" % n,
                                       "value": n * world * f / dtu / 1e9, "unit": "Gbit/s", "ms_per_step": dtu * 1e3,
                                       "h2d_bytes_per_step": n * nsym * 4, "d2h_bytes_per_step": n * nout,
                                       "api": "viterbi_deconvolve_batch_u32 (pinned host buffers)"}
            del h_s32, h_o32
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
            dtp = reduce_ranks(time.perf_counter() - t0) / 10
            pwant = chk.deconvolve_batch(f, dabgen.depuncture(h_rx[:1024].numpy(), keep), chk_threads)
            parity["frames_checked"] += 1024
            parity["mismatches"] += int((pwant != h_pout[:1024].numpy()).any(axis=1).sum())
            extra["e2e_punctured_fic"] = {"workload": "same %d FIC frames per GPU as 2304 transmitted symbols each + keep pattern "
                                                      "(21 blocks PI=16, 3 blocks PI=15, tail), expanded on the device" % n,
                                          "value": n * world * f / dtp / 1e9, "unit": "Gbit/s", "ms_per_step": dtp * 1e3,
                                          "h2d_bytes_per_step": n * int(keep.sum()), "d2h_bytes_per_step": n * nout,
                                          "api": "viterbi_deconvolve_batch_punctured (pinned host buffers)"}
            if e2e is not None:
                e2e["punctured_input"] = {"value": extra["e2e_punctured_fic"]["value"], "unit": "Gbit/s",
                                          "h2d_bytes_per_step": n * int(keep.sum()),
                                          "note": "same frames sent as the 2304 transmitted symbols (what a receiver holds before "
                                                  "depuncturing): the e2e path for hosts where PCIe is the wall"}
            del h_rx, h_pout

        # BASELINE configs[4]: strong scaling of the whole chain with the result gather inside the timed region
        if not args.no_configs4:
            del syms, out
            torch.cuda.empty_cache()
            extra["configs4"] = run_configs4(args, vb, dabgen, chk, chk_threads, parity, dev, rank, world, barrier, reduce_ranks)

    # ---- CPU baseline (rank 0, single-GPU run only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        host_syms = cpu_syms
        all_gbps, frames_done = time_cpu_viterbi(chk, chk_kind, f, host_syms, cores, 4.0)
        one_gbps, _ = time_cpu_viterbi(chk, chk_kind, f, host_syms[:4096], 1, 2.0)
        cpu = {"value": all_gbps, "unit": "Gbit/s", "cores": cores, "kind": chk_kind,
               "sample": "%d-frame slice of the same FIC batch decoded repeatedly for >=4 s on all %d host threads "
                         "(%d frames in total), u32 symbol layout, distinct frames per call" % (host_syms.shape[0], cores, frames_done),
               "single_core": {"value": one_gbps, "unit": "Gbit/s", "cores": 1}, "implementation": chk_what,
               "cpu_model": next((l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "?")}
        if rs is not None:
            rs_sets = [(s, rx[:4000].cpu().numpy()) for s, rx, _, _ in sets]
            rs_all, _ = time_cpu_rs(chk, rs_sets, cores, 2.0)
            rs_one, _ = time_cpu_rs(chk, rs_sets, 1, 1.0)
            rs["cpu_baseline"] = {"value": rs_all, "unit": "superframes/s", "cores": cores, "kind": chk_kind,
                                  "sample": "8 x 4000 superframes of the same batch (s=1..8), >=2 s, all host threads",
                                  "single_core": {"value": rs_one, "unit": "superframes/s", "cores": 1}}

    # every rank's mismatch count, summed: SCALE runs prove bit-exactness at N > 1
    tot = {k: int(reduce_ranks(float(v), "sum")) for k, v in parity.items()}
    if rank == 0:
        line = {
            "metric": "viterbi_decoded_gbit_per_s", "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args, world),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "parity_mismatches": tot["mismatches"],
            "parity": {"checker": chk_what, "frames_checked_all_ranks": tot["frames_checked"],
                       "superframes_checked_all_ranks": tot["superframes_checked"], "ranks": world},
            "rs": rs, "frame_error_rate": fer, "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        barrier()  # no rank tears its communicator down while another is still inside a collective
        dist.destroy_process_group()
    if tot["mismatches"]:
        raise SystemExit("bench.py: %d results differ from the CPU checker" % tot["mismatches"])
    return 0


def configs4_round_sizes(my_sf: int, wave_sf: int, res_cap: int):
    """Superframes per round of one rank's share: whole waves, shrinking towards the end.  The gather of round j hides
    behind the decoding of round j + 1 as long as it is not much bigger than that round can cover (a wave's results take
    about a quarter of a wave's decoding time to gather at N = 8), and the gather of the LAST round cannot hide at all --
    so the last round is the remainder (0.5 .. 1.5 waves), the rounds before it double in size going backwards
    (2, 4, 8 waves, ...) up to the resident-symbol cap, and the first rounds take what is left in cap-sized pieces."""
    if my_sf <= wave_sf + wave_sf // 2:
        return [my_sf]
    tail = my_sf % wave_sf
    if tail < wave_sf // 2:
        tail += wave_sf
    rev, left, step = [tail], my_sf - tail, 2 * wave_sf
    while left > 0:
        sz = min(step, left, res_cap)
        if 0 < left - sz < wave_sf:  # no sliver in front: give it to this round if the cap allows, else split evenly
            sz = left if left <= res_cap else (left // 2) // wave_sf * wave_sf
        rev.append(sz)
        left -= sz
        step *= 2
    return rev[::-1]


def run_configs4(args, vb, dabgen, chk, chk_threads, parity, dev, rank, world, barrier, reduce_ranks):
    """BASELINE configs[4]: `--configs4-frames` (2^24) MSC frames = 3.36 M DAB+ superframes (s = 16), Viterbi + RS check
    on the device, STRONG scaling: the job is the same whatever N is, and the timed region ends when EVERY rank holds
    EVERY result (the gather of SURVEY 8e is inside it).

    The job is cut into rounds; in every round each rank decodes one contiguous slice, and the slices of a round sit
    side by side in the result array, so the gather of a round is one contiguous all-gather in natural superframe
    order.  A round is a whole number of WAVES of the persistent Viterbi grid (148 SMs x 16 warps x 64 frames), so
    cutting the work into rounds costs nothing at the tail of a launch; the rounds shrink towards the end
    (configs4_round_sizes: ..., 4, 2 waves, remainder), because the gather of the last round is the only one that
    cannot hide behind decoding.  A rank keeps at most 2^21 frames of symbols (25.8 GB) resident and reuses them (206 GB of
    symbols do not fit one GPU; a pass streams far more than the 126 MB L2, so a repeated pass costs what a fresh one
    does).  Gather, default (BENCH_C4_GATHER=dma): after each round the copy engines push the rank's results into
    every rank's array over NVLink (fec_memcpy_d2d_async on CUDA IPC mappings of the peers' arrays) -- no SM is taken
    from the decode kernels, measured at N = 8: 50.4 ms against 52.9 with NCCL and 57.9 with the RS kernel's own peer
    stores (49.1 without any gather).  BENCH_C4_GATHER=nccl: all_gather_into_tensor per round, in place, on a
    high-priority stream beside ONE compute stream (the collective and the next round then become runnable together and the priority puts the collective's
    few blocks first; with a second compute stream the next round's blocks refill every slot as it frees and the
    collective's large blocks starve until that grid is exhausted -- measured at N = 8: 58.2 vs 53.5 ms).
    BENCH_C4_GATHER=peer uses the fused alternative instead: the RS kernel stores its result tiles into every rank's
    array itself through CUDA IPC peer mappings (the *_bcast entry points; no collective at all).
    If the IPC mappings cannot be set up the run falls back to NCCL and says so in the line."""
    import torch
    import torch.distributed as dist

    f, s = 3072, 16
    row = 110 * s
    total_sf = args.configs4_frames // 5
    my_sf = -(-total_sf // world)
    wave_sf = (torch.cuda.get_device_properties(dev).multi_processor_count * 16 * 64) // 5  # superframes per wave
    res_cap = max(wave_sf, ((1 << 21) // 5) // wave_sf * wave_sf)  # resident symbols: whole waves, <= 2^21 frames
    # one rank has nothing to gather: cap-sized rounds (the resident symbols) and the remainder
    sizes = configs4_round_sizes(my_sf, wave_sf, res_cap) if world > 1 else [res_cap] * (my_sf // res_cap) + ([my_sf % res_cap] if my_sf % res_cap else [])
    rounds = len(sizes)
    offs = [sum(sizes[:j]) for j in range(rounds)]
    res_sf = min(my_sf, res_cap)
    job_sf = my_sf * world  # superframes actually decoded (>= total_sf)
    starts = []  # which resident superframes a round decodes: consecutive slices, wrapping when the set is exhausted
    cur = 0
    for sz in sizes:
        if cur + sz > res_sf:
            cur = 0
        starts.append(cur)
        cur += sz
    t0 = time.perf_counter()
    syms, payload = dabgen.make_superframe_frames_torch(res_sf, f, 4.0, seed=5000 + 17 * rank, device=dev, max_err=3)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0

    mode = os.environ.get("BENCH_C4_GATHER", "dma") if world > 1 else "none"
    buf_out = buf_ret = None
    if mode in ("peer", "dma"):
        try:
            buf_out = vb.PeerBuffer(job_sf * row, world, rank, dev.index)
            buf_ret = vb.PeerBuffer(job_sf * 4, world, rank, dev.index)
        except Exception as e:  # no IPC / no peer access on this box: fall back to the collective (all ranks fail alike)
            mode, buf_out, buf_ret = "nccl (peer mapping failed: %r)" % (e,), None, None
    if buf_out is not None:
        allout = buf_out.local.view(job_sf, row)
        allret = buf_ret.local.view(torch.int32)
        allout.fill_(0xEE)
        allret.fill_(-7)
    else:
        allout = torch.full((job_sf, row), 0xEE, dtype=torch.uint8, device=dev)
        allret = torch.full((job_sf,), -7, dtype=torch.int32, device=dev)
    barrier()  # every rank's arrays are pre-filled before anybody stores into them

    def region(j):  # rows of round j: `world` slices of sizes[j] superframes side by side
        return offs[j] * world, (offs[j] + sizes[j]) * world

    def mine(j, r=rank):  # rank r's slice of round j
        lo = offs[j] * world + r * sizes[j]
        return lo, lo + sizes[j]

    others = [r for r in range(world) if r != rank]
    nstreams = int(os.environ.get("BENCH_C4_STREAMS", "1"))
    comp = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
    comm = torch.cuda.Stream(device=dev, priority=-1)
    ncopy = int(os.environ.get("BENCH_C4_COPY_STREAMS", "4"))
    copy_streams = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(ncopy)] if mode == "dma" else []
    main = torch.cuda.current_stream()
    sync_flag = torch.zeros(1, dtype=torch.int32, device=dev)

    def run_job(gather: bool):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(main)
        for st in comp + [comm] + copy_streams:
            st.wait_stream(main)
        for j in range(rounds):
            st = comp[j % nstreams]
            sy = syms[starts[j] * 5:(starts[j] + sizes[j]) * 5]
            lo, hi = mine(j)
            o, r_ = allout[lo:hi], allret[lo:hi]
            if gather and mode == "peer":
                vb.dabplus_decode_superframes_device_bcast(f, sy, o, r_, [buf_out.peer_ptr(p, o) for p in others],
                                                           [buf_ret.peer_ptr(p, r_) for p in others], st)
                continue
            vb.dabplus_decode_superframes_device(f, sy, o, r_, st)
            if gather and mode == "dma":
                # the copy engines push this round's results into every peer's array over NVLink (no SM involved)
                done = torch.cuda.Event()
                done.record(st)
                for k, p in enumerate(others):
                    cs = copy_streams[k % ncopy]
                    cs.wait_event(done)
                    vb.memcpy_d2d_async(buf_out.peer_ptr(p, o), o, o.numel(), cs)
                    vb.memcpy_d2d_async(buf_ret.peer_ptr(p, r_), r_, 4 * r_.numel(), cs)
                continue
            if gather and world > 1:
                done = torch.cuda.Event()
                done.record(st)
                comm.wait_event(done)
                g0, g1 = region(j)
                with torch.cuda.stream(comm):
                    dist.all_gather_into_tensor(allout[g0:g1].view(-1), o.view(-1))
                    dist.all_gather_into_tensor(allret[g0:g1], r_)
        for st in comp + [comm] + copy_streams:
            main.wait_stream(st)
        if gather and mode in ("peer", "dma"):
            dist.all_reduce(sync_flag)  # on `main`: completes once EVERY rank's kernels (and their peer stores) are done
        end.record(main)
        torch.cuda.synchronize()
        return start.elapsed_time(end)

    run_job(False)  # warm-up: allocations, NCCL channels
    if world > 1:
        run_job(True)
    barrier()
    l0 = vb.kernel_launches()
    ms_gather = reduce_ranks(run_job(True))
    launches = vb.kernel_launches() - l0
    barrier()
    ms_nogather = reduce_ranks(run_job(False))
    barrier()

    # ---- checks: accepted superframes equal the transmitted payload; what this rank holds of every peer is what that
    # peer decoded; a slice is compared bit for bit with the CPU reference chain ----------------------------------
    run_job(True)
    lo, hi = mine(0)
    ok = allret[lo:hi] >= 0
    pay = payload[starts[0]:starts[0] + sizes[0]]
    wrong = int((allout[lo:hi][ok] != pay[ok]).any(dim=1).sum().item())
    accepted = int(ok.sum().item())
    if world > 1:
        # compare per-(round, rank) checksums of the gathered array with the checksums each producer computes over its
        # own slices (a checksum of checksums)
        barrier()

        def sums(a):
            return torch.stack([torch.stack([a[mine(j, r)[0]:mine(j, r)[1]].sum(dtype=torch.int64) for r in range(world)])
                                for j in range(rounds)])

        local = torch.stack([sums(allout), sums(allret)])  # [2, rounds, world] as seen here
        own = local[:, :, rank].contiguous()
        everyone = torch.empty((world,) + tuple(own.shape), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(everyone.view(-1), own.view(-1))
        wrong += int((everyone.permute(1, 2, 0) != local).sum().item()) + int((allret == -7).sum().item())
    nsl = min(sizes[0], args.parity_superframes // 8)
    h_syms = syms[starts[0] * 5:(starts[0] + nsl) * 5].cpu().numpy()
    dec = chk.deconvolve_batch(f, h_syms, chk_threads)
    c_out, c_ret = chk.rs_batch(dec.reshape(nsl, 120 * s), s, fill=0xEE, nthreads=chk_threads)
    parity["frames_checked"] += nsl * 5
    parity["superframes_checked"] += nsl
    parity["mismatches"] += int((c_ret != allret[lo:lo + nsl].cpu().numpy()).sum()) + \
        int((c_out != allout[lo:lo + nsl].cpu().numpy()).any(axis=1).sum()) + wrong
    frames = job_sf * 5
    gathered = int(allout.numel() + 4 * allret.numel()) if world > 1 else 0
    del allout, allret, ok, pay
    if buf_out is not None:
        buf_out.close()
        buf_ret.close()
    return {
        "workload": "BASELINE configs[4]: %d MSC frames (F=3072) -> %d DAB+ superframes (s=16), Viterbi + RS check on "
                    "device, Eb/N0=4 dB, 0-3 byte errors per codeword before the convolutional encoder" % (frames, job_sf),
        "scaling": "strong", "n_gpus": world, "rounds": rounds, "round_superframes": sizes,
        "resident_frames_per_gpu": res_sf * 5,
        "ms_total": ms_gather, "ms_total_without_gather": ms_nogather,
        "frames_per_s": frames / (ms_gather * 1e-3), "superframes_per_s": job_sf / (ms_gather * 1e-3),
        "viterbi_gbit_per_s": frames * f / (ms_gather * 1e-3) / 1e9,
        "gathered_bytes_per_rank": gathered,
        "gather": {"peer": "inside ms_total: the RS kernel of each rank stores its result tiles into every rank's result array "
                           "over NVLink (CUDA IPC peer mappings, dabplus_decode_superframes_device_bcast); a 4-byte all-reduce "
                           "closes the timed region",
                   "dma": "inside ms_total: after each round the copy engines push the rank's results into every rank's result "
                          "array over NVLink (fec_memcpy_d2d_async on CUDA IPC peer mappings, %d copy streams), overlapped with "
                          "the next round; a 4-byte all-reduce closes the timed region" % ncopy,
                   "none": "single rank: nothing to gather"}.get(mode, "inside ms_total: NCCL all_gather_into_tensor per round on a "
                                                                       "high-priority stream, in place into the full result array, "
                                                                       "overlapped with the next round (%s)" % mode),
        "compute_streams": nstreams,
        "gpu_launches_per_rank": launches, "rs_accepted_local": accepted, "rs_accepted_but_wrong_local": wrong,
        "generate_s": gen_s,
    }


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
    ap.add_argument("--no-extra", action="store_true", help="skip the MSC, drop-in, layout and configs[4] side measurements")
    ap.add_argument("--no-configs4", action="store_true", help="skip the 2^24-frame strong-scaling chain")
    ap.add_argument("--only-configs4", action="store_true", help="run nothing but the configs[4] chain (development aid)")
    ap.add_argument("--msc-frames", type=int, default=262144)
    ap.add_argument("--configs4-frames", type=int, default=1 << 24)
    ap.add_argument("--parity-frames", type=int, default=4096, help="frames per rank compared with the CPU checker")
    ap.add_argument("--parity-superframes", type=int, default=4096, help="superframes per rank compared with the CPU checker")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
