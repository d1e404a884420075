"""Measure the integer-pipe instruction rates of the box WITH an NVML clock / throttle record (profiles/int_peaks.json).

    python profiles/int_peaks.py [iters]

Runs profiles/microbench/intbench (built here with nvcc if missing) and samples the SM clock and the throttle
reasons of GPU 0 through NVML every 5 ms while it runs.  Every op reports the wall-clock window it ran in, so the
clock quoted next to a rate is the median of the samples taken under THAT load.  `alu_pipe_tera_laneops_per_s` (the
P_int of SURVEY.md section 8d) is the best ALU-pipe rate; lane-ops per SM per clock = rate / (SMs x sampled clock).
"""
import json
import os
import subprocess
import sys
import threading
import time

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "microbench", "intbench")
SRC = EXE + ".cu"


def main():
    iters = sys.argv[1] if len(sys.argv) > 1 else "4000"
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(SRC):
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-o", EXE, SRC], check=True)
    import pynvml as nv

    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(0)
    samples, stop = [], threading.Event()
    bad = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
           nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
           nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}

    def sampler():
        while not stop.is_set():
            try:
                samples.append((time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                                nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            except Exception:
                pass
            time.sleep(0.005)

    th = threading.Thread(target=sampler, daemon=True)
    th.start()
    out = subprocess.run([EXE, iters], check=True, capture_output=True, text=True).stdout
    stop.set()
    th.join()
    rows, dev = [], None
    for line in out.splitlines():
        d = json.loads(line)
        if "device" in d:
            dev = d
            continue
        win = [(c, r) for t, c, r in samples if d["t_start"] + 0.05 <= t <= d["t_end"]]
        clocks = sorted(c for c, _ in win)
        reasons = sorted({name for _, r in win for bit, name in bad.items() if r & bit})
        mhz = clocks[len(clocks) // 2] if clocks else None
        d.update({"sm_mhz_median": mhz, "sm_mhz_min": clocks[0] if clocks else None, "clock_samples": len(clocks),
                  "throttle_reasons": reasons,
                  "laneops_per_sm_per_clk": (d["tera_laneops_per_s"] * 1e12 / (dev["sms"] * mhz * 1e6)) if mhz else None})
        d.pop("t_start"), d.pop("t_end")
        rows.append(d)
    alu = [r for r in rows if r["op"] in ("lop3", "viaddmnmx_u16x2", "iadd", "vimnmx_u16x2", "prmt", "shf")]
    best = max(alu, key=lambda r: r["tera_laneops_per_s"])
    res = {"device": dev, "sm_max_mhz": nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM),
           "alu_pipe_tera_laneops_per_s": best["tera_laneops_per_s"], "alu_pipe_op": best["op"],
           "alu_pipe_sm_mhz": best["sm_mhz_median"], "alu_pipe_laneops_per_sm_per_clk": best["laneops_per_sm_per_clk"],
           "throttle_reasons_any": sorted({x for r in rows for x in r["throttle_reasons"]}),
           "source": "profiles/int_peaks.json (profiles/int_peaks.py: intbench + NVML clock record)", "ops": rows}
    with open(os.path.join(HERE, "int_peaks.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "ops"}))


if __name__ == "__main__":
    main()
