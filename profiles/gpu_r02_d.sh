#!/bin/bash
mkdir -p gpurun_out
python profiles/rs_e2e_ab.py > gpurun_out/r02d_rs_zero.json 2> gpurun_out/r02d_rs_zero.err; echo rc=$?
VITERBI_B200_RS_UPLOAD=1 python profiles/rs_e2e_ab.py > gpurun_out/r02d_rs_upload.json 2> gpurun_out/r02d_rs_upload.err; echo rc=$?
python - <<PY
import json
a=json.load(open("gpurun_out/r02d_rs_zero.json")); b=json.load(open("gpurun_out/r02d_rs_upload.json"))
for k in a:
    if isinstance(a[k],dict): print(k, "zero-copy %.2f ms %.1f M/s | upload %.2f ms %.1f M/s | failed %.2f" % (a[k]["ms"],a[k]["M_sf_per_s"],b[k]["ms"],b[k]["M_sf_per_s"],a[k]["failed_frac"]))
    else: print(k, a[k], b[k])
PY
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02d_bench.json").read().splitlines() if l.startswith("{")][-1])
c = d["extra"]["configs4"]
print("N=%d value %.1f e2e %.2f rs %.1fM rs_e2e %.1fM parity %d | configs4 %.1f ms (no gather %.1f) %.1f Gbit/s | dropin %s" % (
    d["n_gpus"], d["value"], d["e2e"]["value"], d["rs"]["value"] / 1e6, d["rs"]["e2e"]["value"] / 1e6, d["parity_mismatches"],
    c["ms_total"], c["ms_total_without_gather"], c["viterbi_gbit_per_s"], d["extra"]["dropin"]))
PY
