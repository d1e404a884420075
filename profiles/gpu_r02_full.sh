#!/bin/bash
# full single-GPU check: parity suite, smoke, bench line, latency, crossover
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02full_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02full_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02full_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02full_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02full_bench.json 2> gpurun_out/r02full_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02full_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02full_ref.json 2> gpurun_out/r02full_ref.err; echo "ref rc=$?"
g++ -O2 -std=c++17 -pthread -o /tmp/latbench profiles/microbench/latbench.cpp -ldl
REF=oracle/_ref/libviterbi_ref_avx512.so; grep -q avx512vl /proc/cpuinfo || REF=oracle/_ref/libviterbi_ref_avx2.so
timeout 300 /tmp/latbench viterbi.dll_b200/libviterbi_b200.so $REF 2000 > gpurun_out/r02full_latbench.jsonl 2>/dev/null; echo "latbench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02full_bench.json").read().splitlines() if l.startswith("{")][-1])
r = json.loads([l for l in open("gpurun_out/r02full_ref.json").read().splitlines() if l.startswith("{")][-1])
c = d["extra"]["configs4"]
print("value %.1f (2-stream %.1f) msc %.1f e2e %.2f punct %.2f rs %.1fM rs_e2e %.1fM parity %d | configs4 %.1f ms %.1f Gbit/s | ref %.2f rs %.2fM" % (
    d["value"], d["extra"]["fic_two_streams"]["value"], d["extra"]["msc"]["value"], d["e2e"]["value"], d["e2e"]["punctured_input"]["value"], d["rs"]["value"] / 1e6, d["rs"]["e2e"]["value"] / 1e6, d["parity_mismatches"],
    c["ms_total"], c["viterbi_gbit_per_s"], r["value"], r["rs"]["value"]/1e6))
print("roofline", d["roofline"]["frac"], d["roofline"]["alu_pipe_frac"], "msc", d["extra"]["msc"]["roofline_issue_frac"], "rs", d["rs"]["roofline"]["frac"], "dropin", d["extra"]["dropin"]["F768"], d["extra"]["dropin"]["F3072"])
PY
cat gpurun_out/r02full_latbench.jsonl | head -4
