#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_viterbi.py tests/test_gpu_rs.py -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02e_pytest.log
timeout 600 python profiles/kernel_crossover.py > gpurun_out/r02e_crossover.jsonl 2> gpurun_out/r02e_crossover.err; echo "crossover rc=$?"
grep frames gpurun_out/r02e_crossover.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02e_bench.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02e_bench.json").read().splitlines() if l.startswith("{")][-1])
c = d["extra"]["configs4"]
print("N=%d value %.1f (2-stream %.1f) msc %.1f e2e %.2f rs %.1fM rs_e2e %.1fM parity %d | configs4 %.1f ms (no gather %.1f) %.1f Gbit/s" % (
    d["n_gpus"], d["value"], d["extra"]["fic_two_streams"]["value"], d["extra"]["msc"]["value"], d["e2e"]["value"], d["rs"]["value"] / 1e6, d["rs"]["e2e"]["value"] / 1e6, d["parity_mismatches"],
    c["ms_total"], c["ms_total_without_gather"], c["viterbi_gbit_per_s"]))
print(d["roofline"]["frac"], d["roofline"]["alu_pipe_frac"], d["extra"]["msc"]["roofline_issue_frac"], d["extra"]["fic_two_streams"])
PY
