#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_viterbi.py -m gpu -x -q -k "punctured or smoke or golden" > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02g_pytest.log
sed -n '/^python - > gpurun_out\/r02f_punct.json/,/^PY$/p' profiles/gpu_r02_f.sh | sed 's/r02f_punct/r02g_punct/g' > /tmp/punct.sh; bash /tmp/punct.sh
cat gpurun_out/r02g_punct.json; tail -3 gpurun_out/r02g_punct.err
timeout 600 python bench.py --only-configs4 > gpurun_out/r02g_c4.json 2> gpurun_out/r02g_c4.err; echo "c4 rc=$?"; tail -c 300 gpurun_out/r02g_c4.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02g_c4.json").read().splitlines() if l.startswith("{")][-1]); c = d["configs4"]
print("N=%d %.2f ms (no gather %.2f) %.1f Gbit/s rounds %s parity %d" % (d["n_gpus"], c["ms_total"], c["ms_total_without_gather"], c["viterbi_gbit_per_s"], c["round_superframes"], d["parity_mismatches"]))
PY
