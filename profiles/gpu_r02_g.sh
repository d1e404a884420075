#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_viterbi.py -m gpu -x -q -k "punctured or smoke or golden" > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02g_pytest.log
sed -n '/^python - > gpurun_out\/r02f_punct.json/,/^PY$/p' profiles/gpu_r02_f.sh | sed 's/r02f_punct/r02g_punct/g' > /tmp/punct.sh; bash /tmp/punct.sh
cat gpurun_out/r02g_punct.json; tail -3 gpurun_out/r02g_punct.err
