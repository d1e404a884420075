#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_viterbi.py -m gpu -x -q -k "punctured" > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02h_pytest.log
sed -n '/^python - > gpurun_out\/r02f_punct.json/,/^PY$/p' profiles/gpu_r02_f.sh | sed 's/r02f_punct/r02h_punct_fused/g' > /tmp/punct.sh; bash /tmp/punct.sh; cat gpurun_out/r02h_punct_fused.json
sed 's/r02h_punct_fused/r02h_punct_separate/g' /tmp/punct.sh > /tmp/punct2.sh; VITERBI_B200_PUNCT_SEPARATE=1 bash /tmp/punct2.sh; cat gpurun_out/r02h_punct_separate.json
