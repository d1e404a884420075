#!/bin/bash
# usage (under gpurun --gpus N): bash profiles/gpu_r02_multi.sh N [extra]
N=$1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02m_${N}_smi.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_rs.py -m gpu -x -q > gpurun_out/r02m_${N}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02m_${N}_pytest.log
g++ -std=c++17 -O2 -pthread -o /tmp/mdc tests/host/multi_device_check.cpp -ldl
REF=oracle/_ref/libviterbi_ref_avx512.so; grep -q avx512vl /proc/cpuinfo || REF=oracle/_ref/libviterbi_ref_avx2.so
timeout 600 /tmp/mdc viterbi.dll_b200/libviterbi_b200.so $REF 65536 > gpurun_out/r02m_${N}_native.json 2> gpurun_out/r02m_${N}_native.err; echo "native rc=$?"
cat gpurun_out/r02m_${N}_native.json; tail -3 gpurun_out/r02m_${N}_native.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02m_${N}_bench.json 2> gpurun_out/r02m_${N}_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02m_${N}_bench.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02m_${N}_bench.json").read().splitlines() if l.startswith("{")][-1])
    c = d["extra"]["configs4"]
    print("N=%d value %.1f e2e %.2f rs %.1fM rs_e2e %.1fM parity %d | configs4 %.1f ms (no gather %.1f) %.1f Gbit/s" % (
        d["n_gpus"], d["value"], d["e2e"]["value"], d["rs"]["value"] / 1e6, d["rs"]["e2e"]["value"] / 1e6, d["parity_mismatches"],
        c["ms_total"], c["ms_total_without_gather"], c["viterbi_gbit_per_s"]))
except Exception as e:
    print("bench parse failed", e)
PY
