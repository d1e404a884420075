#!/bin/bash
# 8-GPU call: host-side H2D ceiling vs ranks (with / without NUMA affinity), native multi-device check, bench N=8 and N=4
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_8_topo.txt 2>&1
lscpu | head -30 > gpurun_out/r02_8_lscpu.txt
numactl -H >> gpurun_out/r02_8_lscpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
: > gpurun_out/r02_8_pcie.jsonl
for N in 1 2 4 8; do
  timeout 200 $TR --nproc-per-node $N --master-port 2953$N profiles/pcie_probe.py 2>/dev/null | grep '^{' >> gpurun_out/r02_8_pcie.jsonl
done
for N in 4 8; do
  timeout 200 $TR --nproc-per-node $N --master-port 2954$N profiles/pcie_probe.py --affinity 2>/dev/null | grep '^{' >> gpurun_out/r02_8_pcie.jsonl
done
cat gpurun_out/r02_8_pcie.jsonl
g++ -std=c++17 -O2 -pthread -o /tmp/mdc tests/host/multi_device_check.cpp -ldl
REF=oracle/_ref/libviterbi_ref_avx512.so; grep -q avx512vl /proc/cpuinfo || REF=oracle/_ref/libviterbi_ref_avx2.so
timeout 600 /tmp/mdc viterbi.dll_b200/libviterbi_b200.so $REF 65536 > gpurun_out/r02_8_native.json 2> gpurun_out/r02_8_native.err; echo "native rc=$?"
cat gpurun_out/r02_8_native.json
for N in 8 4; do
timeout 900 $TR --nproc-per-node $N --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_8_bench$N.json 2> gpurun_out/r02_8_bench$N.err; echo "bench $N rc=$?"
tail -c 300 gpurun_out/r02_8_bench$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_8_bench$N.json").read().splitlines() if l.startswith("{")][-1])
    c = d["extra"]["configs4"]
    print("N=%d value %.1f e2e %.2f (punct %.2f) rs %.1fM rs_e2e %.1fM parity %d | configs4 %.1f ms (no gather %.1f) %.1f Gbit/s" % (
        d["n_gpus"], d["value"], d["e2e"]["value"], d["e2e"]["punctured_input"]["value"], d["rs"]["value"] / 1e6, d["rs"]["e2e"]["value"] / 1e6, d["parity_mismatches"],
        c["ms_total"], c["ms_total_without_gather"], c["viterbi_gbit_per_s"]))
except Exception as e:
    print("bench parse failed", e)
PY
done
