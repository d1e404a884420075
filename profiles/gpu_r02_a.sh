#!/bin/bash
# round 2, first GPU call: parity suite, drop-in latency, bench line, integer peaks, PCIe probe (1 GPU)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
g++ -O2 -std=c++17 -pthread -o /tmp/latbench profiles/microbench/latbench.cpp -ldl
REF=oracle/_ref/libviterbi_ref_avx512.so; grep -q avx512vl /proc/cpuinfo || REF=oracle/_ref/libviterbi_ref_avx2.so
timeout 300 /tmp/latbench viterbi.dll_b200/libviterbi_b200.so $REF 2000 > gpurun_out/r02a_latbench.jsonl 2> gpurun_out/r02a_latbench.err; echo "latbench rc=$?"
cat gpurun_out/r02a_latbench.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02a_bench.err
timeout 300 python profiles/int_peaks.py 4000 > gpurun_out/r02a_int_peaks.log 2>&1; echo "int_peaks rc=$?"; cp profiles/int_peaks.json gpurun_out/ 2>/dev/null
tail -2 gpurun_out/r02a_int_peaks.log
timeout 200 python profiles/pcie_probe.py > gpurun_out/r02a_pcie1.json 2>&1; cat gpurun_out/r02a_pcie1.json
