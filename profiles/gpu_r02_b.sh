#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_viterbi.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02b_pytest.log
g++ -O2 -std=c++17 -pthread -o /tmp/latbench profiles/microbench/latbench.cpp -ldl
REF=oracle/_ref/libviterbi_ref_avx512.so; grep -q avx512vl /proc/cpuinfo || REF=oracle/_ref/libviterbi_ref_avx2.so
timeout 300 /tmp/latbench viterbi.dll_b200/libviterbi_b200.so $REF 2000 > gpurun_out/r02b_latbench.jsonl 2> gpurun_out/r02b_latbench.err; echo "latbench rc=$?"
cat gpurun_out/r02b_latbench.jsonl
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/chainlat profiles/microbench/chainlat.cu && /tmp/chainlat > gpurun_out/r02b_chainlat.jsonl; cat gpurun_out/r02b_chainlat.jsonl
