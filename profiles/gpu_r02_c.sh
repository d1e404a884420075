#!/bin/bash
mkdir -p gpurun_out
g++ -O2 -std=c++17 -pthread -o /tmp/latbench profiles/microbench/latbench.cpp -ldl
REF=oracle/_ref/libviterbi_ref_avx512.so; grep -q avx512vl /proc/cpuinfo || REF=oracle/_ref/libviterbi_ref_avx2.so
timeout 300 /tmp/latbench viterbi.dll_b200/libviterbi_b200.so $REF 2000 > gpurun_out/r02c_latbench.jsonl 2> gpurun_out/r02c_latbench.err; echo "latbench rc=$?"
cat gpurun_out/r02c_latbench.jsonl
