#!/bin/bash
# ncu evidence for round 2 (1 GPU).  Every ncu run follows a plain run of the same command in the same call.
mkdir -p gpurun_out
M="smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed_pipe_lsu.sum"
L="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-e2e"
$L > gpurun_out/ncu_plain_list.log 2>&1 && \
ncu --nvtx --nvtx-include "fec_timed_viterbi/" --nvtx-include "fec_timed_rs/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv $L > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
F="python bench.py --steps 3 --warmup 3 --no-rs --no-e2e --no-cpu-baseline --no-extra"
$F > gpurun_out/ncu_plain_fic.log 2>&1 && \
ncu --set full --metrics $M --clock-control none --import-source on -k regex:viterbi_pair -s 3 -c 1 -f -o gpurun_out/r02_fic $F > gpurun_out/ncu_fic.log 2>&1
echo "fic rc=$?"
MS="python bench.py --steps 3 --warmup 3 --frames 262144 --framebits 3072 --no-rs --no-e2e --no-cpu-baseline --no-extra"
$MS > gpurun_out/ncu_plain_msc.log 2>&1 && \
ncu --set full --metrics $M --clock-control none --import-source on -k regex:viterbi_pair -s 3 -c 1 -f -o gpurun_out/r02_msc $MS > gpurun_out/ncu_msc.log 2>&1
echo "msc rc=$?"
R="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extra"
$R > gpurun_out/ncu_plain_rs.log 2>&1 && \
ncu --set full --metrics $M --clock-control none --import-source on -k regex:rs_superframe -s 27 -c 1 -f -o gpurun_out/r02_rs_s4 $R > gpurun_out/ncu_rs.log 2>&1
echo "rs rc=$?"
W="python profiles/warp_probe.py"
$W > gpurun_out/ncu_plain_warp.log 2>&1 && \
ncu --set full --metrics $M --clock-control none --import-source on -k regex:viterbi_warp -s 2 -c 1 -f -o gpurun_out/r02_warp_1frame $W > gpurun_out/ncu_warp1.log 2>&1
echo "warp1 rc=$?"
ncu --set full --metrics $M --clock-control none --import-source on -k regex:viterbi_warp -s 5 -c 1 -f -o gpurun_out/r02_warp_2048 $W > gpurun_out/ncu_warp2.log 2>&1
echo "warp2 rc=$?"
ls -la gpurun_out/*.ncu-rep
