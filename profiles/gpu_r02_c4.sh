#!/bin/bash
# usage (under gpurun --gpus N): bash profiles/gpu_r02_c4.sh N   -- gather-mode A/B of the configs[4] chain + multi-GPU tests
N=$1
mkdir -p gpurun_out
if [ "$2" != "nopytest" ]; then timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02c4_${N}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02c4_${N}_pytest.log; fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 $TR --master-port 29517 bench.py --gpus $N --only-configs4 > gpurun_out/r02c4_${N}_$name.json 2> gpurun_out/r02c4_${N}_$name.err
  echo "$name rc=$?"; tail -c 200 gpurun_out/r02c4_${N}_$name.err | grep -v OMP | tail -2
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02c4_${N}_$name.json").read().splitlines() if l.startswith("{")][-1])
    c = d["configs4"]
    print("$name: N=%d %.2f ms (no gather %.2f) %.1f Gbit/s streams %s parity %d" % (d["n_gpus"], c["ms_total"], c["ms_total_without_gather"], c["viterbi_gbit_per_s"], c["compute_streams"], d["parity_mismatches"]))
except Exception as e:
    print("$name parse failed", e)
PY
}
run peer2 BENCH_C4_GATHER=peer
run peer1 BENCH_C4_GATHER=peer BENCH_C4_STREAMS=1
run nccl1 BENCH_C4_GATHER=nccl BENCH_C4_STREAMS=1
if [ "$2" != "nopytest" ]; then run nccl2 BENCH_C4_GATHER=nccl BENCH_C4_STREAMS=2; fi
