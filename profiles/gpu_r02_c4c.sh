#!/bin/bash
# usage (under gpurun --gpus N): bash profiles/gpu_r02_c4c.sh N   -- configs[4] gather modes: NCCL vs copy engines (dma) vs RS-kernel peer stores
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 $TR --master-port 29517 bench.py --gpus $N --only-configs4 > gpurun_out/r02t_c4_${N}_$name.json 2> gpurun_out/r02t_c4_${N}_$name.err
  echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02t_c4_${N}_$name.json").read().splitlines() if l.startswith("{")][-1])
    c = d["configs4"]
    print("$name: N=%d %.2f ms (no gather %.2f) %.1f Gbit/s parity %d | %s" % (d["n_gpus"], c["ms_total"], c["ms_total_without_gather"], c["viterbi_gbit_per_s"], d["parity_mismatches"], c["gather"][:60]))
except Exception as e:
    print("$name parse failed", e)
PY
  grep -i "error\|Traceback" gpurun_out/r02t_c4_${N}_$name.err | head -3
}
run dma4 BENCH_C4_GATHER=dma
run nccl BENCH_C4_GATHER=nccl

run dma7 BENCH_C4_GATHER=dma BENCH_C4_COPY_STREAMS=7
