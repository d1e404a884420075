#!/bin/bash
# usage (under gpurun --gpus N): bash profiles/gpu_r02_benchN.sh N   -- the full bench line at N GPUs (torchrun, as the driver launches it)
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29518 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02r_bench$N.json 2> gpurun_out/r02r_bench$N.err; echo "bench $N rc=$?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02r_bench$N.json").read().splitlines() if l.startswith("{")][-1])
    c = d["extra"]["configs4"]
    print("N=%d value %.1f e2e %.2f (punct %.2f) rs %.1fM rs_e2e %.1fM parity %d | configs4 %.1f ms (no gather %.1f) %.1f Gbit/s rounds %s" % (
        d["n_gpus"], d["value"], d["e2e"]["value"], d["e2e"]["punctured_input"]["value"], d["rs"]["value"] / 1e6, d["rs"]["e2e"]["value"] / 1e6, d["parity_mismatches"],
        c["ms_total"], c["ms_total_without_gather"], c["viterbi_gbit_per_s"], c["round_superframes"]))
except Exception as e:
    print("bench parse failed", e)
PY
tail -c 400 gpurun_out/r02r_bench$N.err | grep -v OMP | tail -3
