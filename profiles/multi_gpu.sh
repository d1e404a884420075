# usage (under gpurun --gpus N):  bash profiles/multi_gpu.sh N
# weak-scaling bench line (FIC batch per GPU) and the strong-scaling configs[4] end-to-end job at N GPUs
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -c 300 gpurun_out/bench_${N}gpu.err
$TR --master-port 29512 tests/full_size/e2e_scaling.py > gpurun_out/e2e_${N}gpu.json 2> gpurun_out/e2e_${N}gpu.err
tail -c 300 gpurun_out/e2e_${N}gpu.err
python - <<PY
import json
for f in ("gpurun_out/bench_${N}gpu.json", "gpurun_out/e2e_${N}gpu.json"):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        if "metric" in d:
            print("bench N=%d: %.1f Gbit/s (%.3f ms/step), e2e %.2f Gbit/s, rs %.1f M sf/s, gather %s ms" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["rs"]["value"] / 1e6, d["gather_ms"]))
        else:
            print("e2e N=%d: %.1f ms, %.1f M frames/s, %.2f M sf/s, %.1f Gbit/s, gather %s ms, wrong %d" % (d["n_gpus"], d["ms_total"], d["frames_per_s"] / 1e6, d["superframes_per_s"] / 1e6, d["viterbi_gbit_per_s"], d["gather_ms"], d["rs_accepted_but_wrong"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
