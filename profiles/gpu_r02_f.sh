#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02f_pytest.log
# punctured device-resident FIC: fused fetch vs unpunctured
python - > gpurun_out/r02f_punct.json 2> gpurun_out/r02f_punct.err <<PY
import json, sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
import viterbi_dll_b200 as vb
from viterbi_dll_b200 import dabgen
assert vb.initialize()
res = {}
for f, n in ((768, 65536), (3072, 262144)):
    sym, _ = dabgen.make_frames_torch(n, f, 3.0, seed=5, device="cuda")
    keep = dabgen.fic_puncture_pattern() if f == 768 else dabgen.puncture_pattern(f, [(f // 32, 8)])
    kidx = torch.from_numpy(np.flatnonzero(keep)).cuda()
    rx = sym.index_select(1, kidx).contiguous()
    out = torch.empty((n, f // 8), dtype=torch.uint8, device="cuda")
    def timeit(fn, reps=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    t_plain = timeit(lambda: vb.deconvolve_batch_device(f, sym, out))
    t_punct = timeit(lambda: vb.deconvolve_batch_punctured_device(f, rx, keep, 128, out))
    res["F%d" % f] = {"frames": n, "kept_per_frame": int(keep.sum()), "plain_ms": t_plain, "punctured_fused_ms": t_punct,
                      "plain_gbps": n * f / t_plain / 1e6, "punctured_gbps": n * f / t_punct / 1e6, "ratio": t_punct / t_plain}
print(json.dumps(res))
PY
cat gpurun_out/r02f_punct.json; tail -3 gpurun_out/r02f_punct.err
for S in 1 2; do
BENCH_C4_STREAMS=$S timeout 600 python bench.py --steps 5 --warmup 3 --no-rs --no-e2e --no-cpu-baseline > gpurun_out/r02f_bench_s$S.json 2> gpurun_out/r02f_bench_s$S.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02f_bench_s$S.json").read().splitlines() if l.startswith("{")][-1])
c = d["extra"]["configs4"]
print("streams $S: configs4 %.1f ms (no gather %.1f) %.1f Gbit/s parity %d" % (c["ms_total"], c["ms_total_without_gather"], c["viterbi_gbit_per_s"], d["parity_mismatches"]))
PY
done
