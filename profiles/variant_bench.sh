# A/B harness: bench every prebuilt library variant under gpurun_variants/ (FIC and MSC shapes)
for lib in gpurun_variants/*.so; do
  export VITERBI_B200_LIB=$PWD/$lib
  python - <<PY
import json, subprocess, sys, os
def run(args):
    out = subprocess.run([sys.executable, "bench.py", "--no-rs", "--no-cpu-baseline", "--no-e2e"] + args, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1]); return "%.1f Gbit/s (%.3f ms, FER %.4f)" % (d["value"], d["ms_per_step"], d["frame_error_rate"])
    except Exception as e:
        return "FAILED " + out.stderr[-300:]
print(os.environ["VITERBI_B200_LIB"].split("/")[-1], "FIC", run(["--steps", "30", "--warmup", "5"]), "| MSC", run(["--steps", "5", "--warmup", "3", "--frames", "262144", "--framebits", "3072"]))
PY
done
