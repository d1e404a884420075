# A/B harness for RS variants: default bench (FIC Viterbi + RS mix, device-resident) per prebuilt library
for lib in gpurun_variants/*.so; do
  VITERBI_B200_LIB=$PWD/$lib python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', 'vit %.1f Gbit/s' % d['value'], '| rs %.1f M sf/s (%.3f ms)' % (d['rs']['value']/1e6, d['rs']['ms_per_step']))"
done
