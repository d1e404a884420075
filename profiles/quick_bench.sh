python -m pytest tests/test_gpu_viterbi.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 --no-rs --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('FIC', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_int_alu']['frac'])"
python bench.py --steps 10 --warmup 3 --frames 262144 --framebits 3072 --no-rs --no-cpu-baseline --no-e2e 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MSC', d['value'], d['ms_per_step'], d['roofline_int_alu']['frac'])"
