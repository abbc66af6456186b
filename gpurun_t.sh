python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'])); print(d['phase_ms']); print(d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'])"
