timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py --mode train --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); print(json.dumps(d['train']))"
