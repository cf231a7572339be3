timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python -m pytest tests/test_teacher_forced_gpu.py tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in 1 0; do echo "== M3D_NO_KSKIP=$v"; if [ $v = 1 ]; then export M3D_NO_KSKIP=1; else unset M3D_NO_KSKIP; fi; for i in 1 2; do timeout 100 python bench.py --steps 60 --warmup 20 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
except Exception as e: print('FAILED/timeout', len(t))"; done; done
