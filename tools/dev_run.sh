timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "anab" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_teacher_forced_gpu.py tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -3
run() { timeout 100 python bench.py --steps 60 --warmup 20 --no-cpu-baseline --no-extras "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
except Exception as e: print('FAILED/timeout', len(t))"; }
echo "== old level1"; M3D_LEVEL1_2X2=1 run; M3D_LEVEL1_2X2=1 run
echo "== new level1"; run; run
echo "== ANAB"; run --attention ANAB; run --attention ANAB
