timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "class_head or softmax" 2>&1 | tail -5
M3D_FUSE_CLS=1 timeout 900 python -m pytest tests/test_model_gpu.py tests/test_teacher_forced_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do echo "== M3D_FUSE_CLS=$v"; M3D_FUSE_CLS=$v timeout 60 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
except Exception as e: print('FAILED/timeout', len(t))"; done
