for v in "" level3.tree2 level4 level5; do echo "== M3D_TAIL_SWITCH=$v"; M3D_TAIL_SWITCH=$v timeout 45 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t); print(d['value'], d['ms_per_step'], d['e2e']['value'])
except Exception as e: print('FAILED/timeout', len(t))"; done
