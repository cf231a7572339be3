for v in 0 1; do echo "== M3D_SIDE_PDL=$v (1 = PDL kept on beside branches)"; M3D_SIDE_PDL=$v timeout 90 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t); print(d['value'], d['ms_per_step'], d['e2e']['value'])
except Exception as e: print('FAILED/timeout', len(t))"; done
echo "== M3D_SIDE=0"; M3D_SIDE=0 timeout 90 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
