for v in 0 1; do echo "== M3D_SIDE=$v"; M3D_SIDE=$v timeout 300 python tools/gpu_profile.py 2>&1 | grep -E "^stage=(forward|detect) graph=True"; done
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02_bench_dev.err | tail -1 > gpurun_out/r02_bench_dev.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_dev.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['sustained']['value'])"
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras --attention ANAB 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('anab', d['value'], d['e2e']['value'])"
