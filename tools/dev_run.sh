M3D_DCN_HALO=0 timeout 120 python tools/probe_dcn_timeline.py | head -4
echo "=== bench"; timeout 900 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench.json'))
for k in ('value','ms_per_step','clocks','e2e','sustained','roofline','dcn','step_roofline','surface','cpu_baseline','gpu_launches'):
    print(k, json.dumps(d.get(k))[:700])
for n,k in list(d['kernels'].items())[:14]: print(n,k)
PY
tail -5 gpurun_out/r02h_bench.err
