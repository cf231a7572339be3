timeout 900 python bench.py --mode train --steps 10 --warmup 3 2>gpurun_out/r02_train.err | tail -1 > gpurun_out/r02_bench_train.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_train.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches']); print(json.dumps(d['train']))"
tail -3 gpurun_out/r02_train.err
