timeout 300 python bench.py --backbone dla102 --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02_bench_dla102.err | tail -1 > gpurun_out/r02_bench_dla102.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_dla102.json')); print(d['config']['workload'][:80]); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']); print(d['roofline']['kernel'], d['roofline']['frac'])"
tail -2 gpurun_out/r02_bench_dla102.err
