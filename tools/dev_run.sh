timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_8gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('sustained',{}).get('value'), d['clocks'])"; tail -2 gpurun_out/r02_bench_8gpu.err
