timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_2gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('sustained',{}).get('value'))"; tail -3 gpurun_out/r02_bench_2gpu.err
