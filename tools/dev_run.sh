timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 20 --no-cpu-baseline --no-extras 2>gpurun_out/r02_bench_2gpu.err | tail -1 > gpurun_out/r02_bench_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_2gpu.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'])"
