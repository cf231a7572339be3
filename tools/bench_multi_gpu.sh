# 8- and 4-GPU bench lines of a round (run under: gpurun --gpus 8 -- bash tools/bench_multi_gpu.sh); outputs under gpurun_out/
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02_bench_8gpu.err | tail -1 > gpurun_out/r02_bench_8gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_8gpu.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d.get('sustained',{}).get('value'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>gpurun_out/r02_bench_4gpu.err | tail -1 > gpurun_out/r02_bench_4gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_4gpu.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'])"
