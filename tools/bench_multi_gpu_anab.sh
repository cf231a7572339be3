# 8-GPU bench line of BASELINE configs[2] (ANAB) (run under: gpurun --gpus 8 -- bash tools/bench_multi_gpu_anab.sh)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline --no-extras --attention ANAB 2>gpurun_out/r02_bench_anab_8gpu.err | tail -1 > gpurun_out/r02_bench_anab_8gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_anab_8gpu.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'])"
