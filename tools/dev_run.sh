for r in 1 2 3; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$r bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>gpurun_out/r02_bench_2gpu.err | tail -1 > gpurun_out/r02_bench_2gpu_$r.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_2gpu_$r.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'])"
grep -i "note\|ms/step" gpurun_out/r02_bench_2gpu.err | tail -4
done
