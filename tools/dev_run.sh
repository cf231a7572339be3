# development run on the GPU box, 4 GPUs
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 30 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r01d_bench_4gpu.json; head -c 260 gpurun_out/r01d_bench_4gpu.json; echo
