# development run on the GPU box (tools/gpurun_retry.sh gpurun_out/x.log 900 'bash tools/dev_run.sh')
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_dcn_gpu.py tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -2
B='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["e2e"]["value"], d["ms_per_step"])'
echo "== bench"; timeout 200 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | python -c "$B"
timeout 100 python tools/probe_dcn_timeline.py 2>&1 | tail -27 | head -22
timeout 120 python tools/gpu_profile.py > /dev/null 2>&1; grep "dcn_fused" gpurun_out/profile_ops_bf16.txt; tail -3 gpurun_out/profile_ops_bf16.txt
