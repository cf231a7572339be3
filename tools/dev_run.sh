# development run on the GPU box (tools/gpurun_retry.sh gpurun_out/x.log 900 'bash tools/dev_run.sh')
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -2
B='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["e2e"]["value"], d["ms_per_step"])'
echo "== bench f32 staged off"; M3D_F32_STAGED=0 timeout 200 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | python -c "$B"
for t in 4 6 8; do echo "== bench tail sms $t"; M3D_TAIL_SMS=$t timeout 200 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | python -c "$B"; done
timeout 120 python tools/gpu_profile.py > /dev/null 2>&1; grep "cls.l" gpurun_out/profile_ops_bf16.txt; tail -4 gpurun_out/profile_ops_bf16.txt
