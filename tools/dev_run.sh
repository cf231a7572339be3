timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_teacher_forced_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/gpu_profile.py > gpurun_out/r02z_ops.txt 2>&1; grep -E "^level1|^level0|^total|^stage" gpurun_out/r02z_ops.txt
