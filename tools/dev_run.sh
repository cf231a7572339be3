timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -4
echo "=== new"; timeout 300 python tools/probe_conv_perf.py level2 level3
echo "=== old"; M3D_LIB=$PWD/m3dssd_b200/libm3dssd_b200.old.so timeout 300 python tools/probe_conv_perf.py level2 level3
timeout 300 python tools/gpu_profile.py > gpurun_out/r02z_ops.txt 2>&1; grep -E "level0|level2.tree|^total|^stage" gpurun_out/r02z_ops.txt
