timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "== new"; timeout 300 python tools/probe_conv_perf.py s2 root cls proj
echo "== old"; M3D_LIB=$PWD/m3dssd_b200/libm3dssd_b200.old.so timeout 300 python tools/probe_conv_perf.py s2 root cls proj
timeout 300 python tools/gpu_profile.py > gpurun_out/r02z_ops.txt 2>&1; grep -E "^total|^stage" gpurun_out/r02z_ops.txt
