# Development run on the GPU box (scratch command file shipped with the snapshot).
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -rf 2>&1 | tail -40 > gpurun_out/r02a_gpu_tests.log; tail -25 gpurun_out/r02a_gpu_tests.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== bench"; timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; head -c 600 gpurun_out/r02a_bench.json; echo
echo "=== per-op default"; timeout 200 python tools/gpu_profile.py > gpurun_out/r02a_ops_default.txt 2>&1; tail -4 gpurun_out/r02a_ops_default.txt
echo "=== per-op PAIR256=0"; M3D_PAIR256=0 timeout 200 python tools/gpu_profile.py > gpurun_out/r02a_ops_nopair256.txt 2>&1; tail -4 gpurun_out/r02a_ops_nopair256.txt
for v in headbias epipipe; do
  echo "=== variant $v"
  M3D_LIB=$PWD/m3dssd_b200/libm3dssd_b200.$v.so timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_ops_gpu.py -m gpu -q 2>&1 | tail -3
  M3D_LIB=$PWD/m3dssd_b200/libm3dssd_b200.$v.so timeout 200 python tools/gpu_profile.py > gpurun_out/r02a_ops_$v.txt 2>&1; tail -4 gpurun_out/r02a_ops_$v.txt
done
echo "=== bench ANAB"; timeout 300 python bench.py --steps 30 --warmup 5 --attention ANAB --no-cpu-baseline > gpurun_out/r02a_bench_anab.json 2> gpurun_out/r02a_bench_anab.err; head -c 300 gpurun_out/r02a_bench_anab.json; echo
