tag=r02f
timeout 900 python -m pytest tests -m gpu -x -q -rf > gpurun_out/${tag}_gpu_tests.log 2>&1; tail -2 gpurun_out/${tag}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; head -c 300 gpurun_out/${tag}_bench.json; echo
timeout 400 python bench.py --attention ANAB > gpurun_out/${tag}_bench_anab.json 2> gpurun_out/${tag}_bench_anab.err; head -c 300 gpurun_out/${tag}_bench_anab.json; echo
timeout 200 python tools/gpu_profile.py > /dev/null 2>&1; cp gpurun_out/profile_ops_bf16.txt gpurun_out/${tag}_per_op_cuda_events.txt; tail -2 gpurun_out/${tag}_per_op_cuda_events.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_step_metrics.csv python tools/ncu_target.py --iters 2 --ops step > gpurun_out/ncu_a.log 2>&1
python tools/summarize_ncu.py traffic gpurun_out/${tag}_step_metrics.csv > gpurun_out/${tag}_traffic.json; head -c 300 gpurun_out/${tag}_traffic.json; echo
python tools/summarize_ncu.py launches gpurun_out/${tag}_step_metrics.csv > gpurun_out/${tag}_launch_summary.md
