#!/bin/bash
# Profile pass of a round on the GPU box: tools/profile_round.sh r02   (outputs under gpurun_out/; the summaries are
# copied to profiles/ by hand after reading them; see /opt/skills/guides/B200_PROFILING.md for the ncu flags)
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rf > gpurun_out/${tag}_gpu_tests.log 2>&1; tail -3 gpurun_out/${tag}_gpu_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; head -c 600 gpurun_out/${tag}_bench.json; echo
timeout 600 python bench.py --steps 30 --warmup 5 --attention ANAB > gpurun_out/${tag}_bench_anab.json 2> gpurun_out/${tag}_bench_anab.err; head -c 400 gpurun_out/${tag}_bench_anab.json; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2>/dev/null; head -c 300 gpurun_out/${tag}_bench_reference_arm.json; echo
timeout 200 python tools/gpu_profile.py > /dev/null 2>&1; cp gpurun_out/profile_ops_bf16.txt gpurun_out/${tag}_per_op_cuda_events.txt; tail -4 gpurun_out/${tag}_per_op_cuda_events.txt
# launch list + DRAM traffic + tensor-pipe activity of EVERY launch of one warm step
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_step_metrics.csv python tools/ncu_target.py --iters 2 --ops step > gpurun_out/ncu_a.log 2>&1
python tools/summarize_ncu.py traffic gpurun_out/${tag}_step_metrics.csv > gpurun_out/${tag}_traffic.json; head -c 700 gpurun_out/${tag}_traffic.json
python tools/summarize_ncu.py launches gpurun_out/${tag}_step_metrics.csv > gpurun_out/${tag}_launch_summary.md
# full-set capture of the hot kernels
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/${tag}_hot \
  python tools/ncu_target.py --iters 1 --ops stem,level0,level3.tree1.tree1.conv2,level5.tree1.conv2,dla_up.ida_1.node_1,dla_up.ida_1.node_1.offset,headsA.mlp,cls.l1,cls.l3,shape_align,flatten_heads \
  > gpurun_out/ncu_b.log 2>&1
ncu -i gpurun_out/${tag}_hot.ncu-rep --page raw --csv > gpurun_out/${tag}_hot_raw.csv 2>/dev/null
python tools/summarize_ncu.py hot gpurun_out/${tag}_hot_raw.csv > gpurun_out/${tag}_hot_kernels.md
rm -f gpurun_out/${tag}_hot.ncu-rep
ls -la gpurun_out/${tag}_*
# BASELINE configs[3]: the training step (native kernels + the reference's loss in its static-shape form)
timeout 900 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/${tag}_bench_train.json 2> gpurun_out/${tag}_bench_train.err; head -c 300 gpurun_out/${tag}_bench_train.json; echo
