#!/bin/bash
# Profile pass of a round on the GPU box: tools/profile_round.sh r01d   (outputs under gpurun_out/, then copied to
# profiles/ with tools/summarize_ncu.py; see /opt/skills/guides/B200_PROFILING.md for the ncu flags)
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1; tail -2 gpurun_out/${tag}_gpu_tests.log
timeout 300 python bench.py --steps 50 --warmup 10 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; head -c 900 gpurun_out/${tag}_bench.json; echo
timeout 200 python tools/gpu_profile.py > /dev/null 2>&1; cp gpurun_out/profile_ops_bf16.txt gpurun_out/${tag}_per_op_cuda_events.txt; tail -4 gpurun_out/${tag}_per_op_cuda_events.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${tag}_launches.csv python tools/ncu_target.py --iters 2 --ops step > gpurun_out/ncu_a.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/${tag}_hot \
  python tools/ncu_target.py --iters 1 --ops stem,level0,level3.tree1.tree1.conv2,level5.tree1.conv2,dla_up.ida_1.node_1,dla_up.ida_1.node_1.offset,headsA.mlp,cls.l1,cls.l3,shape_align,flatten_heads \
  > gpurun_out/ncu_b.log 2>&1
ncu -i gpurun_out/${tag}_hot.ncu-rep --page raw --csv > gpurun_out/${tag}_hot_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_hot.ncu-rep
ls -la gpurun_out/${tag}_*
