# Development run on the GPU box: GPU parity tests + one bench line.
#   tools/gpurun_retry.sh gpurun_out/x.log 900 'bash tools/dev_run.sh'
# (edit freely: this is the scratch command file the retry wrapper ships with the snapshot)
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 30 --warmup 5 2>/dev/null | tail -1 | head -c 400; echo
