# development run on the GPU box (tools/gpurun_retry.sh gpurun_out/x.log 900 'bash tools/dev_run.sh')
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 50 --warmup 10 > gpurun_out/r01d_bench_final.json 2> gpurun_out/r01d_bench_final.err; head -c 300 gpurun_out/r01d_bench_final.json; echo
timeout 200 python bench.py --steps 30 --warmup 5 --attention ANAB --no-cpu-baseline 2>/dev/null | tail -1 | head -c 200; echo
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
