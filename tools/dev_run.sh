timeout 200 python tools/probe_elementwise.py 2>&1 | head -2
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k upsample 2>&1 | tail -2
