timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k side_stream 2>&1 | tail -8
