timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k dla102 2>&1 | tail -40
