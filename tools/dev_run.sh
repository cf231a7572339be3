timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k dla102 2>&1 | grep -E "AssertionError|assert |Error" | head -8
