timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k dla102 2>&1 | grep -E "AssertionError|assert " | head
timeout 900 python -m pytest tests/test_teacher_forced_gpu.py -m gpu -x -q 2>&1 | tail -6
