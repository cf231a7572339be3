timeout 600 python -m pytest tests/test_dcn_gpu.py tests/test_train_gpu.py -m gpu -q 2>&1 | tail -4
