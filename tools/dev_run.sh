timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_writer_gpu.py -m gpu -x -q -k "preprocess or detect_images or kitti" 2>&1 | tail -30
