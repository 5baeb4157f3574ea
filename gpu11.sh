timeout 600 python -m pytest tests/test_sdf_gpu.py tests/test_capture.py -m gpu -x -q 2>&1 | tail -15
