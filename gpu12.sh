timeout 600 python -m pytest tests/test_sdf_gpu.py -m gpu -x -q 2>&1 | tail -4
python tools/sdf_bench.py 2>&1 | tail -2 | tee gpurun_out/sdf_bench.txt
