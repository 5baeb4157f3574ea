python tools/sdf_bench.py 2>&1 | tail -2 | tee gpurun_out/sdf_bench.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -3 gpurun_out/bench5.err; cat gpurun_out/bench5.json
