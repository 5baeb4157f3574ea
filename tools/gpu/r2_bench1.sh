# Round 2: new GPU tests + the full bench line on one GPU.
mkdir -p gpurun_out
( free -g | head -2; nproc ) > gpurun_out/r2_box.txt 2>&1
( timeout 900 python -m pytest tests/test_round2_gpu.py -x -q 2>&1 | tail -15 ) > gpurun_out/r2_tests_round2.txt
cat gpurun_out/r2_tests_round2.txt
timeout 400 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; tail -3 gpurun_out/r2_bench1.err; cut -c1-1500 gpurun_out/r2_bench1.json
cat gpurun_out/r2_box.txt
