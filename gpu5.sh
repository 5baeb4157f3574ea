set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
python bench.py --steps 20 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -3 gpurun_out/bench3.err; cat gpurun_out/bench3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v21.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2>gpurun_out/bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 2 -c 1 -f -o gpurun_out/prof_r1_v21 python tools/prof_run.py 2>&1 | tail -3
