set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2>gpurun_out/bench_under_ncu.err
tail -2 gpurun_out/bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 2 -c 1 -f -o gpurun_out/prof_r1_v1 python tools/prof_run.py 2>&1 | tail -5
ls -la gpurun_out
