# Round-2 evidence in one call:  gpurun --timeout 900 -- "bash tools/gpu/r2_final.sh TAG"
# full GPU suite, ncu --set full of the raycast kernel (whole frame in one launch: long-tiles-first off), ncu launch list of a short
# bench run, the bench line and the reference arm.
TAG=${1:-r2_final}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
WX_LONG_FIRST=0 timeout 120 ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 2 -c 1 -f -o gpurun_out/prof_$TAG \
    python tools/prof_run.py 2>&1 | tail -2
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
timeout 300 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench.json
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | cut -c1-400 > gpurun_out/${TAG}_bench_reference.json; cat gpurun_out/${TAG}_bench_reference.json
