# ncu evidence for profiles/: the launch list of a short bench run and one full capture of the raycast kernel.
#   gpurun -- 'bash tools/gpu/profile.sh r1_v22'
TAG=${1:-prof}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 2 -c 1 -f -o gpurun_out/prof_$TAG \
    python tools/prof_run.py 2>&1 | tail -2
# then, here: python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep profiles/${TAG}_raycast_sphere2048_4k.json
