# End-of-round check in one short call:  gpurun --timeout 330 -- "bash tools/gpu/final_check.sh r1_v25"
# full GPU suite, one ncu --set full capture of the raycast kernel, one bench line (kept under profiles/ afterwards).
TAG=${1:-final}
mkdir -p gpurun_out
( timeout 170 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/fc_pytest_gpu.txt
cat gpurun_out/fc_pytest_gpu.txt
timeout 90 ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 2 -c 1 -f -o gpurun_out/prof_$TAG \
    python tools/prof_run.py 2>&1 | tail -2
timeout 130 python bench.py > gpurun_out/fc_bench.json 2> gpurun_out/fc_bench.err
tail -2 gpurun_out/fc_bench.err; cut -c1-400 gpurun_out/fc_bench.json
