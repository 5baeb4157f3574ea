# What the driver runs at round end, in one gpurun call:  gpurun --timeout 900 -- "bash tools/gpu/round_check.sh"
# (the full evidence run -- ncu capture, launch list, bench line, reference arm -- is tools/gpu/r2_final.sh TAG; scaling:
#  gpurun --gpus N -- "python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus N")
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | cut -c1-330
python bench.py > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; tail -2 gpurun_out/bench_check.err; cut -c1-250 gpurun_out/bench_check.json
