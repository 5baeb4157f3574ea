# What the driver runs at round end, in one gpurun call:  gpurun -- "bash tools/gpu/round_check.sh"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | cut -c1-330
python bench.py > gpurun_out/bench6.json 2> gpurun_out/bench6.err; tail -2 gpurun_out/bench6.err; cut -c1-250 gpurun_out/bench6.json
