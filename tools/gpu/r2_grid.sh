# Round 2: world-grid march (default build) vs the round-1 march (-DWX_NO_GRID) and register caps; then the GPU parity suite.
mkdir -p gpurun_out
bash tools/gpu/ab.sh default nogrid mb10 mb8 > /dev/null 2>&1
cp gpurun_out/ab.txt gpurun_out/r2_grid_ab.txt
for m in 3 4; do ( timeout 60 python tools/prof_run.py --frames 12 --mode $m 2>&1 | tail -1 | cut -c1-120 ) >> gpurun_out/r2_grid_ab.txt; done
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) >> gpurun_out/r2_grid_ab.txt
cat gpurun_out/r2_grid_ab.txt
