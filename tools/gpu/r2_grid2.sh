# Round 2: A/B of register caps on the world-grid march + ncu --set full capture of it.
mkdir -p gpurun_out
bash tools/gpu/ab.sh default nogrid mb10 mb8 > /dev/null 2>&1
cp gpurun_out/ab.txt gpurun_out/r2_grid_ab2.txt
timeout 120 ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_grid \
    python tools/prof_run.py 2>&1 | tail -2
cat gpurun_out/r2_grid_ab2.txt
