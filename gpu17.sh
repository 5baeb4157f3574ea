ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 2 -c 1 -f -o gpurun_out/prof_r1_v22 python tools/prof_run.py 2>&1 | tail -2
