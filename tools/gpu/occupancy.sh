# kernel time vs resident CTAs per SM (dynamic shared memory padding): 9 (default), 8, 7, 6, 5, 4
for pad in 0 28000 32000 37000 45000 56000; do
  WX_SMEM_PAD=$pad python tools/prof_run.py --frames 10 2>&1 | tail -1 | sed "s/^/pad $pad: /"
done | tee gpurun_out/occupancy.txt
