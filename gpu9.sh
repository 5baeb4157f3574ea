python -m pytest tests -m gpu -x -q 2>&1 | tail -4
WX_KERNEL=tiled python tools/prof_run.py --frames 12 2>&1 | tail -1
for v in default p8; do
  if [ $v = default ]; then unset WOXEL_B200_LIB; else export WOXEL_B200_LIB=$PWD/build/libwx_$v.so; fi
  python tools/prof_run.py --frames 12 2>&1 | tail -1
done 2>&1 | tee gpurun_out/variants_e.txt
unset WOXEL_B200_LIB
for k in 1 4 8; do WX_RENDER_CHUNKS=$k python tools/e2e_probe.py 2>&1 | tail -1; done
