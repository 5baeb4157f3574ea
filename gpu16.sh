timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in base pred default; do
  if [ $v = default ]; then unset WOXEL_B200_LIB; else export WOXEL_B200_LIB=$PWD/build/libwx_$v.so; fi
  python tools/prof_run.py --frames 12 2>&1 | tail -1
done 2>&1 | tee gpurun_out/variants_f.txt
