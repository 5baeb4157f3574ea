set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in default mb9 mb10 u2 u2mb9; do
  if [ $v = default ]; then unset WOXEL_B200_LIB; else export WOXEL_B200_LIB=$PWD/build/libwx_$v.so; fi
  python tools/prof_run.py --frames 12 2>&1 | tail -1
done 2>&1 | tee gpurun_out/variants_b.txt
