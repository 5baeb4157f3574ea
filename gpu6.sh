set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in default w2 w1 w2r48 w4r48; do
  if [ $v = default ]; then unset WOXEL_B200_LIB; else export WOXEL_B200_LIB=$PWD/build/libwx_$v.so; fi
  python tools/prof_run.py --frames 12 2>&1 | tail -1
done 2>&1 | tee gpurun_out/variants_c.txt
unset WOXEL_B200_LIB
python bench.py --steps 20 --warmup 3 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -3 gpurun_out/bench4.err; cat gpurun_out/bench4.json
