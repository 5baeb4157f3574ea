# A/B of build variants on the bench frame:  gpurun --timeout 200 -- "bash tools/gpu/ab.sh NAME1 NAME2 ..."
# (libraries woxel_b200/libwoxel_b200_NAME.so built beforehand with make -C woxel_b200/csrc EXTRA=... OUT=...; "default" = the shipped one)
mkdir -p gpurun_out
: > gpurun_out/ab.txt
for n in "$@"; do
  if [ "$n" = default ]; then unset WOXEL_B200_LIB; else export WOXEL_B200_LIB=$PWD/woxel_b200/libwoxel_b200_$n.so; fi
  ( timeout 60 python tools/prof_run.py --frames 32 2>&1 | tail -1 | cut -c1-120 ) >> gpurun_out/ab.txt
done
for n in "$@"; do
  [ "$n" = default ] && continue
  export WOXEL_B200_LIB=$PWD/woxel_b200/libwoxel_b200_$n.so
  ( echo "$n: $(timeout 100 python -m pytest tests/test_parity_gpu.py -x -q -k 'assets_all_modes or synthetic or edge_cases' 2>&1 | tail -1)" ) >> gpurun_out/ab.txt
done
cat gpurun_out/ab.txt
