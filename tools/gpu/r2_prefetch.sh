# Round 2: leaf-brick prefetch on entering a leaf (-DWX_PREFETCH_LEAF) vs default: whole 4K frame, one rank's eighth, and the long-tile kernel alone (ncu).
mkdir -p gpurun_out; : > gpurun_out/r2_prefetch.txt
for n in default pf; do
  if [ "$n" = default ]; then unset WOXEL_B200_LIB; else export WOXEL_B200_LIB=$PWD/woxel_b200/libwoxel_b200_$n.so; fi
  echo "== $n" >> gpurun_out/r2_prefetch.txt
  ( timeout 60 python tools/prof_run.py --frames 24 2>&1 | tail -1 | cut -c1-140 ) >> gpurun_out/r2_prefetch.txt
  ( timeout 100 python tools/shard_probe.py --shards 1 8 2>&1 | tail -2 ) >> gpurun_out/r2_prefetch.txt
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:raycast --csv python tools/shard_probe.py --shards 8 --frames 6 2>/dev/null | grep -E 'raycast' | awk -F'","' '{print $5, $NF}' | tail -2 >> gpurun_out/r2_prefetch.txt
done
export WOXEL_B200_LIB=$PWD/woxel_b200/libwoxel_b200_pf.so
( echo "pf: $(timeout 100 python -m pytest tests/test_parity_gpu.py -x -q -k 'assets_all_modes or synthetic or edge_cases' 2>&1 | tail -1)" ) >> gpurun_out/r2_prefetch.txt
cat gpurun_out/r2_prefetch.txt
