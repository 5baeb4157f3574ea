mkdir -p gpurun_out; : > gpurun_out/r2_shardprobe.txt
for cfg in "WX_LONG_FIRST=0" "WX_LONG_THRESHOLD=96" "WX_LONG_THRESHOLD=64" "WX_LONG_THRESHOLD=48" "WX_LONG_THRESHOLD=128"; do
  echo "== $cfg" >> gpurun_out/r2_shardprobe.txt
  ( env $cfg timeout 120 python tools/shard_probe.py 2>&1 | tail -4 ) >> gpurun_out/r2_shardprobe.txt
done
cat gpurun_out/r2_shardprobe.txt
