# Round 2: wx_render's long-tile kernel restricted to the tile rows of chunks >= 2 (the first two chunks leave before it can finish).
mkdir -p gpurun_out; out=gpurun_out/r2_e2e_long2.txt; : > $out
( timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3 ) >> $out
for cfg in "WX_LONG_FIRST=1" "WX_LONG_FIRST=0"; do
  ( echo -n "sphere $cfg: "; env $cfg timeout 120 python tools/e2e_probe.py 2>&1 | tail -1 ) >> $out
  ( echo "fog $cfg:"; env $cfg timeout 300 python tools/fog_bench.py 2>&1 | grep '"mode": 0' | cut -c1-260 ) >> $out
done
cat $out
