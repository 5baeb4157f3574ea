# Round 2: wx_render's chunked read-back with ONE long-tile kernel per frame (default) vs without (WX_LONG_FIRST=0): 4K frame through host buffers.
mkdir -p gpurun_out; : > gpurun_out/r2_e2e_long.txt
for cfg in "WX_LONG_FIRST=1" "WX_LONG_FIRST=0"; do
  ( echo -n "$cfg: "; env $cfg timeout 120 python tools/e2e_probe.py 2>&1 | tail -1 ) >> gpurun_out/r2_e2e_long.txt
done
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) >> gpurun_out/r2_e2e_long.txt
cat gpurun_out/r2_e2e_long.txt
