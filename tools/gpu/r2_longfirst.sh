# Round 2: long-tiles-first on/off on the 4K sphere, the GPU suite with it on, and a 1-GPU bench line.
mkdir -p gpurun_out
: > gpurun_out/r2_longfirst.txt
for lf in 1 0; do ( echo -n "WX_LONG_FIRST=$lf "; WX_LONG_FIRST=$lf timeout 60 python tools/prof_run.py --frames 32 2>&1 | tail -1 | cut -c1-110 ) >> gpurun_out/r2_longfirst.txt; done
for lf in 1 0; do for m in 3 4; do ( echo -n "WX_LONG_FIRST=$lf "; WX_LONG_FIRST=$lf timeout 60 python tools/prof_run.py --frames 16 --mode $m 2>&1 | tail -1 | cut -c1-110 ) >> gpurun_out/r2_longfirst.txt; done; done
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) >> gpurun_out/r2_longfirst.txt
cat gpurun_out/r2_longfirst.txt
