# Round 2, first GPU call: baseline on this round's box + the CTA-queue kernel that had never run.
#   gpurun --timeout 420 -- "bash tools/gpu/r2_first.sh"
mkdir -p gpurun_out
( timeout 60 python tools/prof_run.py --frames 24 2>&1 | tail -1 | cut -c1-200 ) > gpurun_out/r2_first.txt
( WX_KERNEL=persistent timeout 60 python tools/prof_run.py --frames 24 2>&1 | tail -1 | cut -c1-200 ) >> gpurun_out/r2_first.txt
( WX_KERNEL=persistent_cta timeout 60 python tools/prof_run.py --frames 24 2>&1 | tail -1 | cut -c1-200 ) >> gpurun_out/r2_first.txt
for m in 3 4; do
  ( timeout 60 python tools/prof_run.py --frames 12 --mode $m 2>&1 | tail -1 | cut -c1-200 ) >> gpurun_out/r2_first.txt
  ( WX_KERNEL=persistent_cta timeout 60 python tools/prof_run.py --frames 12 --mode $m 2>&1 | tail -1 | cut -c1-200 ) >> gpurun_out/r2_first.txt
done
( WX_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k cta_queue 2>&1 | tail -3 ) >> gpurun_out/r2_first.txt
cat gpurun_out/r2_first.txt
