# A short call for the last GPU minutes of a round:  gpurun --timeout 300 -- "bash tools/gpu/last_minutes.sh"
# 1. smoke on the library as built now; 2. A/B of the root-pointer variant (-DWX_ROOT_PTRS) against it on the bench frame;
# 3. a parity subset through the C ABI for both libraries.  Everything lands in gpurun_out/ as it is produced.
mkdir -p gpurun_out
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/lm_smoke.txt
cat gpurun_out/lm_smoke.txt
( timeout 90 python tools/prof_run.py --frames 24 2>&1 | tail -1 ) > gpurun_out/lm_ab_default.txt
( WOXEL_B200_LIB=$PWD/woxel_b200/libwoxel_b200_rootptr.so timeout 90 python tools/prof_run.py --frames 24 2>&1 | tail -1 ) > gpurun_out/lm_ab_rootptr.txt
cat gpurun_out/lm_ab_default.txt gpurun_out/lm_ab_rootptr.txt
( timeout 150 python -m pytest tests/test_parity_gpu.py -x -q -k "assets_all_modes or synthetic or edge_cases or max_steps or camera_batch" 2>&1 | tail -3 ) > gpurun_out/lm_parity_default.txt
cat gpurun_out/lm_parity_default.txt
( WOXEL_B200_LIB=$PWD/woxel_b200/libwoxel_b200_rootptr.so timeout 150 python -m pytest tests/test_parity_gpu.py -x -q -k "assets_all_modes or synthetic or edge_cases" 2>&1 | tail -3 ) > gpurun_out/lm_parity_rootptr.txt
cat gpurun_out/lm_parity_rootptr.txt
