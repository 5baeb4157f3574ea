# Round 2: BASELINE config 4 at its full 2048^3 against the oracle renderer (opt-in test: ~40 GB of host memory for the oracle's tables).
mkdir -p gpurun_out
( WX_TEST_FULL_FOG=1 timeout 1500 python -m pytest tests/test_round2_gpu.py -x -q -k "fog_2048" 2>&1 | tail -5; free -g | head -2 ) > gpurun_out/r2_config4_full.txt 2>&1
cat gpurun_out/r2_config4_full.txt
