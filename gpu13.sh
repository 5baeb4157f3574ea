python tools/config_bench.py 2> gpurun_out/config_bench.err | tee gpurun_out/config_bench.jsonl
tail -3 gpurun_out/config_bench.err
