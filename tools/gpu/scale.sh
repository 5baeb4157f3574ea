# Weak scaling (bench.py under torchrun) + single-frame strong scaling on every GPU of the box:  gpurun --gpus N -- "bash tools/gpu/scale.sh"
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 300 python -m pytest tests -m gpu -x -q -k "multi_device" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; tail -2 gpurun_out/scale_n$N.err; cat gpurun_out/scale_n$N.json
python tools/multi_device_probe.py 2>&1 | tail -4 | tee gpurun_out/multi_device_n$N.txt
# BASELINE configs 4 and 5 through the multi-device context (one process): full-size fog, 64-camera orbit
python tools/fog_bench.py --devices $N 2>/dev/null | tee gpurun_out/fog2048_n$N.jsonl | grep '"render"' | cut -c1-260
python tools/multi_device_probe.py --orbit 2>&1 | tail -4 | tee gpurun_out/orbit64_n$N.txt
