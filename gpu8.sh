set -x
nvidia-smi --query-gpu=index,name --format=csv
python -m pytest tests -m gpu -x -q -k "multi_device or pipelined" 2>&1 | tail -5
python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; cat gpurun_out/scale_n1.json | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err; tail -5 gpurun_out/scale_n2.err; cat gpurun_out/scale_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
