timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for m in 0 3 4; do python tools/prof_run.py --frames 10 --mode $m 2>&1 | tail -1; done
for m in 3; do python tools/prof_run.py --frames 10 --mode $m --scene cube 2>&1 | tail -1; done
