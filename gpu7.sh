for k in 1 2 4 8 16; do WX_RENDER_CHUNKS=$k python tools/e2e_probe.py 2>&1 | tail -1; done | tee gpurun_out/e2e_chunks.txt
