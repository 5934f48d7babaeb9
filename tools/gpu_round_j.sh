#!/bin/bash
tag=${1:-r02j}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_embed_gpu.py -q -k "tcgen05 or attention" -p no:cacheprovider 2>&1 | tail -2 | tee gpurun_out/${tag}_attn_tests.log
timeout 120 python tools/attn_time.py 1024 2>&1 | tee gpurun_out/${tag}_attn_time.log
timeout 120 python tools/attn_time.py 1024 256 260 129 2>&1 | tee -a gpurun_out/${tag}_attn_time.log
echo "== tail kernel on (default)" | tee gpurun_out/${tag}_ab.log
timeout 200 python tools/chunk_sweep.py 1024 2>&1 | tee -a gpurun_out/${tag}_ab.log
echo "== tail kernel off" | tee -a gpurun_out/${tag}_ab.log
B200_ATTN_TAIL_KERNEL=0 timeout 200 python tools/chunk_sweep.py 1024 2>&1 | tee -a gpurun_out/${tag}_ab.log
echo "== tail kernel on (again)" | tee -a gpurun_out/${tag}_ab.log
timeout 200 python tools/chunk_sweep.py 1024 2>&1 | tee -a gpurun_out/${tag}_ab.log
timeout 500 python -m pytest tests/test_embed_gpu.py tests/test_embed_batch_gpu.py tests/test_graphs_gpu.py -q -p no:cacheprovider -k "not tcgen05" 2>&1 | tail -3 | tee -a gpurun_out/${tag}_ab.log
