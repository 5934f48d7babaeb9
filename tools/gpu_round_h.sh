#!/bin/bash
tag=${1:-r02h}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_embed_gpu.py -q -k tcgen05 -p no:cacheprovider 2>&1 | tail -2 | tee gpurun_out/${tag}_attn_tests.log
timeout 120 python tools/attn_time.py 1024 2>&1 | tee gpurun_out/${tag}_attn_time.log
timeout 120 python tools/attn_time.py 1024 257 129 260 2>&1 | tee -a gpurun_out/${tag}_attn_time.log
for cfg in "2" "3"; do
  echo "== ATTN_GEN=$cfg" | tee -a gpurun_out/${tag}_ab.log
  B200_ATTN_GEN=$cfg timeout 200 python tools/chunk_sweep.py 1024 2>&1 | tee -a gpurun_out/${tag}_ab.log
done
B200_ATTN_GEN=3 timeout 600 python -m pytest tests/test_embed_gpu.py tests/test_embed_batch_gpu.py tests/test_graphs_gpu.py -q -p no:cacheprovider -k "not tcgen05" 2>&1 | tail -3 | tee -a gpurun_out/${tag}_ab.log
