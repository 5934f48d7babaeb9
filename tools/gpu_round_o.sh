#!/bin/bash
tag=${1:-r02o}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_embed_gpu.py -q -k "tcgen05 or attention" -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/${tag}_attn_tests.log
timeout 200 python tools/attn_variants.py 1024 2>&1 | tee gpurun_out/${tag}_attn_variants.log
for cfg in "4 0" "2 6" "2 10" "2 2"; do
  set -- $cfg
  echo "== B200_ATTN_GEN=$1 B200_ATTN_VARIANT=$2" | tee -a gpurun_out/${tag}_ab.log
  B200_ATTN_GEN=$1 B200_ATTN_VARIANT=$2 timeout 200 python tools/chunk_sweep.py 1024 2>&1 | tee -a gpurun_out/${tag}_ab.log
done
B200_GEMM_DEBUG=1 timeout 100 python tools/gemm_wait.py 2>&1 | tee gpurun_out/${tag}_gemm_wait.log
