#!/bin/bash
tag=${1:-r02l}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_embed_gpu.py -q -k "tcgen05 or attention" -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/${tag}_attn_tests.log
timeout 200 python tools/attn_variants.py 1024 2>&1 | tee gpurun_out/${tag}_attn_variants.log
for v in 0 5 7 1; do
  echo "== B200_ATTN_VARIANT=$v" | tee -a gpurun_out/${tag}_ab.log
  B200_ATTN_VARIANT=$v timeout 200 python tools/chunk_sweep.py 1024 2>&1 | tee -a gpurun_out/${tag}_ab.log
done
timeout 200 python tools/gemm_sustained.py --cublas 2>&1 | tee gpurun_out/${tag}_gemm_cublas.log
