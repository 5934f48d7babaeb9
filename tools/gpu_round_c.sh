#!/bin/bash
tag=${1:-r02c}
mkdir -p gpurun_out
timeout 120 python tools/attn_time.py > gpurun_out/${tag}_attn_time.log 2>&1; cat gpurun_out/${tag}_attn_time.log
for cfg in "2 0" "3 0" "2 1" "3 1"; do
  set -- $cfg
  echo "== ATTN_GEN=$1 FUSE_LN=$2" | tee -a gpurun_out/${tag}_ab.log
  B200_ATTN_GEN=$1 B200_FUSE_LN=$2 timeout 200 python tools/chunk_sweep.py 1024 2>&1 | tee -a gpurun_out/${tag}_ab.log
done
B200_ATTN_GEN=3 B200_FUSE_LN=1 timeout 300 python -m pytest tests/test_embed_gpu.py -q -k "embeddings_match or vit_l14 or zero_feature" -p no:cacheprovider 2>&1 | tail -3 | tee -a gpurun_out/${tag}_ab.log
