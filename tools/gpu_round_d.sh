#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/${tag}_gemm_tests.log
for v in 0 1; do
  echo "== B200_GEMM_TMA_STORE=$v" | tee -a gpurun_out/${tag}_gemm.log
  B200_GEMM_TMA_STORE=$v timeout 200 python tools/gemm_sustained.py 2>&1 | grep TFLOP | tee -a gpurun_out/${tag}_gemm.log
done
for cfg in "2 0 0" "2 0 1" "2 1 1"; do
  set -- $cfg
  echo "== ATTN_GEN=$1 FUSE_LN=$2 TMA_STORE=$3" | tee -a gpurun_out/${tag}_ab.log
  B200_ATTN_GEN=$1 B200_FUSE_LN=$2 B200_GEMM_TMA_STORE=$3 timeout 200 python tools/chunk_sweep.py 1024 2>&1 | tee -a gpurun_out/${tag}_ab.log
done
B200_FUSE_LN=1 B200_GEMM_TMA_STORE=1 timeout 400 python -m pytest tests/test_embed_gpu.py tests/test_embed_batch_gpu.py -q -k "embeddings_match or vit_l14 or zero_feature or pair" -p no:cacheprovider 2>&1 | tail -4 | tee -a gpurun_out/${tag}_ab.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc3_kernel -s 2 -c 1 -o gpurun_out/${tag}_attn3 python tools/attn_time.py 256 > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log
