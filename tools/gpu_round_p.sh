#!/bin/bash
tag=${1:-r02p}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gemm_gpu.py -q -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/${tag}_gemm_tests.log
timeout 200 python tools/gemm_sustained.py --spec-ab --cublas 2>&1 | tee gpurun_out/${tag}_gemm_spec.log
B200_GEMM_DEBUG=1 timeout 100 python tools/gemm_wait.py 2>&1 | grep "gemm dbg" | awk 'NR%3==0' | tee gpurun_out/${tag}_gemm_wait.log
for cfg in "0 4 0" "1 4 0" "1 2 6"; do
  set -- $cfg
  echo "== B200_GEMM_SPEC=$1 B200_ATTN_GEN=$2 B200_ATTN_VARIANT=$3" | tee -a gpurun_out/${tag}_ab.log
  B200_GEMM_SPEC=$1 B200_ATTN_GEN=$2 B200_ATTN_VARIANT=$3 timeout 200 python tools/chunk_sweep.py 1024 2>&1 | tee -a gpurun_out/${tag}_ab.log
done
