#!/bin/bash
tag=${1:-r02g1}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_embed_gpu.py -q -k tcgen05 -p no:cacheprovider 2>&1 | tail -2 | tee gpurun_out/${tag}_attn_tests.log
echo "== two-pass" | tee gpurun_out/${tag}_attn_time.log
timeout 120 python tools/attn_time.py 2>&1 | tee -a gpurun_out/${tag}_attn_time.log
echo "== one-pass" | tee -a gpurun_out/${tag}_attn_time.log
B200_ATTN_ONEPASS=1 timeout 120 python tools/attn_time.py 2>&1 | tee -a gpurun_out/${tag}_attn_time.log
for cfg in "2" "3"; do
  echo "== ATTN_GEN=$cfg" | tee -a gpurun_out/${tag}_ab.log
  B200_ATTN_GEN=$cfg timeout 200 python tools/chunk_sweep.py 1024 2>&1 | tee -a gpurun_out/${tag}_ab.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc3_kernel -s 2 -c 1 -o gpurun_out/${tag}_attn3 python tools/attn_time.py 256 > gpurun_out/${tag}_ncu.log 2>&1; tail -1 gpurun_out/${tag}_ncu.log
