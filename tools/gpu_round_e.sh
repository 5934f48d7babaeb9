#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_embed_gpu.py -q -k tcgen05 -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/${tag}_attn_tests.log
timeout 120 python tools/attn_time.py > gpurun_out/${tag}_attn_time.log 2>&1; cat gpurun_out/${tag}_attn_time.log
for cfg in "2 0" "3 0" "2 1"; do
  set -- $cfg
  echo "== ATTN_GEN=$1 FUSE_LN=$2 (TMA store default on)" | tee -a gpurun_out/${tag}_ab.log
  B200_ATTN_GEN=$1 B200_FUSE_LN=$2 timeout 200 python tools/chunk_sweep.py 1024 2>&1 | tee -a gpurun_out/${tag}_ab.log
done
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log; tail -6 gpurun_out/${tag}_pytest.log
