#!/bin/bash
tag=${1:-r02k}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_embed_gpu.py -q -k "tcgen05 or attention" -p no:cacheprovider 2>&1 | tail -2 | tee gpurun_out/${tag}_attn_tests.log
timeout 120 python tools/attn_time.py 1024 2>&1 | tee gpurun_out/${tag}_attn_time.log
timeout 120 python tools/attn_time.py 1024 256 2>&1 | tee -a gpurun_out/${tag}_attn_time.log
timeout 200 python tools/chunk_sweep.py 1024 2>&1 | tee gpurun_out/${tag}_ab.log
timeout 500 python -m pytest tests/test_embed_gpu.py tests/test_embed_batch_gpu.py tests/test_reference_plumbing_gpu.py -q -p no:cacheprovider -k "not tcgen05" 2>&1 | tail -3 | tee -a gpurun_out/${tag}_ab.log
timeout 300 python bench.py --workload plumbing 2>gpurun_out/${tag}_plumbing.err | tee gpurun_out/${tag}_plumbing.json
