#!/bin/bash
tag=${1:-r02q}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_graphs_gpu.py tests/test_service_gpu.py -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/${tag}_tests.log
timeout 200 python tools/serve_shapes.py --arch ViT-L/14 2>&1 | tail -6 | tee gpurun_out/${tag}_serve.log
timeout 200 python tools/serve_shapes.py --arch ViT-H/14 2>&1 | tail -6 | tee -a gpurun_out/${tag}_serve.log
B200_GEMM_SPEC=0 timeout 200 python tools/serve_shapes.py --arch ViT-H/14 2>&1 | tail -6 | tee -a gpurun_out/${tag}_serve.log
