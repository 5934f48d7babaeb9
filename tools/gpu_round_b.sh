#!/bin/bash
# Session b: attention_tc3 (leftover rows by the softmax group) correctness + timing, serving-shape latencies, the GEMM
# epilogue experiment (stores on/off), the tests touched since session a.
tag=${1:-r02b}
mkdir -p gpurun_out
export B200_ATTN_GEN=3 B200_GRAPHS=1
timeout 300 python -m pytest tests/test_embed_gpu.py -q -k tcgen05 -p no:cacheprovider > gpurun_out/${tag}_attn.log 2>&1
echo "attention tests rc=$?" | tee -a gpurun_out/${tag}_attn.log; tail -3 gpurun_out/${tag}_attn.log
timeout 120 python tools/attn_time.py > gpurun_out/${tag}_attn_time.log 2>&1; cat gpurun_out/${tag}_attn_time.log
timeout 900 python -m pytest tests/test_service_gpu.py tests/test_preprocess_gpu.py tests/test_gemm_gpu.py tests/test_graphs_gpu.py tests/test_ivf_gpu.py tests/test_knn_gpu.py tests/test_embed_gpu.py tests/test_sharded_gpu.py -q --timeout 600 -p no:cacheprovider --deselect tests/test_embed_gpu.py::test_tcgen05_attention_matches_fp32_reference > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log; tail -8 gpurun_out/${tag}_pytest.log
timeout 120 python tools/serve_shapes.py --reps 20 > gpurun_out/${tag}_serve_plain.log 2>&1; cat gpurun_out/${tag}_serve_plain.log
timeout 120 python tools/serve_shapes.py --reps 20 --arch open_clip:ViT-H-14 >> gpurun_out/${tag}_serve_plain.log 2>&1; tail -2 gpurun_out/${tag}_serve_plain.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_serve_launches.csv python tools/serve_shapes.py --reps 2 > gpurun_out/${tag}_serve.log 2>&1
B200_NVCC_EXTRA=-DB200_TIMING_EXPERIMENTS python clip-retrieval_b200/build.py -f > /dev/null 2>&1
timeout 300 bash tools/gemm_exp.sh > gpurun_out/${tag}_gemm_exp.log 2>&1; cat gpurun_out/${tag}_gemm_exp.log
python clip-retrieval_b200/build.py -f > /dev/null 2>&1
