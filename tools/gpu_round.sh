#!/bin/bash
# One GPU-box session: new-kernel tests first (under a hard timeout), then the whole -m gpu suite, the bench, and the
# serving-shape launch list.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-r02}
mkdir -p gpurun_out
# new paths of this round are opt-in until verified: switch them on for this session, fall back one by one
export B200_ATTN_GEN=3 B200_FUSE_LN=1 B200_GRAPHS=1
timeout 600 python -m pytest tests/test_embed_gpu.py -q -k tcgen05 -p no:cacheprovider > gpurun_out/${tag}_attn.log 2>&1
rc=$?; echo "attention tests rc=$rc" | tee -a gpurun_out/${tag}_attn.log; tail -4 gpurun_out/${tag}_attn.log
if [ $rc -ne 0 ]; then export B200_ATTN_GEN=2; echo "falling back to attention_tc2 for the rest of this session"; fi
timeout 600 python -m pytest tests/test_embed_gpu.py -q -k "embeddings_match or vit_l14" -p no:cacheprovider > gpurun_out/${tag}_fuse.log 2>&1
rc=$?; echo "fused-LN embed tests rc=$rc" | tee -a gpurun_out/${tag}_fuse.log; tail -4 gpurun_out/${tag}_fuse.log
if [ $rc -ne 0 ]; then export B200_FUSE_LN=0; echo "falling back to the separate LayerNorm kernel for the rest of this session"; fi
timeout 400 python -m pytest tests/test_graphs_gpu.py -q -p no:cacheprovider > gpurun_out/${tag}_graphs.log 2>&1
rc=$?; echo "graph tests rc=$rc" | tee -a gpurun_out/${tag}_graphs.log; tail -4 gpurun_out/${tag}_graphs.log
if [ $rc -ne 0 ]; then export B200_GRAPHS=0; echo "falling back to eager launches for the rest of this session"; fi
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider --deselect tests/test_embed_gpu.py::test_tcgen05_attention_matches_fp32_reference > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log; tail -6 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/${tag}_bench.err; head -c 1200 gpurun_out/${tag}_bench.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_serve_launches.csv python tools/serve_shapes.py --reps 2 > gpurun_out/${tag}_serve.log 2>&1
timeout 120 python tools/serve_shapes.py --reps 20 > gpurun_out/${tag}_serve_plain.log 2>&1; cat gpurun_out/${tag}_serve_plain.log
timeout 400 python tools/chunk_sweep.py > gpurun_out/${tag}_chunk_sweep.log 2>&1; cat gpurun_out/${tag}_chunk_sweep.log
