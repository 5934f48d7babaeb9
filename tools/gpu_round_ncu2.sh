#!/bin/bash
tag=${1:-r02s}
mkdir -p gpurun_out
export B200_GRAPHS=0
timeout 170 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_pair_kernel -s 60 -c 4 -o gpurun_out/${tag}_gemm python tools/ncu_targets.py embed > gpurun_out/${tag}_ncu1.log 2>&1; tail -1 gpurun_out/${tag}_ncu1.log
timeout 120 ncu --set full --clock-control none -k regex:attention_tc2_kernel -s 10 -c 1 -o gpurun_out/${tag}_attn python tools/ncu_targets.py embed > gpurun_out/${tag}_ncu2.log 2>&1; tail -1 gpurun_out/${tag}_ncu2.log
