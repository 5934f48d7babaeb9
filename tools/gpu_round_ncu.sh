#!/bin/bash
tag=${1:-r02n}
mkdir -p gpurun_out
export B200_GRAPHS=0
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_pair_kernel -s 60 -c 4 -o gpurun_out/${tag}_gemm python tools/ncu_targets.py embed > gpurun_out/${tag}_ncu1.log 2>&1; tail -1 gpurun_out/${tag}_ncu1.log
timeout 400 ncu --set full --clock-control none -k regex:attention_tc2_kernel -s 10 -c 1 -o gpurun_out/${tag}_attn python tools/ncu_targets.py embed > gpurun_out/${tag}_ncu2.log 2>&1; tail -1 gpurun_out/${tag}_ncu2.log
timeout 400 ncu --set full --clock-control none -k regex:"flat_scan_staged_kernel|scan_mma_kernel" -s 2 -c 5 -o gpurun_out/${tag}_scan python tools/ncu_targets.py knn > gpurun_out/${tag}_ncu3.log 2>&1; tail -1 gpurun_out/${tag}_ncu3.log
timeout 400 ncu --set full --clock-control none -k regex:ivf_scan_kernel -s 4 -c 2 -o gpurun_out/${tag}_ivf python tools/ncu_targets.py ivf > gpurun_out/${tag}_ncu4.log 2>&1; tail -1 gpurun_out/${tag}_ncu4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 1200 --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --workload vitl14 --steps 1 --warmup 3 --no-cpu --no-verify > gpurun_out/${tag}_bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/${tag}_bench_under_ncu.log
