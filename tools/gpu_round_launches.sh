#!/bin/bash
tag=${1:-r02t}
mkdir -p gpurun_out
timeout 190 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 1200 --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --workload vitl14 --steps 1 --warmup 3 --no-cpu --no-verify > gpurun_out/${tag}_bench_under_ncu.log 2>&1; tail -c 200 gpurun_out/${tag}_bench_under_ncu.log
