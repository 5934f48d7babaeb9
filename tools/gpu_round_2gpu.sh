#!/bin/bash
# Two GPUs of one box: the multi-device tests (C-ABI sharded search in peer and NCCL mode, one-process-per-GPU NCCL
# search) and the bench under torchrun with in-bench parity checks of the merged results.
tag=${1:-r02g}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/${tag}_gpus.log
timeout 600 python -m pytest tests/test_sharded_gpu.py -q --timeout 500 -p no:cacheprovider -rs 2>&1 | tail -8 | tee gpurun_out/${tag}_sharded_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --knn-rows 30000000 --ivf-rows 30000000 --e2e-rows 20000000 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
echo "bench N=2 rc=$?"; tail -c 800 gpurun_out/${tag}_bench_n2.err; python - <<'P'
import json
j=json.load(open('gpurun_out/r02g_bench_n2.json'))
print('vitl14', j['value'], j['parity_checked'], j.get('parity_checked_all'))
for k in ('knn','ivf','e2e_query'):
    d=j[k]; print(k, round(d['value'],1), d['unit'], 'parity', d['parity_checked'], {a:b for a,b in d['parity'].items() if 'merge' in a}, 'nq1 ms', d.get('single_query_ms'), d.get('p50_ms'))
P
