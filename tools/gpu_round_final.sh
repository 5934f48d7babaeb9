#!/bin/bash
tag=${1:-r02z}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider --durations=12 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log; tail -22 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench_ref.json | head -c 600; echo
timeout 120 python tools/serve_shapes.py --reps 20 2>&1 | tee gpurun_out/${tag}_serve_plain.log
timeout 120 python tools/serve_shapes.py --reps 20 --arch open_clip:ViT-H-14 2>&1 | tee -a gpurun_out/${tag}_serve_plain.log
python - $tag <<'P'
import json
import sys
j=json.load(open('gpurun_out/%s_bench.json' % (sys.argv[1] if len(sys.argv) > 1 else 'r02z')))
print('vitl14', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), 'ms', round(j['ms_per_step'],1), 'roofline', round(j['roofline']['frac'],3), 'traffic', j['roofline']['traffic'], j['breakdown_ms_per_step'], 'parity', j['parity_checked'], j['clocks'], 'cpu', j['cpu_baseline']['value'])
pl=j.get('plumbing',{}); print('plumbing', pl.get('value'), pl.get('mapper_ms_per_step'), pl.get('parity'), pl.get('cpu_baseline',{}).get('value'))
for k in ('knn','ivf','e2e_query'):
    d=j[k]; print(k, round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'parity', d['parity_checked'], 'nq1', d.get('single_query_ms'), 'p50', d.get('p50_ms'), d.get('p99_ms'), 'cpu', round(d['cpu_baseline']['value'],2), 'wall', round(d['wall_s'],1))
P
