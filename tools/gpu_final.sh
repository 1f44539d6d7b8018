#!/bin/bash
# Final round of a build: whole GPU suite (no -x), A/B of the Sigma-update placement at N = 1024, the default bench line, in-graph timestamps.
tag=${1:-r02t}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
bash tools/gpu_ab.sh ${tag} EQVIO_SIGMA_AFTER_LIFT "1 0" 1024
timeout 900 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${tag}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_default.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches') if k in d}, d['e2e'], d['roofline'].get('frac'), d.get('cpu_baseline'))
for k,v in d.get('configs',{}).items(): print(k, v.get('value'), v.get('e2e',{}).get('value'), v.get('roofline',{}).get('frac'))
for k in ('fastRiccati','churn'):
    if k in d: print(k, d[k].get('value'), d[k].get('e2e',{}).get('value'))
"
timeout 200 python tools/graph_stamps.py --features 512 > gpurun_out/${tag}_update_stamps_n512.txt 2>&1; tail -16 gpurun_out/${tag}_update_stamps_n512.txt
