#!/bin/bash
# fused int8 Riccati step: its own tests under a timeout (in-kernel waits), then parity subset and a headline-only bench.  Usage: tools/gpu_fused_round.sh <tag>
tag=${1:-fz}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ozaki.py -q -x -k "fused or int8_cov" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_filter.py -q -x -k "single_step_parity or headline_N512 or graph_replay or golden or full_size" 2>&1 | grep -v "^$" | tail -8
for F in 1 0; do
  EQVIO_OZAKI_FUSED=$F timeout 300 python bench.py --no-sub-configs --no-cpu-baseline > gpurun_out/${tag}_bench_f$F.json 2> gpurun_out/${tag}_bench_f$F.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench_f$F.json"))
    r=d["roofline"]
    print("FUSED=$F N512 value %.1f e2e %.1f ms/period %.3f avg launch ms %.4f frac %.3f by_class %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], {k:(round(v["ms_per_period"],3)) for k,v in r["by_class"].items()}))
except Exception as e:
    print("bench FUSED=$F failed", e); print(open("gpurun_out/${tag}_bench_f$F.err").read()[-1500:])
PY
done
