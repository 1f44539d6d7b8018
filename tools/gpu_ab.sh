#!/bin/bash
# A/B of one environment switch on the headline bench (device-timed value, graphs on).  Usage: tools/gpu_ab.sh <tag> VAR "<values>" [features]
tag=$1; var=$2; vals=$3; N=${4:-512}
mkdir -p gpurun_out
for v in $vals; do
  env $var=$v timeout 300 python bench.py --features $N --no-sub-configs --no-cpu-baseline > gpurun_out/${tag}_${var}_$v.json 2> gpurun_out/${tag}_${var}_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_${var}_$v.json")); r=d["roofline"]
    print("$var=$v N=$N value %.1f e2e %.1f ms/period %.3f avg launch ms %.4f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"]))
except Exception as e:
    print("$var=$v failed", e); print(open("gpurun_out/${tag}_${var}_$v.err").read()[-800:])
PY
done
