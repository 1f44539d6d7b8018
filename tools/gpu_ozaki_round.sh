#!/bin/bash
# parity + bench with the Riccati step on the int8 tensor cores.  Usage: tools/gpu_ozaki_round.sh <tag> "<slice list>"
tag=${1:-oz}; SL=${2:-"8"}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ozaki.py -q -x 2>&1 | tail -5
timeout 200 python tools/ozaki_try.py 2>&1 | grep "M=15\|M=30\|DMMA\|transB"
for S in $SL; do
  echo "== EQVIO_OZAKI=$S"
  EQVIO_OZAKI=$S python -m pytest tests/test_gpu_filter.py -q -k "single_step_parity or headline_N512 or config3 or graph_replay or conditioned_N64 or golden" 2>&1 | grep -v "^E  \|^$" | tail -8
  EQVIO_OZAKI=$S python bench.py --no-sub-configs --no-cpu-baseline > gpurun_out/${tag}_bench_oz$S.json 2> gpurun_out/${tag}_bench_oz$S.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_oz$S.json"))
r=d["roofline"]
print("OZAKI=$S N512 value %.1f e2e %.1f ms/period %.3f riccati avg launch ms %s by_class %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["avg_launch_ms"], {k:(round(v["ms_per_period"],3)) for k,v in r["by_class"].items()}))
PY
done
