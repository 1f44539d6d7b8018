#!/bin/bash
# Where the int8 Riccati path pays: steps/s with the int8 path forced on (EQVIO_OZAKI_MIN_TILES=4) vs off (EQVIO_OZAKI=0) over a range of N.
# Usage: tools/gpu_oz_threshold.sh <tag> "<N list>"
tag=${1:-ozth}; NL=${2:-"192 256 320 384"}
mkdir -p gpurun_out
for N in $NL; do
  for S in 0 8; do
    K=20; [ $N -ge 1024 ] && K=6
    EQVIO_OZAKI=$S EQVIO_OZAKI_MIN_TILES=4 python bench.py --features $N --steps $K --warmup 3 --no-sub-configs --no-cpu-baseline > gpurun_out/${tag}_n${N}_oz$S.json 2> gpurun_out/${tag}_n${N}_oz$S.err
    python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_n${N}_oz$S.json"))
r=d["roofline"]
print("N=$N OZAKI=$S value %.1f e2e %.1f ms/period %.3f by_class %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], {k:(round(v["ms_per_period"],3)) for k,v in r["by_class"].items()}))
PY
  done
done
