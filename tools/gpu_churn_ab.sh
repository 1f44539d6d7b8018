#!/bin/bash
# Churn record with / without the "no new captures while landmarks change" rule and with a larger graph cache.
for cfg in "2 16" "0 16" "0 32" "0 64" "2 16"; do
  set -- $cfg
  echo "EQVIO_GRAPH_STABLE=$1 EQVIO_GRAPH_CACHE=$2: $(EQVIO_GRAPH_STABLE=$1 EQVIO_GRAPH_CACHE=$2 timeout 200 python tools/churn_bench.py 40 2>&1 | tail -1)"
done
