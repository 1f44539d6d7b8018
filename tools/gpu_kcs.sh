#!/bin/bash
# Round after a change in the update: the whole GPU suite, the headline sizes, the in-graph timestamps of the update.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash tools/gpu_ab.sh upd EQVIO_SIGMA_KCS "1 1" 512
bash tools/gpu_ab.sh upd EQVIO_SIGMA_KCS "1" 256
bash tools/gpu_ab.sh upd EQVIO_SIGMA_KCS "1" 1024
timeout 200 python tools/graph_stamps.py --features 512 > gpurun_out/upd_update_stamps_n512.txt 2>&1; tail -16 gpurun_out/upd_update_stamps_n512.txt
