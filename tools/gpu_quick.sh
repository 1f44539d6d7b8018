#!/bin/bash
# Quick round: the tests that exercise the lift and the int8 paths, the headline bench twice (+ N = 1024), the in-graph timestamps of the update.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_schur.py tests/test_gpu_ozaki.py -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_ab.sh q EQVIO_SIGMA_KCS "1 1" 512
bash tools/gpu_ab.sh q EQVIO_SIGMA_KCS "1" 1024
timeout 200 python tools/graph_stamps.py --features 512 > gpurun_out/q_update_stamps_n512.txt 2>&1; tail -20 gpurun_out/q_update_stamps_n512.txt
