#!/bin/bash
# Round of the one-product covariance update Sigma - K (C Sigma): the whole GPU suite, then A/B against the reference's association.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash tools/gpu_ab.sh kcs EQVIO_SIGMA_KCS "0 1 0 1" 512
bash tools/gpu_ab.sh kcs EQVIO_SIGMA_KCS "0 1" 256
bash tools/gpu_ab.sh kcs EQVIO_SIGMA_KCS "0 1" 1024
bash tools/gpu_ab.sh kcs EQVIO_SIGMA_KCS "0 1" 64
timeout 200 python tools/graph_stamps.py --features 512 > gpurun_out/kcs_update_stamps_n512.txt 2>&1; tail -16 gpurun_out/kcs_update_stamps_n512.txt
