#!/bin/bash
# ncu evidence of one build: per-kernel launch lists (N = 512, 256, 64, 1024) and one `--set full` capture of the dominant kernel at N = 512.
# Usage: tools/gpu_evidence.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__waves_per_multiprocessor,launch__grid_size
for N in 512 256 64 1024; do
  timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_kernels_n${N}.csv python tools/profile_workload.py --features $N --periods 1 > gpurun_out/${tag}_ncu_n${N}.log 2>&1; echo "ncu N=$N rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_oz_riccati -c 2 -o gpurun_out/${tag}_oz_riccati_n512 -f python tools/profile_workload.py --features 512 --periods 1 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
