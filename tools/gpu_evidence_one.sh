#!/bin/bash
# ncu launch list (per-kernel metrics) of one profiled vision period at one size.  Usage: tools/gpu_evidence_one.sh <tag> <N>
tag=${1:-r02}; N=${2:-512}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__waves_per_multiprocessor,launch__grid_size
timeout 330 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_kernels_n${N}.csv python tools/profile_workload.py --features $N --periods 1 > gpurun_out/${tag}_ncu_n${N}.log 2>&1; echo "ncu N=$N rc=$?"
python tools/kernel_table.py gpurun_out/${tag}_kernels_n${N}.csv > gpurun_out/${tag}_kernels_n${N}.md 2>&1; head -30 gpurun_out/${tag}_kernels_n${N}.md
