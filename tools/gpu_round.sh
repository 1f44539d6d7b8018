#!/bin/bash
# One GPU-box visit: parity tests, bench line, per-kernel ncu tables.  Usage: tools/gpu_round.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__waves_per_multiprocessor,launch__grid_size
for N in 512 256 64; do
  timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_kernels_n${N}.csv python tools/profile_workload.py --features $N --periods 1 > gpurun_out/${tag}_ncu_n${N}.log 2>&1; echo "ncu N=$N rc=$?"
done
