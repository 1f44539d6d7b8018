"""Summarise an .ncu-rep (from `ncu --set full`) into the handful of numbers DESIGN.md / profiles/ quote.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [launch_index]
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
row = data[idx]
col = {h: i for i, h in enumerate(hdr)}
def g(name):
    i = col.get(name)
    return (row[i], units[i]) if i is not None else ("n/a", "")
want = [
    "Kernel Name", "Block Size", "Grid Size",
    "gpu__time_duration.sum",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
]
for w in want:
    v, u = g(w)
    print(f"{w:82s} {v} {u}")
# warp stall reasons (per-issue sampling)
stalls = [(h, row[i]) for h, i in col.items() if h.startswith("smsp__average_warp_latency_issue_stalled") or h.startswith("smsp__average_warps_issue_stalled")]
vals = []
for h, v in stalls:
    try: vals.append((float(v), h))
    except ValueError: pass
for v, h in sorted(vals, reverse=True)[:8]:
    print(f"{h:82s} {v:.3f}")
