"""Per-kernel table from an ncu CSV (one row per launch x metric, `--csv --log-file`), for profiles/rNN_kernels_ncu.md.
    python tools/kernel_table.py gpurun_out/kernels_n512.csv [HBM_GBS] > profiles/r02_kernels_ncu_n512.md
Columns: launches, average duration, DMMA sub-pipe busy (% of elapsed cycles, average over launches weighted by duration),
DRAM bytes per launch (read + write) and the DRAM rate that implies against the measured copy bandwidth."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
hbm = float(sys.argv[2]) if len(sys.argv) > 2 else 6546.2
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}


def short(name):
    name = name.replace("void ", "")
    if "<" in name and name.rfind(">") > 0:
        name = name[: name.rfind(">") + 1]
    else:
        name = name.split("(")[0]
    return name.replace("eqvio::", "")


per = defaultdict(dict)  # launch id -> metric -> value
kname = {}
for r in rows[1:]:
    lid = r[ix["ID"]]
    kname[lid] = short(r[ix["Kernel Name"]])
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[ix["Metric Unit"]]
    m = r[ix["Metric Name"]]
    if m == "gpu__time_duration.sum":
        v = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v * 1e6
    if m.startswith("dram__bytes"):
        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    per[lid][m] = v
agg = defaultdict(lambda: dict(n=0, us=0.0, dmma=0.0, imma=0.0, rd=0.0, wr=0.0, regs=0, waves=0.0, warps=0.0, grid=0.0))
for lid, m in per.items():
    a = agg[kname[lid]]
    us = m.get("gpu__time_duration.sum", 0.0)
    a["n"] += 1
    a["us"] += us
    a["dmma"] += us * m.get("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", 0.0)
    a["imma"] += us * m.get("sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    a["warps"] += us * m.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
    a["rd"] += m.get("dram__bytes_read.sum", 0.0)
    a["wr"] += m.get("dram__bytes_write.sum", 0.0)
    a["regs"] = max(a["regs"], int(m.get("launch__registers_per_thread", 0)))
    a["waves"] = max(a["waves"], m.get("launch__waves_per_multiprocessor", 0.0))
    a["grid"] = max(a["grid"], m.get("launch__grid_size", 0.0))
tot = sum(a["us"] for a in agg.values())
print("| kernel | launches | avg us | share | DMMA pipe %% | int8 tensor pipe (tcgen05) %% | DRAM rd+wr per launch | DRAM GB/s (of %.1f measured) | regs | max grid | max waves/SM | warps active %% |" % hbm)
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
    us = a["us"] / a["n"]
    by = (a["rd"] + a["wr"]) / a["n"]
    gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] > 0 else 0.0
    print(f"| `{k}` | {a['n']} | {us:.1f} | {100 * a['us'] / tot:.1f} % | {a['dmma'] / a['us'] if a['us'] else 0:.1f} | {a['imma'] / a['us'] if a['us'] else 0:.1f} | {by / 1e6:.3f} MB | {gbs:.0f} ({100 * gbs / hbm:.1f} %) | {a['regs']} | {int(a['grid'])} | {a['waves']:.2f} | {a['warps'] / a['us'] if a['us'] else 0:.1f} |")
