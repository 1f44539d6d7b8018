"""ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) -> markdown table per kernel.
    python tools/launch_list.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].replace("void ", "")
    name = name.split("(")[0] if "<" not in name else name[: name.rfind(">") + 1] if name.rfind(">") > 0 else name
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v
    agg[name][0] += 1
    agg[name][1] += us
mine = {k: v for k, v in agg.items() if k.startswith("eqvio::")}
tot = sum(v[1] for v in mine.values())
print("| kernel | launches | total ms | share of this repo's kernels | avg us |")
print("|---|---|---|---|---|")
for k, (n, us) in sorted(mine.items(), key=lambda x: -x[1][1]):
    print(f"| `{k}` | {n} | {us / 1e3:.3f} | {100 * us / tot:.1f} % | {us / n:.1f} |")
other = {k: v for k, v in agg.items() if not k.startswith("eqvio::")}
if other:
    print("\nOther launches in the window: " + ", ".join(f"`{k[:60]}` x{v[0]}" for k, v in other.items()))
