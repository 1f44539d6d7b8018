"""Bracketed launches of the covariance-update end of one vision frame (direct launches, CUDA events: durations are per launch, the
overlap is not what the replayed graph does).   python tools/update_tail.py [N=512]"""
import os, sys, json, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
out = os.path.join(root, "gpurun_out", f"timeline_n{N}.json")
subprocess.check_call([sys.executable, os.path.join(root, "tools", "timeline.py"), "--features", str(N), "--out", out])
d = json.load(open(out))
E = sorted(d["entries"], key=lambda e: e[2])
cls, lanes = d["classes"], d["lanes"]
small = [e for e in E if e[4] and e[4] < 2e8]
print("thin products (flops < 0.2 G):")
for e in small:
    print(f"  {lanes[int(e[1])]:8s} {cls[int(e[0])]:14s} start {e[2]*1e3:9.1f} us  dur {(e[3]-e[2])*1e3:7.1f} us  {e[4]/1e6:8.1f} MFLOP")
print("last 40 bracketed launches:")
for e in E[-40:]:
    print(f"  {lanes[int(e[1])]:8s} {cls[int(e[0])]:14s} start {e[2]*1e3:9.1f} us  dur {(e[3]-e[2])*1e3:7.1f} us  {e[4]/1e6:10.1f} MFLOP")
