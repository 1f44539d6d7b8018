"""Where a vision period goes under landmark churn (template outlier threshold, 5 % of the ids replaced per frame): the library's event
brackets per class and stream for one period, plus wall-clock time of the processVisionData call (which contains the outlier D2H
round trip).      python tools/churn_timeline.py [N=512]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ses = bench.Session(torch, None, N, 6, 3, 0, 0, 1, bench.bench_settings(outlierThreshold=0.01), churn=0.05)
f = ses.f
for _ in range(4):
    ses.run_period_resident(next(ses.it))
f.synchronize()
# wall clock of the calls of one period (graphs on)
evs = next(ses.it)
t = []
for kind, i in evs:
    t0 = time.perf_counter()
    if kind == "imu":
        ses.imu(i)
    else:
        f.processVisionDataDevice(ses.seq.vision_stamps[i], ses.frame_ids[i], ses.ydev[i].data_ptr())
    t.append((kind, 1e6 * (time.perf_counter() - t0)))
f.synchronize()
print("host time per call (us):", " ".join(f"{k[0]}{v:.0f}" for k, v in t))
f.profile_enable(True)
evs = next(ses.it)
t0 = time.perf_counter()
for kind, i in evs:
    if kind == "imu":
        ses.imu(i)
    else:
        f.processVisionDataDevice(ses.seq.vision_stamps[i], ses.frame_ids[i], ses.ydev[i].data_ptr())
f.synchronize()
print(f"one period with event brackets (direct launches): {1e3 * (time.perf_counter() - t0):.2f} ms wall")
tl = f.profile_timeline()
cls, lanes = f.PROFILE_CLASSES, ("main", "side", "lift", "main_h", "lift_h", "state", "other")
rows = tl[np.argsort(tl[:, 2])]
start = rows[0, 2]
# the vision frame's part: from the last Riccati launch on
for row in rows:
    if row[3] - row[2] > 0.02 or cls[int(row[0])] != "small_kernels":
        pass
big = [r for r in rows if (r[3] - r[2]) > 0.015]
for row in big[-40:]:
    print(f"{cls[int(row[0])]:16s} {lanes[int(row[1])]:7s} {1e3*(row[2]-start):9.1f} -> {1e3*(row[3]-start):9.1f} us  ({1e3*(row[3]-row[2]):7.1f})")
