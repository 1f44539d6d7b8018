"""Timeline of a few IMU ticks (Riccati step) from the library's event brackets: class, lane, start, end per launch.
    EQVIO_OZAKI=8 python tools/tick_timeline.py --features 512"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
ap = argparse.ArgumentParser(); ap.add_argument("--features", type=int, default=512); ap.add_argument("--ticks", type=int, default=3)
a = ap.parse_args()
from eqf_vio_b200.filter import VIOFilter
from eqf_vio_b200.settings import conditioned_settings
from eqf_vio_b200.synthetic import period_sequence
s = conditioned_settings()
seq = period_sequence(a.features, 4, camera_offset=tuple(s.cameraOffset))
f = VIOFilter(s)
ev = list(seq.events())
vis = [k for k, (kind, i) in enumerate(ev) if kind == "vision"]
start = vis[2] + 1          # first IMU tick after the third vision frame
for kind, i in ev[:start]:
    f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7]) if kind == "imu" else f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
f.synchronize()
f.profile_enable(True)
for kind, i in ev[start:start + a.ticks]:
    assert kind == "imu"
    f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
f.synchronize()
tl = f.profile_timeline()
cls, lanes = VIOFilter.PROFILE_CLASSES, ("main", "side", "lift", "main_h", "lift_h", "state", "other")
for row in tl[np.argsort(tl[:, 2])]:
    print(f"{cls[int(row[0])]:14s} {lanes[int(row[1])]:7s} {1e3*row[2]:9.1f} -> {1e3*row[3]:9.1f} us  ({1e3*(row[3]-row[2]):7.1f})  {row[4]/1e9:8.3f} GF")
