"""A short filter workload for ncu captures: W warm-up vision periods, then P profiled ones (direct launches, so that every
kernel is a separate launch whatever the graph cache holds).
    ncu --metrics ... python tools/profile_workload.py --features 512 --periods 1 [--fast] [--churn]
cudaProfilerStart/Stop bracket the profiled periods: run ncu with `--profile-from-start off`."""
import argparse
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--features", type=int, default=512)
ap.add_argument("--periods", type=int, default=1)
ap.add_argument("--warm", type=int, default=2)
ap.add_argument("--fast", action="store_true")
ap.add_argument("--graphs", action="store_true")
a = ap.parse_args()

from eqf_vio_b200.filter import VIOFilter  # noqa: E402
from eqf_vio_b200.settings import conditioned_settings  # noqa: E402
from eqf_vio_b200.synthetic import period_sequence  # noqa: E402

rt = ctypes.CDLL("libcudart.so")
s = conditioned_settings(fastRiccati=a.fast)
seq = period_sequence(a.features, a.warm + a.periods, camera_offset=tuple(s.cameraOffset))
f = VIOFilter(s)
if not a.graphs:
    f.set_graphs(False)
started = False
for kind, i in seq.events():
    if kind == "imu":
        f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
    else:
        f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
        if i == a.warm and not started:
            f.synchronize()
            rt.cudaProfilerStart()
            started = True
f.synchronize()
rt.cudaProfilerStop()
print("landmarks", f.numLandmarks, "launches", f.launch_count())
