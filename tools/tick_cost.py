"""Host enqueue time vs device time of the IMU ticks and of the vision update, per N (run under gpurun).
Tells whether a bench period is bound by the host (launch path) or by the GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from eqf_vio_b200.filter import VIOFilter
from eqf_vio_b200.settings import conditioned_settings
from eqf_vio_b200.synthetic import period_sequence

for N in [int(a) for a in sys.argv[1:]] or [64, 256, 512]:
    s = conditioned_settings()
    P = 30
    seq = period_sequence(N, P + 1, camera_offset=tuple(s.cameraOffset))
    f = VIOFilter(s, device=0)
    ext = torch.cuda.ExternalStream(f.stream_ptr())
    ydev = torch.tensor(seq.bearings, dtype=torch.float64, device="cuda").contiguous()
    periods, cur = [], []
    for kind, i in seq.events():
        cur.append((kind, i))
        if kind == "vision":
            periods.append(cur); cur = []
    host_imu, dev_imu, host_vis, dev_vis = [], [], [], []
    for pi, evs in enumerate(periods):
        f.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(ext)
        t0 = time.perf_counter()
        for kind, i in evs[:-1]:
            f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        t1 = time.perf_counter()
        e[1].record(ext)
        f.synchronize()
        kind, i = evs[-1]
        e[1].record(ext)
        t2 = time.perf_counter()
        f.processVisionDataDevice(seq.vision_stamps[i], seq.ids, ydev[i].data_ptr())
        t3 = time.perf_counter()
        e[2].record(ext)
        f.synchronize()
        if pi >= 6:
            nt = len(evs) - 1
            host_imu.append((t1 - t0) / nt * 1e6); host_vis.append((t3 - t2) * 1e6)
            dev_vis.append(e[1].elapsed_time(e[2]) * 1e3)
    # device time of ticks: enqueue 10 ticks after a sync, events around them (host runs ahead if it can)
    for pi, evs in enumerate(periods[:0]):
        pass
    # second pass measuring device time of the 10 ticks with the host far ahead is the same as above's e0..e1 only if
    # the GPU is the slower side; report wall for ticks+sync as the realised per-tick time
    f2 = f
    print(f"N={N}: host enqueue per IMU tick {np.median(host_imu):.1f} us; vision call host {np.median(host_vis):.0f} us, device {np.median(dev_vis):.0f} us", flush=True)
    # realised time of 10 ticks (host + device pipeline), after the update has drained
    seq2 = period_sequence(N, 8, camera_offset=tuple(s.cameraOffset))
    g = VIOFilter(s, device=0)
    ext2 = torch.cuda.ExternalStream(g.stream_ptr())
    real = []
    cnt = 0
    for kind, i in seq2.events():
        if kind == "imu":
            if cnt == 0:
                g.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record(ext2); tw = time.perf_counter()
            g.processIMUData(seq2.imu[i, 0], seq2.imu[i, 1:4], seq2.imu[i, 4:7]); cnt += 1
        else:
            if cnt:
                b.record(ext2); g.synchronize(); real.append((a.elapsed_time(b) * 1e3 / cnt, (time.perf_counter() - tw) * 1e6 / cnt))
            cnt = 0
            g.processVisionData(seq2.vision_stamps[i], seq2.ids, seq2.bearings[i])
    print(f"N={N}: realised per IMU tick: device events {np.median([r[0] for r in real[3:]]):.1f} us, wall incl. final sync {np.median([r[1] for r in real[3:]]):.1f} us", flush=True)
