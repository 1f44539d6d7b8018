#!/usr/bin/env python
"""Per-launch timeline of one vision period (10 IMU ticks + 1 vision frame) on the B200 path.

    python tools/timeline.py --features 512 [--out gpurun_out/timeline_n512.json]

Every launch the library brackets with CUDA events (eqvio_profile_enable) is listed with its stream lane, so
the critical path of the vision update — the two blocked Schur eliminations that stand in for `S.inverse()`
(reference eqf_vio/src/VIOFilter.cpp:277) and `Sigma.inverse()` (eqf_vio/src/EqFMatrices.cpp:239) — can be read
off.  Prints a per-lane / per-class summary of the vision frame and writes the raw entries as JSON.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--features", type=int, default=512)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from eqf_vio_b200.filter import VIOFilter
    from eqf_vio_b200.settings import conditioned_settings
    from eqf_vio_b200.synthetic import period_sequence

    N = args.features
    s = conditioned_settings()
    seq = period_sequence(N, args.warm + 2, camera_offset=tuple(s.cameraOffset))
    f = VIOFilter(s, device=0)
    n_vis = 0
    for kind, i in seq.events():
        if kind == "vision" and n_vis == args.warm + 1:
            f.synchronize()
            f.profile_enable(True)
        if kind == "imu":
            f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else:
            f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
            n_vis += 1
            if n_vis == args.warm + 2:
                break
    f.synchronize()
    tl = f.profile_timeline()
    f.profile_enable(False)
    cls_names, lanes = VIOFilter.PROFILE_CLASSES, VIOFilter.PROFILE_LANES
    t_end = tl[:, 3].max()
    print(f"N={N}: vision frame (Riccati + update) spans {t_end:.3f} ms, {len(tl)} bracketed launches")
    print("per class:  busy ms (sum of launch durations) / launches")
    for c, name in enumerate(cls_names):
        sel = tl[:, 0] == c
        if sel.any():
            print(f"  {name:14s} {np.sum(tl[sel, 3] - tl[sel, 2]):8.3f} ms  {int(sel.sum()):5d}")
    print("per lane:   first start .. last end, busy ms")
    for l, name in enumerate(lanes):
        sel = tl[:, 1] == l
        if sel.any():
            print(f"  {name:12s} {tl[sel, 2].min():7.3f} .. {tl[sel, 3].max():7.3f}   busy {np.sum(tl[sel, 3] - tl[sel, 2]):7.3f} ms")
    # the chain kernels in order, with the gap to the previous chain launch on the same lane
    for lane in (0, 2):
        sel = (tl[:, 1] == lane) & (tl[:, 0] == 3)
        d = tl[sel]
        if len(d) > 1:
            dur = d[:, 3] - d[:, 2]
            per = np.diff(d[:, 2])
            print(f"  {lanes[lane]} chain: {len(d)} diagonal blocks, LU kernel {dur.mean() * 1e3:.1f} us avg, block-to-block period {per.mean() * 1e3:.1f} us avg (min {per.min() * 1e3:.1f}, max {per.max() * 1e3:.1f})")
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump({"features": N, "columns": ["class", "lane", "t0_ms", "t1_ms", "flops"], "classes": cls_names, "lanes": lanes,
                   "entries": tl.tolist()}, open(args.out, "w"))


if __name__ == "__main__":
    main()
