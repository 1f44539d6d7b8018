#!/usr/bin/env python
"""%globaltimer marks inside the vision update while it runs as a replayed CUDA graph (EQVIO_STAMPS=1 adds one
single-thread stamp kernel per mark to the captured sequence; CUDA events cannot time points inside a graph).

    EQVIO_STAMPS=1 python tools/graph_stamps.py --features 512
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys

os.environ.setdefault("EQVIO_STAMPS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NAMES = ["begin", "lift: setup done", "lift: elimination done", "lift: R^T done", "main: C, delta", "main: S formed",
         "main: S^-1 done", "main: K", "main: gamma", "side: Sigma C^T done", "main: lift joined", "main: lift features(gamma)",
         "main: X updated", "side: Sigma update done", "end", "side: K rows split", "helper: Sigma update strips done", "main: S^-1 columns split",
         "main: C Sigma done", "main: C Sigma rows split"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--features", type=int, default=512)
    ap.add_argument("--periods", type=int, default=8)
    args = ap.parse_args()
    from eqf_vio_b200 import abi
    from eqf_vio_b200.filter import VIOFilter
    from eqf_vio_b200.settings import conditioned_settings
    from eqf_vio_b200.synthetic import period_sequence

    s = conditioned_settings()
    seq = period_sequence(args.features, args.periods, camera_offset=tuple(s.cameraOffset))
    f = VIOFilter(s, device=0)
    L = abi.lib()
    L.eqvio_debug_stamps.restype = C.c_int
    L.eqvio_debug_stamps.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]
    for kind, i in seq.events():
        if kind == "imu":
            f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else:
            f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
    f.synchronize()
    out = (C.c_ulonglong * 512)()
    abi.check(L.eqvio_debug_stamps(f._h, out, 512), "eqvio_debug_stamps")
    t0 = out[0]
    for name, base in (("lift", 32), ("S", 272)):
        rows = []
        for j in range(80):
            k, c, t = out[base + 3 * j], out[base + 3 * j + 1], out[base + 3 * j + 2]
            if base + 3 * j + 2 >= (272 if base == 32 else 512) or k == 0:
                break
            rows.append(((k - t0) / 1e3, (c - t0) / 1e3, (t - t0) / 1e3))
        print(f"{name} chain, per block: chain kernel done / column panel done / trailing update done (us)")
        for j, r in enumerate(rows):
            prev = rows[j - 1][0] if j else float("nan")
            print(f"   {j:3d}  {r[0]:9.1f} {r[1]:9.1f} {r[2]:9.1f}   period {r[0] - prev:6.1f}")
    print(f"N={args.features}: last vision update, graph replays so far {f.graph_stats()[0]}")
    for name, t in sorted(((nm, tt) for nm, tt in zip(NAMES, list(out)[: len(NAMES)]) if tt >= t0), key=lambda x: x[1]):
        print(f"  {(t - t0) / 1e3:9.1f} us  {name}")


if __name__ == "__main__":
    main()
