"""Replay driver with the reference's wire formats (reference: eqf_vio/src/main.cpp:42-203).

    python -m eqf_vio_b200.replay IMU.csv MEAS.csv [CONFIG.yaml] [--out-state S.csv] [--out-filter F.csv]

* IMU csv: header line, then `t, wx, wy, wz, ax, ay, az` (main.cpp:184-190)
* measurement csv: header line, then `t, N, (id, x, y, z) x N` (main.cpp:192-203)
* config: the reference's YAML with an `eqf:` section (VIOFilterSettings.h:56-109) and a `main:` section
  (`startTime`, `writeState`, `writeFilter`; main.cpp:67-77)
* state csv: `time, tx, ty, tz, qw, qx, qy, qz, vx, vy, vz, N, (id, x, y, z) x N` (main.cpp:96-98, VIOState.cpp:72-84)
* filter csv: `time, xi0 pose/velocity, X.A, X.w, N, (id, q0, Q quaternion, Q scale) x N, Sigma row-major`
  (main.cpp:99-107, VIOFilter.cpp:311-341)
Merge order: an IMU row is consumed while `imu.stamp < meas.stamp`, else the measurement row; rows with
`stamp <= startTime` are skipped; the state is written after every measurement row (main.cpp:111-140).
The loop ends when either file runs out, like the reference.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

from .settings import default_settings, settings_from_eqf_node


def read_imu_csv(path):
    rows = np.loadtxt(path, delimiter=",", skiprows=1, ndmin=2)
    return rows[:, :7]


def read_meas_csv(path):
    out = []
    with open(path) as f:
        next(f)
        for line in f:
            c = [x for x in line.strip().split(",") if x.strip() != ""]
            if not c:
                continue
            t, n = float(c[0]), int(float(c[1]))
            vals = np.array(c[2 : 2 + 4 * n], dtype=float).reshape(n, 4)
            out.append((t, vals[:, 0].astype(np.int32), vals[:, 1:4].copy()))
    return out


def write_imu_csv(path, imu):
    with open(path, "w") as f:
        f.write("t, wx, wy, wz, ax, ay, az\n")
        for r in imu:
            f.write(", ".join(repr(float(v)) for v in r[:7]) + "\n")


def write_meas_csv(path, stamps, ids, bearings):
    with open(path, "w") as f:
        f.write("t, N, id1, x1, y1, z1, ...\n")
        for t, y in zip(stamps, bearings):
            parts = [repr(float(t)), str(len(ids))]
            for i, p in zip(ids, y):
                parts += [str(int(i)), repr(float(p[0])), repr(float(p[1])), repr(float(p[2]))]
            f.write(", ".join(parts) + "\n")


def _stream_double(v, precision):
    """What `os << std::setprecision(p) << v` prints for a double with the default float field: C's %.{p}g
    (exponent notation for small / large magnitudes, trailing zeros dropped)."""
    return "%.*g" % (precision, float(v))


def format_state_row(t, est, precision=5):
    """`setprecision(20) time, setprecision(5) state` (main.cpp:135-137; operator<< VIOState.cpp:72-84)."""
    g = lambda v: _stream_double(v, precision)
    parts = [_stream_double(t, 20)]
    p = est.pose
    parts += [g(p[0]), g(p[1]), g(p[2]), g(p[3]), g(p[4]), g(p[5]), g(p[6])]
    parts += [g(v) for v in est.velocity]
    parts.append(str(len(est.ids)))
    for i, q in zip(est.ids, est.bodyLandmarks):
        parts += [str(int(i)), g(q[0]), g(q[1]), g(q[2])]
    return ", ".join(parts)


def format_filter_row(t, snap, precision=5):
    """operator<<(ostream&, VIOFilter) (VIOFilter.cpp:311-341) from a snapshot (include/eqvio.h layout)."""
    g = lambda v: _stream_double(v, precision)
    N = int(snap[0])
    n = 11 + 3 * N
    parts = [_stream_double(t, 20)]
    parts += [g(v) for v in (*snap[26:29], *snap[22:26], *snap[29:32], *snap[43:46], *snap[39:43], *snap[46:49])]
    parts.append(str(N))
    L = snap[49 : 49 + 9 * N].reshape(N, 9)
    for r in L:
        parts += [str(int(r[0]))] + [g(v) for v in r[1:9]]
    S = snap[49 + 9 * N :].reshape(n, n, order="F")
    parts += [g(v) for v in S.reshape(-1)]  # row-major, IOFormat(-1, 0, ", ", ", ")
    return ", ".join(parts)


def replay(filt, imu, meas, start_time=0.0, on_state=None, on_filter=None):
    """The reference's merge loop (main.cpp:111-170) on any object with the VIOFilter method names."""
    i = j = n_imu = n_vis = 0
    if len(imu) == 0 or len(meas) == 0:
        return 0, 0
    while True:
        if imu[i, 0] < meas[j][0]:
            if imu[i, 0] > start_time:
                filt.processIMUData(imu[i, 0], imu[i, 1:4], imu[i, 4:7])
                n_imu += 1
            i += 1
            if i == len(imu):
                break
        else:
            t, ids, y = meas[j]
            if t > start_time:
                filt.processVisionData(t, ids, y)
                n_vis += 1
            if on_state is not None:
                on_state(filt.getTime(), filt.stateEstimate())
            if on_filter is not None:
                on_filter(filt.getTime(), filt.get_snapshot())
            j += 1
            if j == len(meas):
                break
    return n_imu, n_vis


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("imu_file")
    ap.add_argument("meas_file")
    ap.add_argument("config_file", nargs="?")
    ap.add_argument("--out-state")
    ap.add_argument("--out-filter")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    settings, start_time, write_state, write_filter = default_settings(), 0.0, bool(a.out_state), bool(a.out_filter)
    if a.config_file:
        import yaml

        with open(a.config_file) as f:
            cfg = yaml.safe_load(f)
        settings = settings_from_eqf_node(cfg.get("eqf", {}))
        m = cfg.get("main", {})
        start_time = float(m.get("startTime", 0.0))
        write_state = write_state or bool(m.get("writeState", False))
        write_filter = write_filter or bool(m.get("writeFilter", False))
    from .filter import VIOFilter

    filt = VIOFilter(settings, device=a.device)
    fs = open(a.out_state or "EQF_VIO_output.csv", "w") if write_state else None
    ff = open(a.out_filter or "EQF_VIO_internal.csv", "w") if write_filter else None
    if fs:
        fs.write("time, tx, ty, tz, qw, qx, qy, qz, vx, vy, vz, N, p1id, p1x, p1y, p1z, ..., ..., ..., ..., pNid, pNx, pNy, pNz\n")
    if ff:
        ff.write("time, t0x, t0y, t0z, q0w, q0x, q0y, q0z, v0x, v0y, v0z, tAx, tAy, tAz, qAw, qAx, qAy, qAz, wx, wy, wz, N, "
                 "p1id, p1x, p1y, p1z, qQ1w, qQ1x, qQ1y, qQ1z, aQ1, ..., pNid, pNx, pNy, pNz, qQNw, qQNx, qQNy, qQNz, aQN, "
                 "Sigma(1,1), Sigma(1,2), ..., Sigma(5+3N, 5+3N)\n")
    n_imu, n_vis = replay(
        filt, read_imu_csv(a.imu_file), read_meas_csv(a.meas_file), start_time,
        on_state=(lambda t, e: fs.write(format_state_row(t, e) + "\n")) if fs else None,
        on_filter=(lambda t, s: ff.write(format_filter_row(t, s) + "\n")) if ff else None,
    )
    for f in (fs, ff):
        if f:
            f.close()
    print(f"Processed {n_imu} IMU and {n_vis} vision measurements.")
    return 0


if __name__ == "__main__":
    sys.exit(main())
