"""Generates the committed golden vectors (tests/golden/*.npz).

The reference ships no known-answer vectors for the filter and cannot be built here, so
these are produced by the C restatement (oracle/eqvio_oracle.c) after it was cross-checked against the
independent numpy restatement and the reference's property tests.  They freeze the oracle's outputs so a
later change to either the oracle or the CUDA path shows up as a diff against history.

    python tests/golden/make_golden.py        # rewrites the .npz files in place
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eqf_vio_b200.settings import template_settings  # noqa: E402
from eqf_vio_b200.synthetic import period_sequence  # noqa: E402
from oracle.c_oracle import COracleFilter  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sequence_case(name, N, periods, **overrides):
    s = template_settings(**overrides)
    seq = period_sequence(N, periods, camera_offset=tuple(s.cameraOffset))
    f = COracleFilter(s)
    snaps, status = [], []
    for kind, i in seq.events():
        if kind == "imu":
            status.append(f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7]))
        else:
            status.append(f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i]))
            snaps.append(f.get_snapshot())
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        N=N, periods=periods, overrides=np.array(repr(sorted(overrides.items()))),
        imu=seq.imu, vision_stamps=seq.vision_stamps, ids=seq.ids, bearings=seq.bearings,
        status=np.array(status), **{f"snap{j}": sn for j, sn in enumerate(snaps)},
    )


def pieces_case(name, N):
    s = template_settings(outlierThreshold=1e9)
    seq = period_sequence(N, 2, camera_offset=tuple(s.cameraOffset))
    f = COracleFilter(s)
    for kind, i in seq.events():
        if (kind, i) == ("vision", 2):
            break
        if kind == "imu":
            f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else:
            f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
    snap = f.get_snapshot()
    omega, T = np.array([0.11, -0.23, 0.07]), 0.005
    F, Bb = f.build_FB(T, omega)
    y = seq.bearings[2]
    C, delta = f.build_C_delta(y)
    rng = np.random.default_rng(7)
    g_eqf = rng.standard_normal(5 + 3 * N) * 1e-2
    Gamma = f.bundle_lift(g_eqf)
    f.riccati_propagate(T, omega)
    Sigma_prop = f.stateCovariance()
    f.set_snapshot(snap)
    K, gamma = f.gain_update(y)
    Sigma_upd = f.stateCovariance()
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        N=N, snapshot=snap, omega=omega, T=T, F=F, Bb=Bb, bearings=y, C=C, delta=delta, gamma_eqf=g_eqf, Gamma=Gamma,
        Sigma_prop=Sigma_prop, K=K, gamma=gamma, Sigma_upd=Sigma_upd,
    )


if __name__ == "__main__":
    sequence_case("seq_config1_N5", 5, 1)                       # BASELINE config 1: 10 IMU ticks + 1 vision frame, N = 5
    sequence_case("seq_N5_p4", 5, 4, outlierThreshold=1e9)
    sequence_case("seq_N16_p4_fastriccati", 16, 4, outlierThreshold=1e9, fastRiccati=True)
    sequence_case("seq_N8_p3_continuous", 8, 3, outlierThreshold=1e9, useDiscreteVelocityLift=False, useDiscreteInnovationLift=False)
    sequence_case("seq_N8_p3_nolift", 8, 3, outlierThreshold=1e9, useInnovationLift=False)
    pieces_case("pieces_N8", 8)
    print("golden vectors written to", HERE)
