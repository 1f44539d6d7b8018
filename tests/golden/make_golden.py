"""Generates the committed golden vectors (tests/golden/*.npz).

Source of truth: the REFERENCE's own, unmodified filter sources (/root/reference/eqf_vio/src/*.cpp,
libs/core/src/*.cpp) compiled here against the Eigen-API stand-in of oracle/refshim (Eigen3 itself is
not installed; see oracle/README.md) and driven through oracle/ref_binding.py.  The reference ships no
known-answer vectors for the filter, so these freeze what its code computes on seeded
synthetic inputs; the restated oracles (C, numpy) and the CUDA path are then tested against them on the
GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py        # rewrites the .npz files in place (needs /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eqf_vio_b200.settings import conditioned_settings, template_settings  # noqa: E402
from eqf_vio_b200.synthetic import period_sequence  # noqa: E402
from oracle import ref_binding  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = "pvangoor/eqf_vio @ 0b1334ec sources compiled against oracle/refshim (Eigen stand-in)"


def make_settings(base, overrides):
    return (conditioned_settings if base == "conditioned" else template_settings)(**overrides)


def sequence_case(name, base, N, periods, select=None, **overrides):
    """Whole sequence through the reference; a snapshot after every vision frame."""
    s = make_settings(base, overrides)
    seq = period_sequence(N, periods, camera_offset=tuple(s.cameraOffset))
    f = ref_binding.ReferenceFilter(s)
    rng = np.random.default_rng(11)
    snaps, sels = [], []
    for kind, i in seq.events():
        if kind == "imu":
            f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else:
            sel = np.arange(N) if select is None else select(rng, i, N)
            f.processVisionData(seq.vision_stamps[i], seq.ids[sel], seq.bearings[i][sel])
            snaps.append(f.get_snapshot())
            sels.append(np.asarray(sel, dtype=np.int32))
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        source=np.array(SOURCE), base=np.array(base), N=N, periods=periods, overrides=np.array(repr(sorted(overrides.items()))),
        imu=seq.imu, vision_stamps=seq.vision_stamps, ids=seq.ids, bearings=seq.bearings,
        **{f"snap{j}": sn for j, sn in enumerate(snaps)}, **{f"sel{j}": se for j, se in enumerate(sels)},
    )


def steps_case(name, base, N, **overrides):
    """One IMU step and one vision step from a given state, plus the reference's free functions
    (A0, Bt, C0, delta, bundleLift) evaluated at that state."""
    s = make_settings(base, overrides)
    seq = period_sequence(N, 3, camera_offset=tuple(s.cameraOffset))
    f = ref_binding.ReferenceFilter(s)
    ev = list(seq.events())
    stop = ev.index(("vision", 2)) - 1  # the last IMU tick before vision 2
    for kind, i in ev[:stop]:
        if kind == "imu":
            f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else:
            f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
    snap0 = f.get_snapshot()
    omega = np.array([0.11, -0.23, 0.07])
    A0, Bt, C0 = f.state_matrix_A(omega), f.input_matrix_B(), f.output_matrix_C()
    y = seq.bearings[2]
    delta = f.delta(y)
    g_eqf = np.random.default_rng(7).standard_normal(5 + 3 * N) * 1e-2
    Gamma = f.bundle_lift(g_eqf)
    kind, i = ev[stop]
    assert kind == "imu"
    imu_row = seq.imu[i].copy()
    f.processIMUData(imu_row[0], imu_row[1:4], imu_row[4:7])
    snap_imu = f.get_snapshot()
    f.processVisionData(seq.vision_stamps[2], seq.ids, y)
    snap_vis = f.get_snapshot()
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        source=np.array(SOURCE), base=np.array(base), N=N, overrides=np.array(repr(sorted(overrides.items()))),
        snapshot=snap0, omega=omega, A0=A0, Bt=Bt, C0=C0, bearings=y, delta=delta, gamma_eqf=g_eqf, Gamma=Gamma,
        imu_row=imu_row, snap_after_imu=snap_imu, vision_stamp=seq.vision_stamps[2], ids=seq.ids, snap_after_vision=snap_vis,
    )


def ragged(rng, i, N):
    return np.arange(2 * N // 3) if i == 0 else np.sort(rng.choice(N, size=int(rng.integers(N // 4, N)), replace=False))


if __name__ == "__main__":
    if not ref_binding.build():
        raise SystemExit("needs /root/reference (the reference sources) to build oracle/_ref")
    for old in os.listdir(HERE):
        if old.endswith(".npz"):
            os.remove(os.path.join(HERE, old))
    sequence_case("seq_config1_N5", "template", 5, 1)                      # BASELINE config 1: 10 IMU ticks + 1 vision frame, N = 5
    sequence_case("seq_template_N5_p4", "template", 5, 4, outlierThreshold=1e9)
    sequence_case("seq_conditioned_N16_p8", "conditioned", 16, 8)
    sequence_case("seq_conditioned_N16_p6_fastriccati", "conditioned", 16, 6, fastRiccati=True)
    sequence_case("seq_conditioned_N8_p5_continuous", "conditioned", 8, 5, useDiscreteVelocityLift=False, useDiscreteInnovationLift=False)
    sequence_case("seq_conditioned_N8_p5_nolift", "conditioned", 8, 5, useInnovationLift=False)
    sequence_case("seq_bookkeeping_N12_p8", "template", 12, 8, select=ragged)  # landmarks come and go, template outlier threshold
    steps_case("steps_template_N8", "template", 8, outlierThreshold=1e9)
    steps_case("steps_conditioned_N24", "conditioned", 24)
    print("golden vectors written to", HERE)
