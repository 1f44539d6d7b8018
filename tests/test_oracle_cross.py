"""The two CPU restatements (C: oracle/eqvio_oracle.c, numpy: oracle/eqvio_numpy.py) against each other
and against the committed golden vectors.  Tolerances: the per-function pieces agree to round-off; whole
sequences agree to ~1e-9 because the reference's own formulation loses digits (C annihilates the radial
direction that carries variance 5000, so K = Sigma C^T S^-1 is only reproducible to ~1e-11 per step
between any two fp64 implementations — see DESIGN.md "Numerical conditioning")."""
import glob
import os

import numpy as np
import pytest

from eqf_vio_b200.settings import template_settings
from eqf_vio_b200.synthetic import period_sequence
from helpers import feed, golden_paths, golden_settings, golden_tolerances, np_settings, rel, replay_golden, run, split_snapshot
from oracle import c_oracle, eqvio_numpy as onp
from oracle.c_oracle import COracleFilter



def test_dense_helpers():
    rng = np.random.default_rng(0)
    A, B = rng.standard_normal((37, 29)), rng.standard_normal((29, 41))
    assert rel(c_oracle.dgemm(A, B), A @ B) < 1e-14
    assert rel(c_oracle.dgemm(A.T.copy(), B, transA=True), A @ B) < 1e-14
    assert rel(c_oracle.dgemm(A, B.T.copy(), transB=True), A @ B) < 1e-14
    S = rng.standard_normal((50, 50))
    S = S @ S.T + 50 * np.eye(50)
    assert rel(c_oracle.inverse(S), np.linalg.inv(S)) < 1e-12


@pytest.mark.parametrize("N", [5, 16])
def test_pieces_c_vs_numpy(N):
    s = template_settings(outlierThreshold=1e9)
    seq = period_sequence(N, 2, camera_offset=tuple(s.cameraOffset))
    fc, fn = COracleFilter(s), onp.VIOFilter(np_settings(s))
    run(fc, seq, ("vision", 2))
    fn.set_snapshot(fc.get_snapshot())
    assert np.abs(fn.get_snapshot() - fc.get_snapshot()).max() == 0.0
    omega = np.array([0.2, -0.1, 0.3])
    xi0m = onp.project_to_manifold(fn.xi0)
    assert np.abs(fc.state_matrix_A(omega) - onp.state_matrix_A(fn.X, xi0m, omega)).max() < 1e-12
    assert np.abs(fc.input_matrix_B() - onp.input_matrix_B(fn.X, xi0m)).max() < 1e-12
    assert np.abs(fc.output_matrix_C() - onp.output_matrix_C(xi0m)).max() < 1e-13
    y = seq.bearings[2]
    C1, d1 = fc.build_C_delta(y)
    C2, d2 = fn.build_C_delta(y)
    assert np.abs(C1 - C2).max() < 1e-13 and np.abs(d1 - d2).max() < 1e-13
    g = np.random.default_rng(1).standard_normal(5 + 3 * N) * 1e-2
    G1 = fc.bundle_lift(g)
    G2 = onp.bundle_lift(g, fn.xi0, fn.X, fn.Sigma[6:, 6:])
    assert np.abs(G1 - G2).max() < 1e-9 * max(1.0, np.abs(G2).max())
    a1 = fc.lift_innovation(g)
    a2 = onp.lift_innovation(g, xi0m)
    assert np.abs(a1 - np.concatenate([a2.U, a2.u, a2.W.reshape(-1)])).max() < 1e-13
    a1 = fc.lift_innovation(g, wls=True)
    a2 = onp.lift_innovation_wls(g, fn.xi0, fn.X, fn.Sigma[6:, 6:])
    assert np.abs(a1 - np.concatenate([a2.U, a2.u, a2.W.reshape(-1)])).max() < 1e-9
    # one Riccati step from identical state
    F1, B1 = fc.build_FB(0.005, omega)
    fc.riccati_propagate(0.005, omega)
    F2, B2 = fn.riccati(0.005, omega)
    assert np.abs(F1 - F2).max() < 1e-13 and np.abs(B1 - B2).max() < 1e-13
    assert rel(fc.stateCovariance(), fn.Sigma) < 1e-14


@pytest.mark.parametrize("N,periods", [(5, 6), (16, 6)])
def test_sequence_c_vs_numpy(N, periods):
    s = template_settings(outlierThreshold=1e9)
    seq = period_sequence(N, periods, camera_offset=tuple(s.cameraOffset))
    fc, fn = COracleFilter(s), onp.VIOFilter(np_settings(s))
    for kind, i in seq.events():
        r1, r2 = feed(fc, seq, kind, i), feed(fn, seq, kind, i)
        assert r1 == r2
        if kind == "vision":
            h1, S1 = split_snapshot(fc.get_snapshot())
            h2, S2 = split_snapshot(fn.get_snapshot())
            assert rel(S1, S2) < 5e-9
            assert np.abs(h1 - h2).max() < 1e-6


def test_bookkeeping_c_vs_numpy():
    """Landmarks that come and go, default outlier threshold (VIOFilter.cpp:345-443)."""
    rng = np.random.default_rng(3)
    s = template_settings()
    seq = period_sequence(12, 6, camera_offset=tuple(s.cameraOffset))
    fc, fn = COracleFilter(s), onp.VIOFilter(np_settings(s))
    for kind, i in seq.events():
        if kind == "imu":
            feed(fc, seq, kind, i), feed(fn, seq, kind, i)
        else:
            sel = np.sort(rng.choice(12, size=9, replace=False)) if i > 0 else np.arange(8)
            r1, r2 = feed(fc, seq, kind, i, sel=sel), feed(fn, seq, kind, i, sel=sel)
            assert r1 == r2
            a, b = fc.get_snapshot(), fn.get_snapshot()
            assert a.size == b.size and int(a[0]) == int(b[0])
            h1, S1 = split_snapshot(a)
            h2, S2 = split_snapshot(b)
            assert [int(v) for v in h1[49::9]] == fn.X.ids
            assert rel(S1, S2) < 5e-9


@pytest.mark.parametrize("path", golden_paths("seq_"), ids=os.path.basename)
@pytest.mark.parametrize("which", ["c", "numpy"])
def test_golden_sequences(path, which):
    """Both restatements against the vectors recorded from the reference's own sources."""
    z = np.load(path)
    s = golden_settings(z)
    f = COracleFilter(s) if which == "c" else onp.VIOFilter(np_settings(s))
    tol_s, tol_h = golden_tolerances(z)

    def check(j, gold):
        snap = f.get_snapshot()
        assert snap.size == gold.size, (j, snap[0], gold[0])  # same landmark count
        hg, Sg = split_snapshot(gold)
        h, S = split_snapshot(snap)
        assert np.array_equal(h[49::9], hg[49::9])  # same ids, same order
        assert rel(S, Sg) < tol_s and np.abs(h - hg).max() < tol_h, (j, rel(S, Sg), np.abs(h - hg).max())

    replay_golden(z, f, check)


@pytest.mark.parametrize("path", golden_paths("steps_"), ids=os.path.basename)
def test_golden_steps(path):
    """The reference's free functions (A0, Bt, C0, delta, bundleLift) and one IMU / one vision step from
    a recorded state: per-step, so the tolerance is tight for both start-ups."""
    z = np.load(path)
    s = golden_settings(z)
    N = int(z["N"])
    fc, fn = COracleFilter(s), onp.VIOFilter(np_settings(s))
    fc.set_snapshot(z["snapshot"])
    fn.set_snapshot(z["snapshot"])
    xi0m = onp.project_to_manifold(fn.xi0)
    for A0, Bt, C0 in ((fc.state_matrix_A(z["omega"]), fc.input_matrix_B(), fc.output_matrix_C()),
                       (onp.state_matrix_A(fn.X, xi0m, z["omega"]), onp.input_matrix_B(fn.X, xi0m), onp.output_matrix_C(xi0m))):
        assert np.abs(A0 - z["A0"]).max() < 1e-12 and np.abs(Bt - z["Bt"]).max() < 1e-12 and np.abs(C0 - z["C0"]).max() < 1e-13
    assert np.abs(fc.build_C_delta(z["bearings"])[1] - z["delta"]).max() < 1e-14
    assert np.abs(fn.build_C_delta(z["bearings"].reshape(N, 3))[1] - z["delta"]).max() < 1e-14
    assert np.abs(fc.bundle_lift(z["gamma_eqf"]) - z["Gamma"]).max() < 1e-10
    assert np.abs(onp.bundle_lift(z["gamma_eqf"], fn.xi0, fn.X, fn.Sigma[6:, 6:]) - z["Gamma"]).max() < 1e-9
    row = z["imu_row"]
    for f in (fc, fn):
        f.processIMUData(row[0], row[1:4], row[4:7])
        h, S = split_snapshot(f.get_snapshot())
        hg, Sg = split_snapshot(z["snap_after_imu"])
        assert rel(S, Sg) < 1e-13 and np.abs(h - hg).max() < 1e-12
        f.processVisionData(float(z["vision_stamp"]), z["ids"], z["bearings"])
        h, S = split_snapshot(f.get_snapshot())
        hg, Sg = split_snapshot(z["snap_after_vision"])
        assert rel(S, Sg) < 1e-9 and np.abs(h - hg).max() < 1e-8, (rel(S, Sg), np.abs(h - hg).max())


def test_silent_skips():
    """Reference quirks: first IMU sample only initialises; dt <= 0 skipped;
    vision before any IMU is dropped; vision with stamp <= currentTime is dropped entirely."""
    s = template_settings()
    for make in (lambda: COracleFilter(s), lambda: onp.VIOFilter(np_settings(s))):
        f = make()
        y = np.array([[0.0, 0.6, 0.8]])
        assert f.processVisionData(0.5, [0], y) == onp.SKIPPED_DT
        assert f.processIMUData(1.0, [0, 0, 0], [0.1, 0.2, 9.8]) == onp.SKIPPED_DT
        assert f.processIMUData(1.0, [0, 0, 0], [0.1, 0.2, 9.8]) == onp.SKIPPED_DT
        assert f.processVisionData(0.9, [0], y) == onp.SKIPPED_DT
        assert f.processIMUData(1.005, [0, 0, 0], [0.1, 0.2, 9.8]) == onp.OK
        assert f.processVisionData(1.0075, [0], y) == onp.OK
        assert f.getTime() == 1.0075
