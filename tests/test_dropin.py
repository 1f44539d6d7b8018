"""The reference-typed drop-in `class VIOFilterB200` (include/eqf_vio_b200/VIOFilterB200.h) against the reference's own
`class VIOFilter` (eqf_vio/include/eqf_vio/VIOFilter.h:64-88), both driven by the replay loop of the reference's
driver (eqf_vio/src/main.cpp:108-170) in ONE executable, `oracle/_ref/dropin_replay` (source: oracle/refshim/
dropin_main.cpp; built by `make -C oracle/refshim` in THIS container, where /root/reference is mounted, from the
reference's unmodified sources + the Eigen stand-in; the binary travels to the GPU box with the snapshot).

    --impl ref   runs the reference class;   --impl b200   runs the drop-in over libeqvio_b200.so (sm_100a kernels).

The GPU test feeds both the same CSV files and compares the two output files the loop writes."""
import os
import subprocess

import numpy as np
import pytest

from eqf_vio_b200 import replay as rp
from eqf_vio_b200.settings import conditioned_settings, template_settings
from eqf_vio_b200.synthetic import period_sequence

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_replay")
HEADER = os.path.join(ROOT, "include", "eqf_vio_b200", "VIOFilterB200.h")


def write_config(path, s, start_time=0.0):
    """The flat two-section YAML subset EQVIO_config_template.yaml uses (what dropin_main.cpp reads)."""
    with open(path, "w") as f:
        f.write("eqf:\n")
        for k, v in s.as_dict().items():
            if k == "cameraOffset":
                f.write('  cameraOffset: ["xw", %s]\n' % ", ".join(repr(float(x)) for x in v))
            elif isinstance(v, tuple):
                f.write("  %s: [%s]\n" % (k, ", ".join(repr(float(x)) for x in v)))
            elif isinstance(v, bool):
                f.write("  %s: %s\n" % (k, "true" if v else "false"))
            else:
                f.write("  %s: %r\n" % (k, float(v)))
        f.write("main:\n  startTime: %r\n  writeState: true\n  writeFilter: true\n" % start_time)


def run_replay(impl, d, tag, extra=()):
    out_s, out_f = str(d / f"state_{tag}.csv"), str(d / f"filter_{tag}.csv")
    r = subprocess.run([BIN, "--impl", impl, str(d / "imu.csv"), str(d / "meas.csv"), str(d / "cfg.yaml"), out_s, out_f, *extra],
                       capture_output=True, text=True, timeout=900)
    return r, out_s, out_f


def rows_of(path):
    with open(path) as f:
        return [np.array([float(c) for c in line.split(",")]) for line in list(f)[1:]]


def prepare(tmp_path, s, N, periods, sel=None):
    seq = period_sequence(N, periods, camera_offset=tuple(s.cameraOffset))
    rp.write_imu_csv(str(tmp_path / "imu.csv"), seq.imu)
    if sel is None:
        rp.write_meas_csv(str(tmp_path / "meas.csv"), seq.vision_stamps, seq.ids, seq.bearings)
    else:   # ragged: a different id subset per frame
        with open(tmp_path / "meas.csv", "w") as f:
            f.write("t, N, id1, x1, y1, z1, ...\n")
            for j, t in enumerate(seq.vision_stamps):
                ids = sel(j)
                parts = [repr(float(t)), str(len(ids))]
                for i in ids:
                    parts += [str(int(i))] + [repr(float(v)) for v in seq.bearings[j][i]]
                f.write(", ".join(parts) + "\n")
    write_config(str(tmp_path / "cfg.yaml"), s)
    return seq


needs_binary = pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/dropin_replay not built (needs /root/reference: python __graft_entry__.py)")


@needs_binary
def test_reference_side_of_the_driver_matches_the_oracle(tmp_path):
    """The driver's reference arm (the reference's own class through the reference's own loop and operator<<) against
    the C restatement fed the same rows through the Python replay loop: pins the driver itself, on CPU."""
    from oracle.c_oracle import COracleFilter

    s = conditioned_settings()
    seq = prepare(tmp_path, s, 6, 4)
    r, out_s, out_f = run_replay("ref", tmp_path, "ref", ("--precision", "17"))
    assert r.returncode == 0, r.stderr
    assert "Processed 42 IMU and 4 vision measurements." in r.stdout
    o = COracleFilter(s)
    want = []
    rp.replay(o, rp.read_imu_csv(str(tmp_path / "imu.csv")), rp.read_meas_csv(str(tmp_path / "meas.csv")), 0.0,
              on_filter=lambda t, sn: want.append(np.concatenate([[t], sn])))
    got = rows_of(out_f)
    assert len(got) == len(want) == 4
    for g, w in zip(got, want):
        N = int(w[1])
        n = 11 + 3 * N
        assert g[0] == w[0] and int(g[21]) == N and g.size == 22 + 9 * N + n * n
        S = w[1 + 49 + 9 * N:].reshape(n, n, order="F")
        assert np.allclose(g[22 + 9 * N:].reshape(n, n), S, rtol=1e-10, atol=1e-13)   # Sigma is written row-major
        assert np.allclose(g[1:4], w[1 + 26:1 + 29], atol=1e-12) and np.allclose(g[4:8], w[1 + 22:1 + 26], atol=1e-12)


@needs_binary
def test_dropin_fails_loudly_without_a_gpu(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    prepare(tmp_path, conditioned_settings(), 4, 1)
    r, _, _ = run_replay("b200", tmp_path, "b200")
    assert r.returncode == 4 and "no CUDA device" in r.stderr      # no CPU fallback behind the drop-in class


def test_header_is_complete():
    """No elisions: every Settings field of VIOFilterSettings.h:29-50 is converted by name, every public member of
    VIOFilter.h:64-88 exists."""
    src = open(HEADER).read()
    from eqf_vio_b200.settings import _BOOL_FIELDS, _DOUBLE_FIELDS

    for name in list(_DOUBLE_FIELDS) + list(_BOOL_FIELDS) + ["initialAccelBias", "initialOmegaBias", "cameraOffset"]:
        assert f"s.{name}" in src, name
    for member in ("VIOFilterB200()", "VIOFilterB200(const AuxiliaryFilterData& auxiliaryData)",
                   "VIOFilterB200(const AuxiliaryFilterData& auxiliaryData, const VIOFilter::Settings& settings_)",
                   "VIOFilterB200(const VIOFilter::Settings& settings_)", "void initialiseFromIMUData(const IMUVelocity& imuVelocity)",
                   "void reset()", "void setAuxiliaryData(const AuxiliaryFilterData& auxiliaryData)",
                   "void setInertialPoints(const std::vector<Point3d>& inertialPoints)", "void processIMUData(const IMUVelocity& imuVelocity)",
                   "void processVisionData(const VisionMeasurement& measurement)", "double getTime() const", "VIOState stateEstimate() const",
                   "Eigen::MatrixXd stateCovariance() const", "friend std::ostream& operator<<(std::ostream& os, const VIOFilterB200& filter)",
                   "std::unique_ptr<VIOFilter::Settings> settings;"):
        assert member in src, member
    assert "..." not in src.replace("p1z, ...", "")


def _compare(tmp_path, extra, sigma_tol, state_tol, periods_checked=None):
    ra, sa, fa = run_replay("ref", tmp_path, "ref", ("--precision", "17", *extra))
    rb, sb, fb = run_replay("b200", tmp_path, "b200", ("--precision", "17", *extra))
    assert ra.returncode == 0 and rb.returncode == 0, (ra.stderr, rb.stderr)
    assert ra.stdout == rb.stdout                      # "Processed K IMU and M vision measurements."
    A, B = rows_of(fa), rows_of(fb)
    assert len(A) == len(B) > 0
    worst_s = worst_h = 0.0
    for a, b in zip(A, B):
        assert a.size == b.size and a[0] == b[0] and a[21] == b[21]
        N = int(a[21])
        assert np.array_equal(a[22:22 + 9 * N:9], b[22:22 + 9 * N:9])           # same ids, same order
        Sa, Sb = a[22 + 9 * N:], b[22 + 9 * N:]
        worst_s = max(worst_s, float(np.linalg.norm(Sa - Sb) / np.linalg.norm(Sa)))
        worst_h = max(worst_h, float(np.abs(a[1:22 + 9 * N] - b[1:22 + 9 * N]).max()))
    assert worst_s < sigma_tol and worst_h < state_tol, (worst_s, worst_h)
    SA, SB = rows_of(sa), rows_of(sb)
    assert len(SA) == len(SB) == len(A)
    for a, b in zip(SA, SB):
        assert a.size == b.size and np.abs(a - b).max() < state_tol
    return worst_s, worst_h


@pytest.mark.gpu
@needs_binary
def test_dropin_matches_reference_class_through_the_reference_loop(tmp_path):
    """N = 24, 12 vision periods, conditioned start-up, full precision: Sigma rel-Frobenius < 1e-9 and every state entry
    of both output files < 1e-8 at every frame (north_star tolerance)."""
    prepare(tmp_path, conditioned_settings(), 24, 12)
    _compare(tmp_path, (), 1e-9, 1e-8)
    # and the files as the reference's loop writes them (setprecision(5)): equal up to the last printed digit
    ra, sa, fa = run_replay("ref", tmp_path, "ref5")
    rb, sb, fb = run_replay("b200", tmp_path, "b2005")
    assert ra.returncode == 0 and rb.returncode == 0
    for pa, pb in ((sa, sb), (fa, fb)):
        for a, b in zip(rows_of(pa), rows_of(pb)):
            assert a.size == b.size and np.allclose(a, b, rtol=2e-5, atol=1e-9)


@pytest.mark.gpu
@needs_binary
def test_dropin_other_constructors_and_live_settings(tmp_path):
    """--aux: VIOFilter(aux, settings), move construction / assignment, setAuxiliaryData, initialiseFromIMUData, a
    settings change through the public pointer, stateCovariance() before any data — same calls on both classes."""
    prepare(tmp_path, conditioned_settings(), 10, 6)
    _compare(tmp_path, ("--aux",), 1e-9, 1e-8)


@pytest.mark.gpu
@needs_binary
def test_dropin_with_landmark_bookkeeping_and_template_settings(tmp_path):
    """Template settings as shipped (outlierThreshold 0.01, variance 5000, depth 1 m) with a different id subset in every
    frame: the reference's removeOldLandmarks / removeOutliers / addNewLandmarks against the drop-in's; identical id lists
    at every frame, free-running tolerance of the template start-up (tests/helpers.golden_tolerances)."""
    rng = np.random.default_rng(11)
    sets = [np.arange(9)] + [np.sort(rng.choice(14, size=int(rng.integers(4, 14)), replace=False)) for _ in range(8)]
    prepare(tmp_path, template_settings(), 14, 8, sel=lambda j: sets[j])
    _compare(tmp_path, (), 5e-9, 2e-7)
