"""The restated oracle against the reference ITSELF: pvangoor/eqf_vio's own translation units compiled
against the Eigen-API stand-in (oracle/refshim -> oracle/_ref/libeqvio_ref.so, built by
__graft_entry__.build() where /root/reference is mounted; the prebuilt .so travels to the GPU box).
Skipped only if that library is absent."""
import numpy as np
import pytest

from eqf_vio_b200.settings import conditioned_settings, template_settings
from eqf_vio_b200.synthetic import period_sequence
from helpers import feed, rel, split_snapshot
from oracle import ref_binding
from oracle.c_oracle import COracleFilter

pytestmark = pytest.mark.skipif(not (ref_binding.available() or ref_binding.build()), reason="oracle/_ref not built (needs /root/reference)")


def test_free_functions_match_reference():
    s = template_settings(outlierThreshold=1e9)
    seq = period_sequence(10, 2, camera_offset=tuple(s.cameraOffset))
    o = COracleFilter(s)
    for kind, i in seq.events():
        if (kind, i) == ("vision", 2):
            break
        feed(o, seq, kind, i)
    f = ref_binding.ReferenceFilter(s)
    f.set_snapshot(o.get_snapshot())
    assert np.array_equal(f.get_snapshot(), o.get_snapshot())
    om = np.array([0.2, -0.1, 0.3])
    assert np.abs(f.state_matrix_A(om) - o.state_matrix_A(om)).max() < 1e-13   # EqFMatrices.cpp:277
    assert np.abs(f.input_matrix_B() - o.input_matrix_B()).max() < 1e-13       # :346
    assert np.abs(f.output_matrix_C() - o.output_matrix_C()).max() < 1e-14     # :319
    g = np.random.default_rng(1).standard_normal(5 + 30) * 1e-2
    assert np.abs(f.bundle_lift(g) - o.bundle_lift(g)).max() < 1e-12           # :173
    assert np.abs(f.delta(seq.bearings[2]) - o.build_C_delta(seq.bearings[2])[1]).max() < 1e-14


@pytest.mark.parametrize("mode", [{}, {"fastRiccati": True}, {"useDiscreteVelocityLift": False, "useDiscreteInnovationLift": False}, {"useInnovationLift": False}],
                         ids=["default", "fastRiccati", "continuous", "nolift"])
def test_sequences_match_reference(mode):
    s = conditioned_settings(**mode)
    seq = period_sequence(16, 10, camera_offset=tuple(s.cameraOffset))
    f, o = ref_binding.ReferenceFilter(s), COracleFilter(s)
    for kind, i in seq.events():
        feed(f, seq, kind, i), feed(o, seq, kind, i)
        if kind == "vision":
            h1, S1 = split_snapshot(f.get_snapshot())
            h2, S2 = split_snapshot(o.get_snapshot())
            assert rel(S1, S2) < 1e-12 and np.abs(h1 - h2).max() < 1e-11, (i, rel(S1, S2), np.abs(h1 - h2).max())


def test_bookkeeping_matches_reference():
    """removeOldLandmarks / matchMeasurementsToState / removeOutliers / addNewLandmarks with the template
    outlier threshold and ragged id sets: identical landmark sets at every frame."""
    rng = np.random.default_rng(3)
    s = template_settings()
    seq = period_sequence(14, 10, camera_offset=tuple(s.cameraOffset))
    f, o = ref_binding.ReferenceFilter(s), COracleFilter(s)
    for kind, i in seq.events():
        if kind == "imu":
            feed(f, seq, kind, i), feed(o, seq, kind, i)
            continue
        sel = np.sort(rng.choice(14, size=int(rng.integers(4, 14)), replace=False)) if i > 0 else np.arange(9)
        feed(f, seq, kind, i, sel=sel), feed(o, seq, kind, i, sel=sel)
        a, b = f.get_snapshot(), o.get_snapshot()
        assert a.size == b.size
        h1, S1 = split_snapshot(a)
        h2, S2 = split_snapshot(b)
        assert np.array_equal(h1[49::9], h2[49::9])
        assert rel(S1, S2) < 5e-9
