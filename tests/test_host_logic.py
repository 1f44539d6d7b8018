"""Host-side logic that needs no GPU: settings mirror, YAML key handling, synthetic generator, GEMM tile
heuristic inputs, replay merge order."""
import numpy as np
import pytest

from eqf_vio_b200.settings import TEMPLATE_EQF, default_settings, settings_from_eqf_node, template_settings
from eqf_vio_b200.synthetic import make_sequence, period_sequence


def test_defaults_and_template():
    d = default_settings()
    assert d.measurementVariance == 0.1 and d.initialPointVariance == 1.0 and d.outlierThreshold == 0.01
    assert d.fastRiccati == 0 and d.useInnovationLift == 1
    t = template_settings()
    assert t.initialPointVariance == 5000.0 and t.measurementVariance == 0.003 and t.velOmegaVariance == 1e-4
    assert tuple(t.cameraOffset)[3] == pytest.approx(0.7123014606690344)
    # absent keys keep the struct default (safeConfig, reference libs/core/include/common.h:22-29)
    s = settings_from_eqf_node({"fastRiccati": True})
    assert s.fastRiccati == 1 and s.measurementVariance == 0.1
    with pytest.raises(ValueError):
        settings_from_eqf_node({"cameraOffset": ["wx", 0, 0, 0, 1, 0, 0, 0]})
    assert set(TEMPLATE_EQF) >= {"initialSceneDepth", "outlierThreshold", "cameraOffset"}


def test_synthetic_sequence_shape_and_constraints():
    s = template_settings()
    seq = make_sequence(32, 1.0, camera_offset=tuple(s.cameraOffset))
    assert seq.imu.shape[1] == 7 and np.all(np.diff(seq.imu[:, 0]) > 0)
    assert np.allclose(np.diff(seq.imu[:, 0]), 1 / 200.0)
    assert np.allclose(np.diff(seq.vision_stamps), 1 / 20.0)
    # vision stamps sit half an IMU period off the IMU grid: dt > 0 always
    assert np.min(np.abs(seq.vision_stamps[:, None] - seq.imu[None, :, 0])) > 2e-3
    assert np.allclose(np.linalg.norm(seq.bearings, axis=2), 1.0)
    assert np.all(np.diff(seq.ids) > 0)
    # chart pole away from the camera +z axis for the first frame
    assert np.all(seq.bearings[0][:, 2] < np.cos(5e-3))
    # gravity visible in the first accelerometer sample
    assert abs(np.linalg.norm(seq.imu[0, 4:7]) - 9.81) < 1.0
    # same seed -> same data
    seq2 = make_sequence(32, 1.0, camera_offset=tuple(s.cameraOffset))
    assert np.array_equal(seq.bearings, seq2.bearings) and np.array_equal(seq.imu, seq2.imu)


def test_period_sequence_event_order():
    seq = period_sequence(5, 3)
    ev = [k for k, _ in seq.events()]
    assert ev[:3] == ["imu", "imu", "vision"]
    assert ev.count("vision") == 4 and ev[-1] == "vision"
    # 10 IMU ticks between consecutive vision frames (IMU 200 Hz / vision 20 Hz)
    idx = [i for i, k in enumerate(ev) if k == "vision"]
    assert all(b - a == 11 for a, b in zip(idx, idx[1:]))
