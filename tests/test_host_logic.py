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


def _run_bench(*args, env_extra=None):
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(root, "bench.py"), *args], capture_output=True, text=True, env=env, timeout=600)


def test_bench_reference_arm_prints_one_json_line():
    """The reference arm (the CPU restatement of the reference timed on the host cores) honours the bench contract:
    exactly ONE line on stdout, JSON, with the keys the driver reads — also when a library writes to fd 1."""
    import json

    r = _run_bench("--impl", "reference", "--features", "8", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-400:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:400]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("N=8")


def test_bench_reference_arm_other_ranks_do_no_work():
    """Under torchrun only rank 0 runs the reference arm; the other ranks exit 0 without output."""
    r = _run_bench("--impl", "reference", "--features", "8", "--steps", "1", "--warmup", "1", "--gpus", "2",
                   env_extra={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_product_arm_fails_loudly_without_a_gpu():
    """No CUDA device: the product arm must not fall back to anything."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run_bench("--features", "8", "--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
