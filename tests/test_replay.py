"""The reference's wire formats and replay merge loop (eqf_vio/src/main.cpp:42-203), §8 row f-3."""
import numpy as np
import pytest

from eqf_vio_b200 import replay as rp
from eqf_vio_b200.settings import conditioned_settings
from eqf_vio_b200.synthetic import period_sequence
from helpers import rel, split_snapshot
from oracle.c_oracle import COracleFilter


class _OracleAdapter:
    """Gives the C oracle the attribute-style state estimate the replay formatter expects."""

    def __init__(self, s):
        self.f = COracleFilter(s)

    def processIMUData(self, *a):
        return self.f.processIMUData(*a)

    def processVisionData(self, *a):
        return self.f.processVisionData(*a)

    def getTime(self):
        return self.f.getTime()

    def get_snapshot(self):
        return self.f.get_snapshot()

    def stateEstimate(self):
        from eqf_vio_b200.filter import VIOStateEstimate

        e = self.f.stateEstimate()
        return VIOStateEstimate(e["pose"], e["velocity"], e["cameraOffset"], e["ids"], e["landmarks"])


def _write_files(tmp_path, seq):
    imu_p, meas_p = str(tmp_path / "imu.csv"), str(tmp_path / "meas.csv")
    rp.write_imu_csv(imu_p, seq.imu)
    rp.write_meas_csv(meas_p, seq.vision_stamps, seq.ids, seq.bearings)
    return imu_p, meas_p


def test_csv_round_trip_and_merge_order(tmp_path):
    s = conditioned_settings()
    seq = period_sequence(7, 3, camera_offset=tuple(s.cameraOffset))
    imu_p, meas_p = _write_files(tmp_path, seq)
    imu, meas = rp.read_imu_csv(imu_p), rp.read_meas_csv(meas_p)
    assert np.array_equal(imu, seq.imu) and len(meas) == len(seq.vision_stamps)
    assert all(np.array_equal(m[1], seq.ids) and np.array_equal(m[2], seq.bearings[j]) and m[0] == seq.vision_stamps[j] for j, m in enumerate(meas))
    calls = []

    class Rec:
        def processIMUData(self, t, *_):
            calls.append(("imu", t))

        def processVisionData(self, t, *_):
            calls.append(("vision", t))

    n_imu, n_vis = rp.replay(Rec(), imu, meas, start_time=0.0)
    # same order as the generator's merge (imu while imu.stamp < meas.stamp), loop ends when a file runs out
    want = [(k, (seq.imu[i, 0] if k == "imu" else seq.vision_stamps[i])) for k, i in seq.events()]
    # (the reference breaks out as soon as the IMU file is exhausted, main.cpp:120-122, so a trailing frame is never fed)
    assert calls == want[: len(calls)] and n_imu == len(seq.imu) and n_vis == len(seq.vision_stamps) - 1
    # startTime skips early rows
    calls.clear()
    rp.replay(Rec(), imu, meas, start_time=float(seq.vision_stamps[1]))
    assert all(t > seq.vision_stamps[1] for _, t in calls)


def test_replay_outputs_with_oracle(tmp_path):
    s = conditioned_settings()
    seq = period_sequence(6, 3, camera_offset=tuple(s.cameraOffset))
    imu_p, meas_p = _write_files(tmp_path, seq)
    rows_state, rows_filter = [], []
    f = _OracleAdapter(s)
    rp.replay(f, rp.read_imu_csv(imu_p), rp.read_meas_csv(meas_p), 0.0,
              on_state=lambda t, e: rows_state.append(rp.format_state_row(t, e)),
              on_filter=lambda t, sn: rows_filter.append(rp.format_filter_row(t, sn)))
    assert len(rows_state) == len(seq.vision_stamps) - 1 == len(rows_filter)
    c = [x.strip() for x in rows_state[-1].split(",")]
    assert float(c[0]) == seq.vision_stamps[-2] and int(c[11]) == 6 and len(c) == 12 + 4 * 6
    cf = [x.strip() for x in rows_filter[-1].split(",")]
    n = 11 + 3 * 6
    assert int(cf[21]) == 6 and len(cf) == 22 + 9 * 6 + n * n
    # Sigma is written row-major: first entries are Sigma(1,1), Sigma(1,2)
    _, S = split_snapshot(f.get_snapshot())
    assert abs(float(cf[22 + 54]) - S[0, 0]) <= 1e-4 * abs(S[0, 0]) and abs(float(cf[22 + 54 + 1]) - S[0, 1]) <= 1e-4 * abs(S[0, 1]) + 1e-12


@pytest.mark.gpu
def test_replay_cli_matches_oracle(tmp_path):
    import yaml

    from eqf_vio_b200.settings import TEMPLATE_EQF

    node = dict(TEMPLATE_EQF)
    node.update(outlierThreshold=1e9, initialSceneDepth=8.0, initialPointVariance=100.0)
    cfg = str(tmp_path / "cfg.yaml")
    with open(cfg, "w") as fh:
        yaml.safe_dump({"eqf": node, "main": {"startTime": 0.0, "writeState": True, "writeFilter": True}}, fh)
    s = conditioned_settings()
    seq = period_sequence(9, 4, camera_offset=tuple(s.cameraOffset))
    imu_p, meas_p = _write_files(tmp_path, seq)
    out_s, out_f = str(tmp_path / "state.csv"), str(tmp_path / "filter.csv")
    assert rp.main([imu_p, meas_p, cfg, "--out-state", out_s, "--out-filter", out_f]) == 0
    f = _OracleAdapter(s)
    rows = []
    rp.replay(f, rp.read_imu_csv(imu_p), rp.read_meas_csv(meas_p), 0.0, on_state=lambda t, e: rows.append(rp.format_state_row(t, e)))
    got = open(out_s).read().strip().split("\n")[1:]
    assert len(got) == len(rows)
    for a, b in zip(got, rows):
        va, vb = np.array(a.split(","), dtype=float), np.array(b.split(","), dtype=float)
        assert np.allclose(va, vb, rtol=2e-4, atol=1e-9)  # 5 significant digits in the file format
