"""Multi-GPU path = independent sessions, one per rank, plus an all-gather of the 8-double pose record
(SURVEY.md §8e).  The collective plumbing is exercised here with world_size 2 on the gloo backend; the
filter itself is replaced by the CPU oracle (this is a test, the product path has no CPU mode)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eqf_vio_b200.sessions import gather_pose_records, session_seed


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from eqf_vio_b200.settings import template_settings
    from eqf_vio_b200.synthetic import period_sequence
    from oracle.c_oracle import COracleFilter

    s = template_settings(outlierThreshold=1e9)
    seq = period_sequence(5, 2, seed=session_seed(5, rank), camera_offset=tuple(s.cameraOffset))
    f = COracleFilter(s)
    recs = []
    for kind, i in seq.events():
        if kind == "imu":
            f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else:
            f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
            rec = torch.tensor(np.concatenate([[f.getTime()], f.stateEstimate()["pose"]]))
            recs.append(gather_pose_records(rec).numpy())
    if rank == 0:
        np.save(out, np.stack(recs))
    dist.destroy_process_group()


def test_pose_gather_world2(tmp_path):
    out = str(tmp_path / "recs.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    recs = np.load(out)  # frames x ranks x 8
    assert recs.shape == (3, 2, 8)
    assert np.allclose(recs[:, 0, 0], recs[:, 1, 0])          # same stamps
    assert not np.allclose(recs[-1, 0, 1:4], recs[-1, 1, 1:4])  # different sessions (seeds)
    assert np.allclose(np.linalg.norm(recs[:, :, 4:8], axis=2), 1.0, atol=1e-9)


def test_session_seed():
    assert session_seed(256, 0) == 1256 and session_seed(256, 3) == 1259
