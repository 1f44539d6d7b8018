"""Shared helpers for the tests (oracle-side; never imported by the product)."""
import numpy as np

from eqf_vio_b200.settings import Settings, template_settings
from oracle import eqvio_numpy as onp


def np_settings(s: Settings) -> onp.Settings:
    return onp.Settings(**s.as_dict())


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def feed(filt, seq, kind, i, ids=None, sel=None):
    if kind == "imu":
        return filt.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
    if sel is None:
        return filt.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
    return filt.processVisionData(seq.vision_stamps[i], seq.ids[sel], seq.bearings[i][sel])


def run(filt, seq, stop_before=None):
    """Feed a whole sequence; stop_before=("vision", j) stops before that event."""
    for kind, i in seq.events():
        if stop_before is not None and (kind, i) == stop_before:
            return
        feed(filt, seq, kind, i)


def split_snapshot(d):
    N = int(d[0])
    hn = 49 + 9 * N
    n = 11 + 3 * N
    return d[:hn], d[hn : hn + n * n].reshape(n, n, order="F")


# reference test/testing_utilities.cpp:23-40,57-78 (random elements; N = 5 ids in the reference tests)
def random_unit_quat(rng):
    q = rng.standard_normal(4)
    return q / np.linalg.norm(q)


def random_state(rng, ids):
    return onp.VIOState(
        onp.SE3(random_unit_quat(rng), rng.uniform(-1, 1, 3)),
        rng.uniform(-1, 1, 3),
        rng.uniform(-1, 1, (len(ids), 3)),
        list(ids),
        onp.SE3(random_unit_quat(rng), np.zeros(3)),
    )


def random_group(rng, ids):
    return onp.VIOGroup(
        onp.SE3(random_unit_quat(rng), rng.uniform(-1, 1, 3)),
        rng.uniform(-1, 1, 3),
        [onp.SOT3(random_unit_quat(rng), 5.0 * rng.uniform() + 1.0) for _ in ids],
        list(ids),
    )


def state_vec_diff(x1, x2):  # testing_utilities.cpp:42-55
    out = [onp.SE3.log(x1.pose.inverse() * x2.pose), x2.velocity - x1.velocity, (x2.landmarks - x1.landmarks).reshape(-1)]
    return np.concatenate(out)


def log_norm(X):  # testing_utilities.cpp:80-88
    return np.linalg.norm(onp.SE3.log(X.A)) + np.linalg.norm(X.w) + sum(np.linalg.norm(onp.SOT3.log(Q)) for Q in X.Q)


def manifold_distance(a, b):  # testing_utilities.cpp:102-112
    return np.linalg.norm(a.gravityDir - b.gravityDir) + np.linalg.norm(a.velocity - b.velocity) + sum(np.linalg.norm(p - q) for p, q in zip(a.landmarks, b.landmarks))


# ---- golden vectors (tests/golden/*.npz, generated from the reference's own sources) ----
import glob
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_paths(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def golden_settings(z):
    from eqf_vio_b200.settings import conditioned_settings

    ov = dict(eval(str(z["overrides"])))
    return conditioned_settings(**ov) if str(z["base"]) == "conditioned" else template_settings(**ov)


def golden_tolerances(z):
    """(Sigma rel-Frobenius, state abs).  Conditioned start-up: the recursion is well conditioned and
    fp64 implementations agree to ~1e-13.  Template start-up (initialPointVariance 5000, depth 1 m vs
    3-15 m): the reference's own formulas lose ~5 digits per update (numerical conditioning),
    so free-running sequences separate to ~1e-9 between any two fp64 implementations."""
    return (1e-11, 1e-10) if str(z["base"]) == "conditioned" else (5e-9, 2e-7)


def replay_golden(z, filt, check):
    """Feed the recorded inputs; call check(j, snapshot_expected) after vision frame j."""
    imu, vs, ids, y = z["imu"], z["vision_stamps"], z["ids"], z["bearings"]
    i = j = 0
    while i < len(imu) or j < len(vs):
        if i < len(imu) and (j >= len(vs) or imu[i, 0] < vs[j]):
            filt.processIMUData(imu[i, 0], imu[i, 1:4], imu[i, 4:7])
            i += 1
        else:
            sel = z[f"sel{j}"]
            filt.processVisionData(vs[j], ids[sel], y[j][sel])
            check(j, z[f"snap{j}"])
            j += 1
