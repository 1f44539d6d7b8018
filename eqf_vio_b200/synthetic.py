"""Seeded synthetic IMU + bearing sequences for parity tests and the benchmark.

A smooth analytic trajectory gives body-frame angular velocity and specific force
(accel = R^T (p'' + g e3), the convention of liftVelocity, reference eqf_vio/src/VIOGroup.cpp:187),
N static landmarks in a 3-15 m shell give unit bearings in the camera frame.  IMU at 200 Hz, vision at
20 Hz offset by half an IMU period so dt > 0 always (eqf_vio/src/VIOFilter.cpp:151-152,235).  All N
landmarks are seen in every frame (the filter has no field-of-view test), ids are 0..N-1 ascending.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

GRAVITY = 9.81


def _rot_xyz(roll, pitch, yaw):
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _quat_to_mat(q):
    w, x, y, z = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
        ]
    )


@dataclass
class Sequence:
    imu: np.ndarray  # (K, 7): stamp, omega xyz, accel xyz
    vision_stamps: np.ndarray  # (M,)
    ids: np.ndarray  # (N,) int32
    bearings: np.ndarray  # (M, N, 3)
    landmarks_world: np.ndarray  # (N, 3)
    true_pose: np.ndarray  # (M, 7) x y z qw qx qy qz at the vision stamps (rotation as matrix->not needed)

    def events(self):
        """Merge order of the reference replay loop (eqf_vio/src/main.cpp:111-170): an IMU row is
        consumed while imu.stamp < meas.stamp, otherwise the vision row."""
        i = j = 0
        K, M = len(self.imu), len(self.vision_stamps)
        while i < K or j < M:
            if i < K and (j >= M or self.imu[i, 0] < self.vision_stamps[j]):
                yield ("imu", i)
                i += 1
            else:
                yield ("vision", j)
                j += 1


def trajectory(t):
    """Position, velocity, acceleration (world, z up) and rotation body->world at time t."""
    amp = np.array([1.0, 0.8, 0.3])
    w = np.array([0.5, 0.35, 0.7])
    ph = np.array([0.0, 0.9, 0.3])
    p = amp * np.sin(w * t + ph)
    a = -amp * w * w * np.sin(w * t + ph)
    R = _rot_xyz(0.15 * np.sin(0.5 * t + 1.0), 0.2 * np.sin(0.7 * t), 0.6 * np.sin(0.4 * t) + 0.1 * t)
    return p, a, R


def make_sequence(n_features: int, duration: float, seed: int | None = None, camera_offset=None, imu_rate=200.0,
                  vision_rate=20.0, imu_noise=1e-2, bearing_noise=1e-3, bias=(0.01, -0.008, 0.005, 0.02, 0.01, -0.015)) -> Sequence:
    rng = np.random.default_rng(1000 + n_features if seed is None else seed)
    if camera_offset is None:
        camera_offset = (0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0)
    x_ic = np.array(camera_offset[0:3], dtype=float)
    q = np.array(camera_offset[3:7], dtype=float)
    R_ic = _quat_to_mat(q / np.linalg.norm(q))

    dt = 1.0 / imu_rate
    K = int(round(duration * imu_rate)) + 1
    stamps = 0.01 + dt * np.arange(K)
    imu = np.zeros((K, 7))
    h = 1e-6
    b = np.array(bias, dtype=float)
    for k, t in enumerate(stamps):
        _, a, R = trajectory(t)
        _, _, R1 = trajectory(t + h)
        dR = R.T @ R1
        omega = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / (2 * h)
        accel = R.T @ (a + np.array([0.0, 0.0, GRAVITY]))
        imu[k, 0] = t
        imu[k, 1:4] = omega + b[0:3] + imu_noise * rng.standard_normal(3)
        imu[k, 4:7] = accel + b[3:6] + imu_noise * rng.standard_normal(3)

    # landmarks in a 3-15 m shell around the origin of the trajectory; reject initial bearings within
    # 1e-2 rad of the camera +z axis (chart pole singularity, eqf_vio/src/VIOState.cpp:243, SO3.cpp:159-161)
    vstep = 1.0 / vision_rate
    M = int(np.floor((duration - 0.5 * dt) / vstep)) + 1
    vstamps = stamps[0] + 0.5 * dt + dt + vstep * np.arange(M)
    vstamps = vstamps[vstamps < stamps[-1]]
    M = len(vstamps)
    p0, _, R0 = trajectory(vstamps[0])
    lm = np.zeros((n_features, 3))
    count = 0
    while count < n_features:
        d = rng.standard_normal(3)
        d /= np.linalg.norm(d)
        r = rng.uniform(3.0, 15.0)
        cand = p0 + r * d
        qc = R_ic.T @ (R0.T @ (cand - p0) - x_ic)
        if qc[2] / np.linalg.norm(qc) > np.cos(1e-2):
            continue
        lm[count] = cand
        count += 1

    bearings = np.zeros((M, n_features, 3))
    true_pose = np.zeros((M, 7))
    for j, t in enumerate(vstamps):
        p, _, R = trajectory(t)
        qc = ((lm - p) @ R - x_ic) @ R_ic  # rows: R_ic^T (R^T (lm - p) - x_ic)
        y = qc / np.linalg.norm(qc, axis=1, keepdims=True)
        y = y + bearing_noise * rng.standard_normal(y.shape)
        bearings[j] = y / np.linalg.norm(y, axis=1, keepdims=True)
        true_pose[j, 0:3] = p
    return Sequence(imu, vstamps, np.arange(n_features, dtype=np.int32), bearings, lm, true_pose)


def period_sequence(n_features: int, n_periods: int, seed: int | None = None, camera_offset=None) -> Sequence:
    """`n_periods` vision periods (10 IMU ticks + 1 vision frame each) after an initial
    IMU sample + vision frame that initialise the filter and add the landmarks (config 1 shape)."""
    s = make_sequence(n_features, duration=(n_periods + 1) * 0.05 + 0.006, seed=seed, camera_offset=camera_offset)
    keep = n_periods + 1
    s.vision_stamps, s.bearings, s.true_pose = s.vision_stamps[:keep], s.bearings[:keep], s.true_pose[:keep]
    s.imu = s.imu[s.imu[:, 0] < s.vision_stamps[-1]]
    return s
