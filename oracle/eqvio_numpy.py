"""
eqvio_numpy.py — second, independent CPU restatement (numpy, fp64) of the reference EqF-VIO filter.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
reference arm; never by the product path (eqf_vio_b200/).

It follows pvangoor/eqf_vio @ 0b1334ec function by function (citations are relative to
/root/reference) with the reference's operation order: dense products associated left to right,
explicit inverses (numpy.linalg.inv = LAPACK getrf/getri, the same LU-with-partial-pivoting family as
Eigen's PartialPivLU inverse).  Dense products go through OpenBLAS, so this restatement is also the
"reference-equivalent CPU path" timed as the CPU baseline: Eigen's GEMM is replaced by
a BLAS that is at least as fast.

PARITY PIN STATUS: see oracle/eqvio_oracle.h — the reference holds no golden vectors for the filter
recursion; pins are the reference's property tests (tests/test_oracle_properties.py), agreement of
this file with the C restatement, and the refshim build of the reference's own sources.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field

import numpy as np

GRAVITY_CONSTANT = 9.81  # eqf_vio/include/eqf_vio/IMUVelocity.h:22
SIGMA_BASE_SIZE = 11  # eqf_vio/include/eqf_vio/VIOFilter.h:28
E3 = np.array([0.0, 0.0, 1.0])


class SingularChart(ArithmeticError):
    """std::domain_error from SO3::SO3FromVectors (libs/core/src/SO3.cpp:160-161)."""


# ------------------------------------------------------------------------------------------------
# SO3 as a quaternion [w, x, y, z] with Eigen::Quaterniond semantics (libs/core/src/SO3.cpp)
# ------------------------------------------------------------------------------------------------
def skew(v):  # SO3.cpp:110-114
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def q_identity():
    return np.array([1.0, 0.0, 0.0, 0.0])


def q_mul(a, b):  # Eigen quaternion product (SO3.cpp:66-70), not renormalised
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by + ay * bw + az * bx - ax * bz,
            aw * bz + az * bw + ax * by - ay * bx,
        ]
    )


def q_inv(a):  # Eigen Quaternion::inverse = conjugate / squaredNorm (SO3.cpp:74)
    n2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]
    return np.array([a[0], -a[1], -a[2], -a[3]]) / n2


def q_rot(q, v):  # Eigen _transformVector (SO3.cpp:58)
    u = q[1:]
    uv = np.cross(u, v)
    uv = uv + uv
    return v + q[0] * uv + np.cross(u, uv)


def q_mat(q):  # Eigen toRotationMatrix (SO3.cpp:92)
    w, x, y, z = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array(
        [
            [1 - (tyy + tzz), txy - twz, txz + twy],
            [txy + twz, 1 - (txx + tzz), tyz - twx],
            [txz - twy, tyz + twx, 1 - (txx + tyy)],
        ]
    )


def q_from_mat(m):  # Eigen Quaterniond(Matrix3d) (SO3.cpp:100)
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4)
    if t > 0:
        t = np.sqrt(t + 1.0)
        q[0] = 0.5 * t
        t = 0.5 / t
        q[1] = (m[2, 1] - m[1, 2]) * t
        q[2] = (m[0, 2] - m[2, 0]) * t
        q[3] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[1 + i] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[k, j] - m[j, k]) * t
        q[1 + j] = (m[j, i] + m[i, j]) * t
        q[1 + k] = (m[k, i] + m[i, k]) * t
    return q


def so3_exp(w):  # SO3.cpp:122-140
    th = np.linalg.norm(w)
    if abs(th) >= 1e-8:
        A = np.sin(th) / th
        B = (1 - np.cos(th)) / th**2
    else:
        A, B = 1.0, 0.5
    wx = skew(w)
    return q_from_mat(np.eye(3) + A * wx + B * wx @ wx)


def so3_log(q):  # SO3.cpp:142-153
    R = q_mat(q)
    theta = np.arccos((np.trace(R) - 1.0) / 2.0)
    coefficient = 0.5
    if abs(theta) >= 1e-6:
        coefficient = theta / (2.0 * np.sin(theta))
    Om = coefficient * (R - R.T)
    return np.array([Om[2, 1], Om[0, 2], Om[1, 0]])


def so3_from_vectors(origin, dest):  # SO3.cpp:155-167
    a = origin / np.linalg.norm(origin)
    b = dest / np.linalg.norm(dest)
    v = np.cross(a, b)
    c = float(a @ b)
    if abs(1 + c) <= 1e-8:
        raise SingularChart("The vectors cannot be exactly opposing.")
    sv = skew(v)
    return q_from_mat(np.eye(3) + (sv + 1 / (1 + c) * sv @ sv))


# ------------------------------------------------------------------------------------------------
# SE3 (libs/core/src/SE3.cpp), SOT3 (libs/core/src/SOT3.cpp)
# ------------------------------------------------------------------------------------------------
@dataclass
class SE3:
    R: np.ndarray = field(default_factory=q_identity)
    x: np.ndarray = field(default_factory=lambda: np.zeros(3))

    def __mul__(self, other):
        if isinstance(other, SE3):  # SE3.cpp:73-78
            return SE3(q_mul(self.R, other.R), self.x + q_rot(self.R, other.x))
        return q_rot(self.R, other) + self.x  # SE3.cpp:65

    def inverse(self):  # SE3.cpp:82-85
        Ri = q_inv(self.R)
        return SE3(Ri, -q_rot(Ri, self.x))

    def adjoint(self):  # SE3.cpp:95-103
        Rm = q_mat(self.R)
        Ad = np.zeros((6, 6))
        Ad[0:3, 0:3] = Rm
        Ad[3:6, 0:3] = skew(self.x) @ Rm
        Ad[3:6, 3:6] = Rm
        return Ad

    def copy(self):
        return SE3(self.R.copy(), self.x.copy())

    @staticmethod
    def exp(u):  # SE3.cpp:139-164
        w, v = u[0:3], u[3:6]
        th = np.linalg.norm(w)
        if abs(th) >= 1e-12:
            A = np.sin(th) / th
            B = (1 - np.cos(th)) / th**2
            C = (1 - A) / th**2
        else:
            A, B, C = 1.0, 0.5, 1.0 / 6.0
        wx = skew(w)
        wx2 = wx @ wx
        R = np.eye(3) + A * wx + B * wx2
        V = np.eye(3) + B * wx + C * wx2
        return SE3(q_from_mat(R), V @ v)

    @staticmethod
    def log(P):  # SE3.cpp:166-189
        Om = skew(so3_log(P.R))
        theta = np.linalg.norm(np.array([Om[2, 1], Om[0, 2], Om[1, 0]]))
        coefficient = 1.0 / 12.0
        if abs(theta) > 1e-8:
            coefficient = 1 / (theta * theta) * (1 - (theta * np.sin(theta)) / (2 * (1 - np.cos(theta))))
        VInv = np.eye(3) - 0.5 * Om + coefficient * Om @ Om
        return np.concatenate([np.array([Om[2, 1], Om[0, 2], Om[1, 0]]), VInv @ P.x])


@dataclass
class SOT3:
    R: np.ndarray = field(default_factory=q_identity)
    a: float = 1.0

    def __mul__(self, other):
        if isinstance(other, SOT3):  # SOT3.cpp:69-74
            return SOT3(q_mul(self.R, other.R), self.a * other.a)
        return self.a * q_rot(self.R, other)  # SOT3.cpp:62

    def inverse(self):  # SOT3.cpp:78-81
        return SOT3(q_inv(self.R), 1.0 / self.a)

    def as_matrix3(self):  # SOT3.cpp:107-110
        return self.a * q_mat(self.R)

    def copy(self):
        return SOT3(self.R.copy(), float(self.a))

    @staticmethod
    def exp(w):  # SOT3.cpp:127-132
        return SOT3(so3_exp(w[0:3]), float(np.exp(w[3])))

    @staticmethod
    def log(T):  # SOT3.cpp:134-139
        return np.concatenate([so3_log(T.R), [np.log(T.a)]])


# ------------------------------------------------------------------------------------------------
# sphere charts (eqf_vio/src/VIOState.cpp:199-251)
# ------------------------------------------------------------------------------------------------
def e3_project_sphere(eta):  # :199-204
    return (eta - E3)[0:2] / (1 - eta[2])


def e3_project_sphere_inv(y):  # :206-211
    ybar = np.array([y[0], y[1], 0.0])
    return E3 + 2.0 / (ybar @ ybar + 1) * (ybar - E3)


def e3_project_sphere_diff(eta):  # :213-220
    D = (np.eye(3) * (1 - eta[2]) + np.outer(eta - E3, E3))[0:2, :]
    return (1 - eta[2]) ** -2.0 * D


def e3_project_sphere_inv_diff(y):  # :222-228
    n2 = y @ y
    D = np.zeros((3, 2))
    D[0:2, :] = np.eye(2) * (n2 + 1.0) - 2 * np.outer(y, y)
    D[2, :] = 2 * y
    return 2.0 * (n2 + 1.0) ** -2.0 * D


def stereo_sphere_chart(eta, pole):  # :230-234
    return e3_project_sphere(q_rot(so3_from_vectors(-pole, E3), eta))


def stereo_sphere_chart_inv(y, pole):  # :236-240
    return q_rot(q_inv(so3_from_vectors(-pole, E3)), e3_project_sphere_inv(y))


def stereo_sphere_chart_diff(eta, pole):  # :242-246
    rot = so3_from_vectors(-pole, E3)
    return e3_project_sphere_diff(q_rot(rot, eta)) @ q_mat(rot)


def stereo_sphere_chart_inv_diff(y, pole):  # :248-251
    rot = so3_from_vectors(-pole, E3)
    return q_mat(q_inv(rot)) @ e3_project_sphere_inv_diff(y)


# ------------------------------------------------------------------------------------------------
# state / group / algebra (VIOState.h:38-60, VIOGroup.h:24-45)
# ------------------------------------------------------------------------------------------------
@dataclass
class VIOState:
    pose: SE3 = field(default_factory=SE3)
    velocity: np.ndarray = field(default_factory=lambda: np.zeros(3))
    landmarks: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))  # bodyLandmarks[i].p
    ids: list = field(default_factory=list)
    cameraOffset: SE3 = field(default_factory=SE3)


@dataclass
class VIOManifoldState:
    gravityDir: np.ndarray
    velocity: np.ndarray
    landmarks: np.ndarray
    ids: list
    cameraOffset: SE3


def project_to_manifold(Xi: VIOState) -> VIOManifoldState:  # VIOState.cpp:88-95
    return VIOManifoldState(q_rot(q_inv(Xi.pose.R), E3), Xi.velocity.copy(), Xi.landmarks.copy(), list(Xi.ids), Xi.cameraOffset)


@dataclass
class VIOGroup:
    A: SE3 = field(default_factory=SE3)
    w: np.ndarray = field(default_factory=lambda: np.zeros(3))
    Q: list = field(default_factory=list)
    ids: list = field(default_factory=list)

    def __mul__(self, other):  # VIOGroup.cpp:92-110
        assert self.ids == other.ids
        return VIOGroup(self.A * other.A, self.w + q_rot(self.A.R, other.w), [a * b for a, b in zip(self.Q, other.Q)], list(self.ids))

    def inverse(self):  # VIOGroup.cpp:124-134
        return VIOGroup(self.A.inverse(), -q_rot(q_inv(self.A.R), self.w), [q.inverse() for q in self.Q], list(self.ids))

    @staticmethod
    def identity(ids=()):  # VIOGroup.cpp:112-122
        return VIOGroup(SE3(), np.zeros(3), [SOT3() for _ in ids], list(ids))


@dataclass
class VIOAlgebra:
    U: np.ndarray
    u: np.ndarray
    W: np.ndarray  # N x 4
    ids: list

    def __mul__(self, c):  # VIOGroup.cpp:136-145
        return VIOAlgebra(self.U * c, self.u * c, self.W * c, list(self.ids))

    __rmul__ = __mul__

    def __neg__(self):
        return VIOAlgebra(-self.U, -self.u, -self.W, list(self.ids))

    def __add__(self, o):
        return VIOAlgebra(self.U + o.U, self.u + o.u, self.W + o.W, list(self.ids))

    def __sub__(self, o):
        return self + (-o)


def state_group_action(X: VIOGroup, state):  # VIOGroup.cpp:23-69
    lm = np.array([Qi.inverse() * p for Qi, p in zip(X.Q, state.landmarks)]).reshape(-1, 3)
    RAinv = q_inv(X.A.R)
    if isinstance(state, VIOState):
        return VIOState(state.pose * X.A, q_rot(RAinv, state.velocity - X.w), lm, list(state.ids), state.cameraOffset)
    return VIOManifoldState(q_rot(RAinv, state.gravityDir), q_rot(RAinv, state.velocity - X.w), lm, list(state.ids), state.cameraOffset)


def output_group_action(X: VIOGroup, bearings):  # VIOGroup.cpp:71-90
    return np.array([q_rot(q_inv(Qi.R), y) for Qi, y in zip(X.Q, bearings)]).reshape(-1, 3)


def measure_system_state(state):  # VIOState.cpp:58-70
    return np.array([p / np.linalg.norm(p) for p in state.landmarks]).reshape(-1, 3)


def output_coordinate_chart(y, y0):  # VisionMeasurement.cpp:24-34
    return np.concatenate([stereo_sphere_chart(a, b) for a, b in zip(y, y0)]) if len(y) else np.zeros(0)


def output_coordinate_chart_inv(delta, y0):  # VisionMeasurement.cpp:36-50
    return np.array([stereo_sphere_chart_inv(delta[2 * i : 2 * i + 2], y0[i]) for i in range(len(y0))])


def euclid_coordinate_chart(xi, xi0):  # VIOState.cpp:97-110
    N = len(xi0.ids)
    eps = np.zeros(5 + 3 * N)
    eps[0:2] = stereo_sphere_chart(xi.gravityDir, xi0.gravityDir)
    eps[2:5] = xi.velocity - xi0.velocity
    eps[5:] = (xi.landmarks - xi0.landmarks).reshape(-1)
    return eps


def euclid_coordinate_chart_inv(eps, xi0):  # VIOState.cpp:112-128
    N = len(xi0.ids)
    return VIOManifoldState(
        stereo_sphere_chart_inv(eps[0:2], xi0.gravityDir),
        xi0.velocity + eps[2:5],
        xi0.landmarks + eps[5:].reshape(N, 3),
        list(xi0.ids),
        xi0.cameraOffset,
    )


def integrate_system_function(state: VIOState, omega, accel, dt):  # VIOState.cpp:26-56 (test only)
    poseVel = np.concatenate([omega, state.velocity])
    pose = state.pose * SE3.exp(dt * poseVel)
    vel = state.velocity + dt * (-skew(omega) @ state.velocity + accel - q_rot(q_inv(state.pose.R), np.array([0, 0, GRAVITY_CONSTANT])))
    U_C = state.cameraOffset.inverse().adjoint() @ poseVel
    camInv = SE3.exp(-dt * U_C)
    lm = np.array([camInv * p for p in state.landmarks]).reshape(-1, 3)
    return VIOState(pose, vel, lm, list(state.ids), state.cameraOffset)


def lift_velocity(state: VIOManifoldState, omega, accel) -> VIOAlgebra:  # VIOGroup.cpp:178-207
    U = np.concatenate([omega, state.velocity])
    u = -accel + state.gravityDir * GRAVITY_CONSTANT
    U_C = state.cameraOffset.inverse().adjoint() @ U
    om_C, v_C = U_C[0:3], U_C[3:6]
    W = np.zeros((len(state.ids), 4))
    for i, p in enumerate(state.landmarks):
        n2 = p @ p
        W[i, 0:3] = om_C + skew(p) @ v_C / n2
        W[i, 3] = p @ v_C / n2
    return VIOAlgebra(U, u, W, list(state.ids))


def lift_velocity_discrete(state: VIOManifoldState, omega, accel, dt) -> VIOGroup:  # VIOGroup.cpp:209-243
    AVel = np.concatenate([omega, state.velocity])
    A = SE3.exp(dt * AVel)
    w = state.velocity - q_rot(A.R, state.velocity + dt * (-skew(omega) @ state.velocity + accel - state.gravityDir * GRAVITY_CONSTANT))
    U_C = state.cameraOffset.inverse().adjoint() @ AVel
    camInv = SE3.exp(-dt * U_C)
    Q = []
    for p0 in state.landmarks:
        p1 = camInv * p0
        Q.append(SOT3(so3_from_vectors(p1 / np.linalg.norm(p1), p0 / np.linalg.norm(p0)), np.linalg.norm(p0) / np.linalg.norm(p1)))
    return VIOGroup(A, w, Q, list(state.ids))


def vio_exp(lam: VIOAlgebra) -> VIOGroup:  # VIOGroup.cpp:245-256
    return VIOGroup(SE3.exp(lam.U), lam.u.copy(), [SOT3.exp(Wi) for Wi in lam.W], list(lam.ids))


# ------------------------------------------------------------------------------------------------
# EqF matrices and innovation lifts (eqf_vio/src/EqFMatrices.cpp)
# ------------------------------------------------------------------------------------------------
def state_matrix_A(X: VIOGroup, xi0: VIOManifoldState, omega):  # :277-317
    N = len(xi0.ids)
    A0 = np.zeros((5 + 3 * N, 5 + 3 * N))
    A0[2:5, 0:2] = -stereo_sphere_chart_inv_diff(np.zeros(2), xi0.gravityDir) * GRAVITY_CONSTANT
    R_IC = q_mat(xi0.cameraOffset.R)
    R_Ahat = q_mat(X.A.R)
    for i in range(N):
        Qhat = q_mat(X.Q[i].R) * X.Q[i].a
        A0[5 + 3 * i : 8 + 3 * i, 2:5] = -Qhat @ R_IC.T @ R_Ahat.T
    xi_hat = state_group_action(X, xi0)
    U_I = np.concatenate([omega, xi_hat.velocity])
    v_C = (xi0.cameraOffset.inverse().adjoint() @ U_I)[3:6]
    for i in range(N):
        Qhat = q_mat(X.Q[i].R) * X.Q[i].a
        qh = xi_hat.landmarks[i]
        A_qi = -Qhat @ (skew(qh) @ skew(v_C) - 2 * np.outer(v_C, qh) + np.outer(qh, v_C)) @ np.linalg.inv(Qhat) * (1 / (qh @ qh))
        A0[5 + 3 * i : 8 + 3 * i, 5 + 3 * i : 8 + 3 * i] = A_qi
    return A0


def input_matrix_B(X: VIOGroup, xi0: VIOManifoldState):  # :346-382
    N = len(xi0.ids)
    Bt = np.zeros((5 + 3 * N, 6))
    xi_hat = state_group_action(X, xi0)
    R_A = q_mat(X.A.R)
    Bt[0:2, 0:3] = stereo_sphere_chart_diff(xi0.gravityDir, xi0.gravityDir) @ R_A @ skew(xi_hat.gravityDir)
    Bt[2:5, 0:3] = R_A @ skew(xi_hat.velocity)
    Bt[2:5, 3:6] = R_A
    RT_IC = q_mat(q_inv(xi0.cameraOffset.R))
    x_IC = xi0.cameraOffset.x
    for i in range(N):
        Qhat = q_mat(X.Q[i].R) * X.Q[i].a
        qh = xi_hat.landmarks[i]
        Bt[5 + 3 * i : 8 + 3 * i, 0:3] = Qhat @ (skew(qh) @ RT_IC + RT_IC @ skew(x_IC))
    return Bt


def output_matrix_C(xi0):  # :319-344
    N = len(xi0.ids)
    C0 = np.zeros((2 * N, 5 + 3 * N))
    for i in range(N):
        qi0 = xi0.landmarks[i]
        yi0 = qi0 / np.linalg.norm(qi0)
        C0[2 * i : 2 * i + 2, 5 + 3 * i : 8 + 3 * i] = 1 / np.linalg.norm(qi0) * stereo_sphere_chart_diff(yi0, yi0) @ (np.eye(3) - np.outer(yi0, yi0))
    return C0


def lift_innovation(gamma, xi0: VIOManifoldState) -> VIOAlgebra:  # :35-67
    N = len(xi0.ids)
    U = np.zeros(6)
    U[0:3] = -skew(xi0.gravityDir) @ stereo_sphere_chart_inv_diff(np.zeros(2), xi0.gravityDir) @ gamma[0:2]
    u = -gamma[2:5] - skew(U[0:3]) @ xi0.velocity
    W = np.zeros((N, 4))
    for i in range(N):
        g = gamma[5 + 3 * i : 8 + 3 * i]
        q = xi0.landmarks[i]
        W[i, 0:3] = -np.cross(q, g) / (q @ q)
        W[i, 3] = -(q @ g) / (q @ q)
    return VIOAlgebra(U, u, W, list(xi0.ids))


def lift_total_space_innovation(Gamma, xi0: VIOState) -> VIOAlgebra:  # :69-96
    N = len(xi0.ids)
    U = Gamma[0:6].copy()
    u = -Gamma[6:9] - skew(U[0:3]) @ xi0.velocity
    W = np.zeros((N, 4))
    for i in range(N):
        g = Gamma[9 + 3 * i : 12 + 3 * i]
        q = xi0.landmarks[i]
        W[i, 0:3] = -np.cross(q, g) / (q @ q)
        W[i, 3] = -(q @ g) / (q @ q)
    return VIOAlgebra(U, u, W, list(xi0.ids))


def _wls(gamma, xi0: VIOState, X: VIOGroup, Sigma, DeltaU):
    """Shared least-squares core of bundleLift (:173-252) and liftInnovation/4 (:98-171)."""
    xiHat = state_group_action(X, xi0)
    eta0 = project_to_manifold(xi0).gravityDir
    eta0 = eta0 / np.linalg.norm(eta0)
    N = len(xi0.ids)
    KPara = np.zeros((6, 4))
    KPara[0:3, 0] = eta0
    KPara[3:6, 1:4] = np.eye(3)
    KPerp = np.zeros((6, 6))
    KPerp[0:3, 0:3] = np.eye(3) - np.outer(eta0, eta0)
    R_C = q_mul(xiHat.pose.R, xiHat.cameraOffset.R)
    R_CT = q_mat(q_inv(R_C))
    AdP0 = xi0.pose.adjoint()
    DUF = KPerp @ DeltaU
    coeff = np.zeros((3 * N, 4))
    obs = np.zeros(3 * N)
    D = np.zeros((5 + 3 * N, 3 * N))
    PT = xiHat.pose * xiHat.cameraOffset
    for i in range(N):
        g = gamma[5 + 3 * i : 8 + 3 * i]
        pHat = PT * xiHat.landmarks[i]
        alpha = -q_rot(R_C, X.Q[i].inverse() * g)
        pHatMat = np.hstack([-skew(pHat), np.eye(3)])
        obs[3 * i : 3 * i + 3] = alpha - pHatMat @ AdP0 @ DUF
        coeff[3 * i : 3 * i + 3, :] = pHatMat @ AdP0 @ KPara
        D[5 + 3 * i : 8 + 3 * i, 3 * i : 3 * i + 3] = X.Q[i].as_matrix3() @ R_CT
    Wm = D.T @ np.linalg.inv(Sigma) @ D
    sol = np.linalg.solve(coeff.T @ Wm @ coeff, coeff.T @ Wm @ obs)  # reference: 4x4 householderQr().solve
    return DUF + KPara @ sol


def bundle_lift(gamma, xi0: VIOState, X: VIOGroup, Sigma):  # :173-252
    N = len(xi0.ids)
    eta0 = project_to_manifold(xi0).gravityDir
    eta0 = eta0 / np.linalg.norm(eta0)
    DeltaU = np.zeros(6)
    DeltaU[0:3] = -skew(eta0) @ stereo_sphere_chart_inv_diff(np.zeros(2), eta0) @ gamma[0:2]
    DeltaU = _wls(gamma, xi0, X, Sigma, DeltaU)
    out = np.zeros(9 + 3 * N)
    out[0:6] = DeltaU
    out[6:] = gamma[2:]
    return out


def lift_innovation_wls(gamma, xi0: VIOState, X: VIOGroup, Sigma) -> VIOAlgebra:  # :98-171
    Delta = lift_innovation(gamma, project_to_manifold(xi0))
    Delta.U = _wls(gamma, xi0, X, Sigma, Delta.U)
    Delta.u = -gamma[2:5] - skew(Delta.U[0:3]) @ xi0.velocity
    return Delta


def lift_total_space_innovation_discrete(Gamma, xi0: VIOState) -> VIOGroup:  # :254-275
    A = SE3.exp(Gamma[0:6])
    w = xi0.velocity - q_rot(A.R, xi0.velocity + Gamma[6:9])
    Q = []
    for i, qi in enumerate(xi0.landmarks):
        q1 = qi + Gamma[9 + 3 * i : 12 + 3 * i]
        Q.append(SOT3(so3_from_vectors(q1 / np.linalg.norm(q1), qi / np.linalg.norm(qi)), np.linalg.norm(qi) / np.linalg.norm(q1)))
    return VIOGroup(A, w, Q, list(xi0.ids))


# ------------------------------------------------------------------------------------------------
# settings + filter (VIOFilterSettings.h:28-50, VIOFilter.cpp)
# ------------------------------------------------------------------------------------------------
@dataclass
class Settings:
    biasOmegaProcessVariance: float = 0.001
    biasAccelProcessVariance: float = 0.001
    gravityProcessVariance: float = 0.001
    velocityProcessVariance: float = 0.001
    pointProcessVariance: float = 0.001
    velOmegaVariance: float = 0.1
    velAccelVariance: float = 0.1
    measurementVariance: float = 0.1
    initialGravityVariance: float = 1.0
    initialVelocityVariance: float = 1.0
    initialPointVariance: float = 1.0
    initialBiasOmegaVariance: float = 1.0
    initialBiasAccelVariance: float = 1.0
    initialSceneDepth: float = 1.0
    outlierThreshold: float = 0.01
    useInnovationLift: bool = True
    useDiscreteInnovationLift: bool = True
    useDiscreteVelocityLift: bool = True
    fastRiccati: bool = False
    initialAccelBias: tuple = (0.0, 0.0, 0.0)
    initialOmegaBias: tuple = (0.0, 0.0, 0.0)
    cameraOffset: tuple = (0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0)  # x y z qw qx qy qz


OK, SKIPPED_DT, NOT_INITIALISED, EMPTY_MEASUREMENT = 0, 1, 2, 3


class VIOFilter:
    """VIOFilter (eqf_vio/src/VIOFilter.cpp), same method names."""

    def __init__(self, settings: Settings):  # :60-73
        s = self.settings = settings
        self.Sigma = np.eye(SIGMA_BASE_SIZE)
        self.Sigma[0:3, 0:3] = np.eye(3) * s.initialBiasOmegaVariance
        self.Sigma[3:6, 3:6] = np.eye(3) * s.initialBiasAccelVariance
        self.Sigma[6:8, 6:8] = np.eye(2) * s.initialGravityVariance
        self.Sigma[8:11, 8:11] = np.eye(3) * s.initialVelocityVariance
        co = s.cameraOffset
        self.xi0 = VIOState(SE3(), np.zeros(3), np.zeros((0, 3)), [], SE3(np.array(co[3:7], dtype=float), np.array(co[0:3], dtype=float)))
        self.X = VIOGroup.identity()
        self.inputBias = np.concatenate([np.array(s.initialOmegaBias, dtype=float), np.array(s.initialAccelBias, dtype=float)])
        self.initialisedFlag = False
        self.currentTime = -1.0
        self.curOmega = np.zeros(3)
        self.curAccel = np.zeros(3)
        self.accOmega = np.zeros(3)
        self.accAccel = np.zeros(3)
        self.accumulatedTime = 0.0
        self.counters = {"propagate": 0, "update": 0}

    def stateEstimate(self) -> VIOState:  # :304
        return state_group_action(self.X, self.xi0)

    def stateCovariance(self):  # :306-309
        return self.Sigma.copy()

    def getTime(self):  # :343
        return self.currentTime

    def processIMUData(self, stamp, omega, accel):  # :120-131
        uo = np.asarray(omega, dtype=float) - self.inputBias[0:3]
        ua = np.asarray(accel, dtype=float) - self.inputBias[3:6]
        if not self.initialisedFlag:  # :133-144
            self.xi0.pose = SE3()
            self.xi0.velocity = np.zeros(3)
            self.initialisedFlag = True
            self.xi0.pose.R = so3_from_vectors(ua / np.linalg.norm(ua), E3)
        r = self.integrateUpToTime(stamp, not self.settings.fastRiccati)
        self.curOmega, self.curAccel = uo, ua
        self.currentTime = stamp
        return OK if r else SKIPPED_DT

    def riccati(self, T, omega):  # :162-189
        s = self.settings
        N = len(self.xi0.ids)
        n = SIGMA_BASE_SIZE + 3 * N
        PMat = np.eye(n)
        PMat[0:3, 0:3] *= s.biasOmegaProcessVariance
        PMat[3:6, 3:6] *= s.biasAccelProcessVariance
        PMat[6:8, 6:8] *= s.gravityProcessVariance
        PMat[8:11, 8:11] *= s.velocityProcessVariance
        PMat[11:, 11:] *= s.pointProcessVariance
        xi0m = project_to_manifold(self.xi0)
        A0t = state_matrix_A(self.X, xi0m, omega)
        Bt = input_matrix_B(self.X, xi0m)
        R = np.eye(6)
        R[0:3, 0:3] *= s.velOmegaVariance
        R[3:6, 3:6] *= s.velAccelVariance
        Ab = np.zeros((n, n))
        Ab[6:, 6:] = A0t
        Ab[6:, 0:6] = -Bt
        F = np.eye(n) + Ab * T
        Bb = np.zeros((n, 6))
        Bb[6:, :] = Bt
        self.Sigma = T * (PMat + Bb @ R @ Bb.T) + F @ self.Sigma @ F.T
        self.counters["propagate"] += 1
        return F, Bb

    def integrateUpToTime(self, newTime, doRiccati=True):  # :146-209
        if self.currentTime < 0:
            return False
        dt = newTime - self.currentTime
        if dt <= 0:
            return False
        self.accumulatedTime += dt
        self.accOmega = self.accOmega + self.curOmega * dt
        self.accAccel = self.accAccel + self.curAccel * dt
        currentState = self.stateEstimate()
        if doRiccati:
            self.riccati(self.accumulatedTime, self.accOmega * (1.0 / self.accumulatedTime))
            self.accOmega = np.zeros(3)
            self.accAccel = np.zeros(3)
            self.accumulatedTime = 0.0
        cur_m = project_to_manifold(currentState)
        if self.settings.useDiscreteVelocityLift:
            self.X = self.X * lift_velocity_discrete(cur_m, self.curOmega, self.curAccel, dt)
        else:
            self.X = self.X * vio_exp(lift_velocity(cur_m, self.curOmega, self.curAccel) * dt)
        self.currentTime = newTime
        return True

    def _remove_landmark_at(self, idx):  # :421-427
        self.xi0.landmarks = np.delete(self.xi0.landmarks, idx, axis=0)
        del self.xi0.ids[idx]
        del self.X.ids[idx]
        del self.X.Q[idx]
        rows = list(range(SIGMA_BASE_SIZE + 3 * idx, SIGMA_BASE_SIZE + 3 * idx + 3))
        self.Sigma = np.delete(np.delete(self.Sigma, rows, axis=0), rows, axis=1)

    def processVisionData(self, stamp, ids, bearings):  # :232-302
        ids = [int(i) for i in ids]
        bearings = np.asarray(bearings, dtype=float).reshape(-1, 3)
        if not self.integrateUpToTime(stamp) or not self.initialisedFlag:
            return SKIPPED_DT
        assert all(ids[i] <= ids[i + 1] for i in range(len(ids) - 1))
        # removeOldLandmarks :393-419
        for li in reversed([i for i, sid in enumerate(self.X.ids) if sid not in ids]):
            self._remove_landmark_at(li)
        # matchMeasurementsToState :211-230
        m_ids = [None] * len(ids)
        m_y = np.zeros((len(ids), 3))
        newPos = len(self.X.ids) - 1
        for mid, y in zip(ids, bearings):
            if mid in self.X.ids:
                idx = self.X.ids.index(mid)
            else:
                newPos += 1
                idx = newPos
            m_ids[idx] = mid
            m_y[idx] = y
        # removeOutliers :429-443
        yHat = measure_system_state(self.stateEstimate())
        for i in range(len(yHat) - 1, -1, -1):
            if np.linalg.norm(m_y[i] - yHat[i]) > self.settings.outlierThreshold:
                self._remove_landmark_at(i)
                m_y = np.delete(m_y, i, axis=0)
                del m_ids[i]
        # addNewLandmarks :345-391
        oldN = len(self.X.ids)
        if len(m_ids) > oldN:
            lm = self.stateEstimate().landmarks
            median = self.settings.initialSceneDepth
            if oldN > 0:
                d2 = np.sort(np.sum(lm * lm, axis=1))
                median = d2[oldN // 2] ** 0.5
            new_p = m_y[oldN:] * median
            self.xi0.landmarks = np.vstack([self.xi0.landmarks.reshape(-1, 3), new_p])
            self.xi0.ids += m_ids[oldN:]
            self.X.ids += m_ids[oldN:]
            self.X.Q += [SOT3() for _ in m_ids[oldN:]]
            og = self.Sigma.shape[0]
            newN = len(m_ids) - oldN
            S = np.zeros((og + 3 * newN, og + 3 * newN))
            S[:og, :og] = self.Sigma
            S[og:, og:] = np.eye(3 * newN) * self.settings.initialPointVariance
            self.Sigma = S
        if len(m_ids) == 0:
            return EMPTY_MEASUREMENT

        C, delta = self.build_C_delta(m_y)
        N = len(self.xi0.ids)
        QMat = self.settings.measurementVariance * np.eye(2 * N)
        S = C @ self.Sigma @ C.T + QMat  # :276
        K = self.Sigma @ C.T @ np.linalg.inv(S)  # :277
        gamma = K @ delta  # :279
        g_eqf, g_bias = gamma[6:], gamma[0:6]
        if self.settings.useInnovationLift:
            Gamma = bundle_lift(g_eqf, self.xi0, self.X, self.Sigma[6:, 6:])  # :285
            if self.settings.useDiscreteInnovationLift:
                Delta = lift_total_space_innovation_discrete(Gamma, self.xi0)
            else:
                Delta = vio_exp(lift_total_space_innovation(Gamma, self.xi0))
        else:
            Delta = vio_exp(lift_innovation(g_eqf, project_to_manifold(self.xi0)))
        self.inputBias = self.inputBias + g_bias  # :295
        self.X = Delta * self.X  # :296
        self.Sigma = self.Sigma - K @ C @ self.Sigma  # :297
        self.counters["update"] += 1
        self.last = {"K": K, "gamma": gamma, "delta": delta, "C": C, "S": S}
        return OK

    def build_C_delta(self, m_y):  # :264-273
        xi0m = project_to_manifold(self.xi0)
        y0 = measure_system_state(xi0m)
        ye = output_group_action(self.X.inverse(), m_y)
        delta = output_coordinate_chart(ye, y0)
        C0 = output_matrix_C(xi0m)
        C = np.zeros((C0.shape[0], C0.shape[1] + 6))
        C[:, 6:] = C0
        return C, delta

    # ---- snapshot in the shared layout (include/eqvio.h) ----
    def get_snapshot(self):
        N = len(self.xi0.ids)
        hdr = np.zeros(49)
        hdr[0], hdr[1], hdr[2], hdr[3] = N, self.currentTime, float(self.initialisedFlag), self.accumulatedTime
        hdr[4:10] = self.inputBias
        hdr[10:13], hdr[13:16] = self.curOmega, self.curAccel
        hdr[16:19], hdr[19:22] = self.accOmega, self.accAccel
        hdr[22:26], hdr[26:29] = self.xi0.pose.R, self.xi0.pose.x
        hdr[29:32] = self.xi0.velocity
        hdr[32:36], hdr[36:39] = self.xi0.cameraOffset.R, self.xi0.cameraOffset.x
        hdr[39:43], hdr[43:46] = self.X.A.R, self.X.A.x
        hdr[46:49] = self.X.w
        L = np.zeros((N, 9))
        for i in range(N):
            L[i, 0] = self.xi0.ids[i]
            L[i, 1:4] = self.xi0.landmarks[i]
            L[i, 4:8] = self.X.Q[i].R
            L[i, 8] = self.X.Q[i].a
        return np.concatenate([hdr, L.reshape(-1), self.Sigma.reshape(-1, order="F")])

    def set_snapshot(self, d):
        d = np.asarray(d, dtype=float)
        N = int(d[0])
        self.currentTime, self.initialisedFlag, self.accumulatedTime = float(d[1]), bool(d[2]), float(d[3])
        self.inputBias = d[4:10].copy()
        self.curOmega, self.curAccel = d[10:13].copy(), d[13:16].copy()
        self.accOmega, self.accAccel = d[16:19].copy(), d[19:22].copy()
        self.xi0 = VIOState(SE3(d[22:26].copy(), d[26:29].copy()), d[29:32].copy(), np.zeros((N, 3)), [], SE3(d[32:36].copy(), d[36:39].copy()))
        self.X = VIOGroup(SE3(d[39:43].copy(), d[43:46].copy()), d[46:49].copy(), [], [])
        L = d[49 : 49 + 9 * N].reshape(N, 9)
        for i in range(N):
            self.xi0.ids.append(int(L[i, 0]))
            self.X.ids.append(int(L[i, 0]))
            self.xi0.landmarks[i] = L[i, 1:4]
            self.X.Q.append(SOT3(L[i, 4:8].copy(), float(L[i, 8])))
        n = SIGMA_BASE_SIZE + 3 * N
        self.Sigma = d[49 + 9 * N : 49 + 9 * N + n * n].reshape(n, n, order="F").copy()

    def clone(self):
        return copy.deepcopy(self)
