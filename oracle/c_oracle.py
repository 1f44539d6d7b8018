"""ctypes binding of the C oracle (oracle/libeqvio_oracle.so).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from eqf_vio_b200.settings import Settings

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libeqvio_oracle.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "eqvio_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


def _p(a):
    return a.ctypes.data_as(_dp)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.eqo_create.restype = C.c_void_p
        L.eqo_create.argtypes = [C.POINTER(Settings)]
        L.eqo_destroy.argtypes = [C.c_void_p]
        L.eqo_process_imu.argtypes = [C.c_void_p, C.c_double, _dp, _dp]
        L.eqo_process_vision.argtypes = [C.c_void_p, C.c_double, C.c_int, _ip, _dp]
        L.eqo_set_inertial_points.argtypes = [C.c_void_p, C.c_int, _ip, _dp]
        L.eqo_get_time.restype = C.c_double
        L.eqo_get_time.argtypes = [C.c_void_p]
        L.eqo_get_num_landmarks.argtypes = [C.c_void_p]
        L.eqo_get_state.argtypes = [C.c_void_p, _dp, _dp, _dp, _ip, C.c_int, _ip, _dp]
        L.eqo_get_covariance.argtypes = [C.c_void_p, _dp, C.c_int]
        L.eqo_get_bias.argtypes = [C.c_void_p, _dp]
        L.eqo_snapshot_size.restype = C.c_size_t
        L.eqo_snapshot_size.argtypes = [C.c_int]
        L.eqo_get_snapshot.argtypes = [C.c_void_p, _dp, C.c_size_t]
        L.eqo_set_snapshot.argtypes = [C.c_void_p, _dp, C.c_size_t]
        L.eqo_state_matrix_A.argtypes = [C.c_void_p, _dp, _dp]
        L.eqo_input_matrix_B.argtypes = [C.c_void_p, _dp]
        L.eqo_output_matrix_C.argtypes = [C.c_void_p, _dp]
        L.eqo_build_FB.argtypes = [C.c_void_p, C.c_double, _dp, _dp, _dp]
        L.eqo_riccati_propagate.argtypes = [C.c_void_p, C.c_double, _dp]
        L.eqo_build_C_delta.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.eqo_gain_update.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.eqo_bundle_lift.argtypes = [C.c_void_p, _dp, _dp]
        L.eqo_lift_innovation_wls.argtypes = [C.c_void_p, _dp, _dp]
        L.eqo_lift_innovation.argtypes = [C.c_void_p, _dp, _dp]
        L.eqo_stereo_sphere_chart.argtypes = [_dp, _dp, _dp]
        L.eqo_stereo_sphere_chart_inv.argtypes = [_dp, _dp, _dp]
        L.eqo_dgemm.restype = None
        L.eqo_dgemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _dp, C.c_int, _dp, C.c_int, C.c_double, _dp, C.c_int]
        L.eqo_inverse.argtypes = [C.c_int, _dp, C.c_int, _dp, C.c_int]
        _lib = L
    return _lib


def _vec(a, n=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    if n is not None:
        assert a.size == n, (a.size, n)
    return a


class COracleFilter:
    """The C restatement behind the reference's method names."""

    def __init__(self, settings: Settings):
        self._L = lib()
        self._s = settings.copy()
        self._h = self._L.eqo_create(C.byref(self._s))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.eqo_destroy(self._h)
            self._h = None

    @property
    def N(self):
        return self._L.eqo_get_num_landmarks(self._h)

    @property
    def n(self):
        return 11 + 3 * self.N

    def processIMUData(self, stamp, omega, accel):
        return self._L.eqo_process_imu(self._h, float(stamp), _p(_vec(omega, 3)), _p(_vec(accel, 3)))

    def processVisionData(self, stamp, ids, bearings):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        y = _vec(bearings, 3 * len(ids))
        return self._L.eqo_process_vision(self._h, float(stamp), len(ids), ids.ctypes.data_as(_ip), _p(y))

    def setInertialPoints(self, ids, points):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        return self._L.eqo_set_inertial_points(self._h, len(ids), ids.ctypes.data_as(_ip), _p(_vec(points, 3 * len(ids))))

    def getTime(self):
        return self._L.eqo_get_time(self._h)

    def stateEstimate(self):
        N = self.N
        pose, vel, cam = np.zeros(7), np.zeros(3), np.zeros(7)
        ids = np.zeros(max(N, 1), dtype=np.int32)
        lm = np.zeros(3 * max(N, 1))
        n = C.c_int(0)
        self._L.eqo_get_state(self._h, _p(pose), _p(vel), _p(cam), C.byref(n), N, ids.ctypes.data_as(_ip), _p(lm))
        return {"pose": pose, "velocity": vel, "cameraOffset": cam, "ids": ids[:N].copy(), "landmarks": lm[: 3 * N].reshape(N, 3).copy()}

    def stateCovariance(self):
        n = self.n
        S = np.zeros((n, n), order="F")
        self._L.eqo_get_covariance(self._h, _p(S), n)
        return S

    def bias(self):
        b = np.zeros(6)
        self._L.eqo_get_bias(self._h, _p(b))
        return b

    def get_snapshot(self):
        d = np.zeros(self._L.eqo_snapshot_size(self.N))
        st = self._L.eqo_get_snapshot(self._h, _p(d), d.size)
        assert st == 0
        return d

    def set_snapshot(self, d):
        d = _vec(d)
        st = self._L.eqo_set_snapshot(self._h, _p(d), d.size)
        assert st == 0, st

    # pieces
    def state_matrix_A(self, omega):
        p = 5 + 3 * self.N
        A = np.zeros((p, p), order="F")
        st = self._L.eqo_state_matrix_A(self._h, _p(_vec(omega, 3)), _p(A))
        assert st == 0, st
        return A

    def input_matrix_B(self):
        p = 5 + 3 * self.N
        B = np.zeros((p, 6), order="F")
        assert self._L.eqo_input_matrix_B(self._h, _p(B)) == 0
        return B

    def output_matrix_C(self):
        N = self.N
        Cm = np.zeros((2 * N, 5 + 3 * N), order="F")
        assert self._L.eqo_output_matrix_C(self._h, _p(Cm)) == 0
        return Cm

    def build_FB(self, T, omega):
        n = self.n
        F = np.zeros((n, n), order="F")
        Bb = np.zeros((n, 6), order="F")
        assert self._L.eqo_build_FB(self._h, float(T), _p(_vec(omega, 3)), _p(F), _p(Bb)) == 0
        return F, Bb

    def riccati_propagate(self, T, omega):
        return self._L.eqo_riccati_propagate(self._h, float(T), _p(_vec(omega, 3)))

    def build_C_delta(self, bearings):
        N, n = self.N, self.n
        Cm = np.zeros((2 * N, n), order="F")
        d = np.zeros(2 * N)
        assert self._L.eqo_build_C_delta(self._h, _p(_vec(bearings, 3 * N)), _p(Cm), _p(d)) == 0
        return Cm, d

    def gain_update(self, bearings):
        N, n = self.N, self.n
        K = np.zeros((n, 2 * N), order="F")
        g = np.zeros(n)
        st = self._L.eqo_gain_update(self._h, _p(_vec(bearings, 3 * N)), _p(K), _p(g))
        assert st == 0, st
        return K, g

    def bundle_lift(self, gamma_eqf):
        N = self.N
        G = np.zeros(9 + 3 * N)
        st = self._L.eqo_bundle_lift(self._h, _p(_vec(gamma_eqf, 5 + 3 * N)), _p(G))
        assert st == 0, st
        return G

    def lift_innovation(self, gamma_eqf, wls=False):
        N = self.N
        a = np.zeros(9 + 4 * N)
        fn = self._L.eqo_lift_innovation_wls if wls else self._L.eqo_lift_innovation
        st = fn(self._h, _p(_vec(gamma_eqf, 5 + 3 * N)), _p(a))
        assert st == 0, st
        return a


def dgemm(A, B, transA=False, transB=False):
    A = np.asfortranarray(A, dtype=np.float64)
    B = np.asfortranarray(B, dtype=np.float64)
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    N = B.shape[0] if transB else B.shape[1]
    Cm = np.zeros((M, N), order="F")
    lib().eqo_dgemm(int(transA), int(transB), M, N, K, 1.0, _p(A), A.shape[0], _p(B), B.shape[0], 0.0, _p(Cm), M)
    return Cm


def inverse(A):
    A = np.asfortranarray(A, dtype=np.float64)
    n = A.shape[0]
    Ai = np.zeros((n, n), order="F")
    st = lib().eqo_inverse(n, _p(A), n, _p(Ai), n)
    assert st == 0
    return Ai
