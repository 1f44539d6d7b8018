"""ctypes binding of oracle/_ref/libeqvio_ref.so — the reference's own sources compiled against the
Eigen stand-in (oracle/refshim).  TEST INFRASTRUCTURE ONLY; exists only where /root/reference is mounted
(this container) or where the prebuilt .so travelled with the repo snapshot."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from eqf_vio_b200.settings import Settings

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libeqvio_ref.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def available() -> bool:
    return os.path.exists(LIB)


def build() -> bool:
    if os.path.exists("/root/reference/eqf_vio/src/VIOFilter.cpp"):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "refshim"), "-s"])
    return available()


def _p(a):
    return a.ctypes.data_as(_dp)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.POINTER(Settings)]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_process_imu.argtypes = [C.c_void_p, C.c_double, _dp, _dp]
        L.ref_process_vision.argtypes = [C.c_void_p, C.c_double, C.c_int, _ip, _dp]
        L.ref_num_landmarks.argtypes = [C.c_void_p]
        L.ref_get_time.restype = C.c_double
        L.ref_get_time.argtypes = [C.c_void_p]
        L.ref_snapshot_size.restype = C.c_size_t
        L.ref_snapshot_size.argtypes = [C.c_int]
        L.ref_get_snapshot.argtypes = [C.c_void_p, _dp]
        L.ref_set_snapshot.argtypes = [C.c_void_p, _dp]
        L.ref_state_matrix_A.argtypes = [C.c_void_p, _dp, _dp]
        L.ref_input_matrix_B.argtypes = [C.c_void_p, _dp]
        L.ref_output_matrix_C.argtypes = [C.c_void_p, _dp]
        L.ref_bundle_lift.argtypes = [C.c_void_p, _dp, _dp]
        L.ref_delta.argtypes = [C.c_void_p, _dp, _dp]
        _lib = L
    return _lib


def _vec(a, n=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    if n is not None:
        assert a.size == n
    return a


class ReferenceFilter:
    """The reference's `VIOFilter`, method names unchanged."""

    def __init__(self, settings: Settings):
        self._L = lib()
        self._s = settings.copy()
        self._h = self._L.ref_create(C.byref(self._s))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_destroy(self._h)
            self._h = None

    @property
    def N(self):
        return self._L.ref_num_landmarks(self._h)

    def processIMUData(self, stamp, omega, accel):
        return self._L.ref_process_imu(self._h, float(stamp), _p(_vec(omega, 3)), _p(_vec(accel, 3)))

    def processVisionData(self, stamp, ids, bearings):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        return self._L.ref_process_vision(self._h, float(stamp), len(ids), ids.ctypes.data_as(_ip), _p(_vec(bearings, 3 * len(ids))))

    def getTime(self):
        return self._L.ref_get_time(self._h)

    def get_snapshot(self):
        d = np.zeros(self._L.ref_snapshot_size(self.N))
        self._L.ref_get_snapshot(self._h, _p(d))
        return d

    def set_snapshot(self, d):
        self._L.ref_set_snapshot(self._h, _p(_vec(d)))

    def state_matrix_A(self, omega):
        p = 5 + 3 * self.N
        A = np.zeros((p, p), order="F")
        self._L.ref_state_matrix_A(self._h, _p(_vec(omega, 3)), _p(A))
        return A

    def input_matrix_B(self):
        p = 5 + 3 * self.N
        B = np.zeros((p, 6), order="F")
        self._L.ref_input_matrix_B(self._h, _p(B))
        return B

    def output_matrix_C(self):
        N = self.N
        Cm = np.zeros((2 * N, 5 + 3 * N), order="F")
        self._L.ref_output_matrix_C(self._h, _p(Cm))
        return Cm

    def bundle_lift(self, gamma_eqf):
        N = self.N
        G = np.zeros(9 + 3 * N)
        self._L.ref_bundle_lift(self._h, _p(_vec(gamma_eqf, 5 + 3 * N)), _p(G))
        return G

    def delta(self, bearings):
        N = self.N
        d = np.zeros(2 * N)
        self._L.ref_delta(self._h, _p(_vec(bearings, 3 * N)), _p(d))
        return d
