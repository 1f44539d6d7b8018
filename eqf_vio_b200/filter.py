"""Host-side mirror of the reference's `class VIOFilter` (eqf_vio/include/eqf_vio/VIOFilter.h:41-88)
over the C ABI: same method names, argument meaning and silent-skip behaviour; every call lands in
csrc/libeqvio_b200.so (hand-written sm_100a kernels).  Status codes the reference expresses by silently
returning are returned as small positive ints (abi.SKIPPED_DT ...); errors raise EqvioError."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import abi
from .settings import Settings


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _vec(a, n=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} values, got {a.size}")
    return a


@dataclass
class VIOStateEstimate:
    """VIOState (eqf_vio/include/eqf_vio/VIOState.h:51-60) as plain arrays; poses are x y z qw qx qy qz."""

    pose: np.ndarray
    velocity: np.ndarray
    cameraOffset: np.ndarray
    ids: np.ndarray
    bodyLandmarks: np.ndarray


class VIOFilter:
    def __init__(self, settings: Settings, device: int = 0):
        self._L = abi.lib()
        self.settings = settings.copy()
        self._h = C.c_void_p()
        abi.check(self._L.eqvio_create(C.byref(self.settings), int(device), C.byref(self._h)), "eqvio_create")
        self.device = device
        # the IMU call runs 200 times per second of data: its six doubles go through one preallocated ctypes buffer
        # (numpy -> ctypes pointer conversion costs ~4 us per array, more than the C call itself)
        self._imu6 = (C.c_double * 6)()
        self._imu_om = C.cast(self._imu6, C.POINTER(C.c_double))
        self._imu_ac = C.cast(C.byref(self._imu6, 24), C.POINTER(C.c_double))
        self._process_imu = self._L.eqvio_process_imu

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.eqvio_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs (VIOFilter.h:74-81) ----
    def reset(self):
        abi.check(self._L.eqvio_reset(self._h), "eqvio_reset")

    def setAuxiliaryData(self, attitude_wxyz, position, cameraOffset):
        """VIOFilter::setAuxiliaryData (VIOFilter.cpp:75-82); cameraOffset is x y z qw qx qy qz."""
        abi.check(self._L.eqvio_set_auxiliary_data(self._h, _p(_vec(attitude_wxyz, 4)), _p(_vec(position, 3)), _p(_vec(cameraOffset, 7))), "eqvio_set_auxiliary_data")

    def initialiseFromIMUData(self, omega, accel):
        """VIOFilter::initialiseFromIMUData (VIOFilter.cpp:133-144): the sample is used as given."""
        abi.check(self._L.eqvio_initialise_from_imu(self._h, _p(_vec(omega, 3)), _p(_vec(accel, 3))), "eqvio_initialise_from_imu")

    def processIMUData(self, stamp, omega, accel) -> int:
        o = omega.tolist() if hasattr(omega, "tolist") else list(omega)
        a = accel.tolist() if hasattr(accel, "tolist") else list(accel)
        if len(o) != 3 or len(a) != 3:
            raise ValueError("omega and accel must have 3 values each")
        self._imu6[:] = o + a
        st = self._process_imu(self._h, float(stamp), self._imu_om, self._imu_ac)
        return st if st >= 0 else abi.check(st, "eqvio_process_imu")

    def processVisionData(self, stamp, ids, bearings) -> int:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        y = _vec(bearings, 3 * ids.size)
        return abi.check(
            self._L.eqvio_process_vision(self._h, float(stamp), int(ids.size), ids.ctypes.data_as(C.POINTER(C.c_int)), _p(y)),
            "eqvio_process_vision",
        )

    def processVisionDataDevice(self, stamp, ids, bearings_dev_ptr: int) -> int:
        """Bearings already resident in device memory (3n doubles at `bearings_dev_ptr`)."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        return abi.check(
            self._L.eqvio_process_vision_dev(self._h, float(stamp), int(ids.size), ids.ctypes.data_as(C.POINTER(C.c_int)), C.c_void_p(bearings_dev_ptr)),
            "eqvio_process_vision_dev",
        )

    def setInertialPoints(self, ids, points):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        p = _vec(points, 3 * ids.size)
        abi.check(self._L.eqvio_set_inertial_points(self._h, int(ids.size), ids.ctypes.data_as(C.POINTER(C.c_int)), _p(p)), "eqvio_set_inertial_points")

    # ---- outputs (VIOFilter.h:84-86) ----
    def getTime(self) -> float:
        t = C.c_double()
        abi.check(self._L.eqvio_get_time(self._h, C.byref(t)), "eqvio_get_time")
        return t.value

    @property
    def numLandmarks(self) -> int:
        n = C.c_int()
        abi.check(self._L.eqvio_get_num_landmarks(self._h, C.byref(n)), "eqvio_get_num_landmarks")
        return n.value

    def stateEstimate(self) -> VIOStateEstimate:
        N = self.numLandmarks
        pose, vel, cam = np.zeros(7), np.zeros(3), np.zeros(7)
        ids = np.zeros(max(N, 1), dtype=np.int32)
        lm = np.zeros(3 * max(N, 1))
        n = C.c_int()
        abi.check(
            self._L.eqvio_get_state(self._h, _p(pose), _p(vel), _p(cam), C.byref(n), N, ids.ctypes.data_as(C.POINTER(C.c_int)), _p(lm)),
            "eqvio_get_state",
        )
        return VIOStateEstimate(pose, vel, cam, ids[:N].copy(), lm[: 3 * N].reshape(N, 3).copy())

    def poseRecord(self) -> np.ndarray:
        rec = np.zeros(8)
        abi.check(self._L.eqvio_get_pose_record(self._h, _p(rec)), "eqvio_get_pose_record")
        return rec

    def poseRecordDevicePtr(self) -> int:
        p = C.c_void_p()
        abi.check(self._L.eqvio_pose_record_dev(self._h, C.byref(p)), "eqvio_pose_record_dev")
        return p.value

    def posePublish(self):
        """(gather stream handle, device pointer of the published 8-double pose record): see eqvio_pose_publish."""
        st, p = C.c_void_p(), C.c_void_p()
        abi.check(self._L.eqvio_pose_publish(self._h, C.byref(st), C.byref(p)), "eqvio_pose_publish")
        return st.value or 0, p.value

    def deviceFlags(self, clear=False) -> int:
        """Sticky device-side error bits as a status code (abi.OK / ERR_SINGULAR_CHART / ERR_NOT_SPD / ERR_NAN)."""
        v = C.c_int()
        abi.check(self._L.eqvio_get_flags(self._h, C.byref(v), int(clear)), "eqvio_get_flags")
        return v.value

    def stateCovariance(self) -> np.ndarray:
        n = 11 + 3 * self.numLandmarks
        S = np.zeros((n, n), order="F")
        abi.check(self._L.eqvio_get_covariance(self._h, _p(S), n), "eqvio_get_covariance")
        return S

    def inputBias(self) -> np.ndarray:
        b = np.zeros(6)
        abi.check(self._L.eqvio_get_bias(self._h, _p(b)), "eqvio_get_bias")
        return b

    def schur_inverse(self, S) -> np.ndarray:
        """S^-1 by the update's own blocked Schur elimination (kernel-level entry point, eqvio_schur_inverse)."""
        S = np.asfortranarray(S, dtype=np.float64)
        m = S.shape[0]
        out = np.zeros((m, m), order="F")
        abi.check(self._L.eqvio_schur_inverse(self._h, m, _p(S), m, _p(out), m), "eqvio_schur_inverse")
        return out

    # ---- snapshot / restore ----
    def get_snapshot(self) -> np.ndarray:
        d = np.zeros(self._L.eqvio_snapshot_size(self.numLandmarks))
        abi.check(self._L.eqvio_get_snapshot(self._h, _p(d), d.size), "eqvio_get_snapshot")
        return d

    def set_snapshot(self, d):
        d = _vec(d)
        abi.check(self._L.eqvio_set_snapshot(self._h, _p(d), d.size), "eqvio_set_snapshot")

    # ---- kernel-level entry points ----
    def build_FB(self, T, omega):
        n = 11 + 3 * self.numLandmarks
        F = np.zeros((n, n), order="F")
        Bb = np.zeros((n, 6), order="F")
        abi.check(self._L.eqvio_build_FB(self._h, float(T), _p(_vec(omega, 3)), _p(F), _p(Bb)), "eqvio_build_FB")
        return F, Bb

    def riccati_propagate(self, T, omega):
        abi.check(self._L.eqvio_riccati_propagate(self._h, float(T), _p(_vec(omega, 3))), "eqvio_riccati_propagate")

    def build_C_delta(self, bearings):
        N = self.numLandmarks
        Cm = np.zeros((2 * N, 11 + 3 * N), order="F")
        d = np.zeros(2 * N)
        abi.check(self._L.eqvio_build_C_delta(self._h, _p(_vec(bearings, 3 * N)), _p(Cm), _p(d)), "eqvio_build_C_delta")
        return Cm, d

    def gain_update(self, bearings):
        N = self.numLandmarks
        n = 11 + 3 * N
        K = np.zeros((n, 2 * N), order="F")
        g = np.zeros(n)
        abi.check(self._L.eqvio_gain_update(self._h, _p(_vec(bearings, 3 * N)), _p(K), _p(g)), "eqvio_gain_update")
        return K, g

    def bundle_lift(self, gamma_eqf):
        N = self.numLandmarks
        G = np.zeros(9 + 3 * N)
        abi.check(self._L.eqvio_bundle_lift(self._h, _p(_vec(gamma_eqf, 5 + 3 * N)), _p(G)), "eqvio_bundle_lift")
        return G

    # ---- instrumentation ----
    def synchronize(self):
        abi.check(self._L.eqvio_synchronize(self._h), "eqvio_synchronize")

    def launch_count(self, reset=False) -> int:
        c = C.c_longlong()
        abi.check(self._L.eqvio_launch_count(self._h, C.byref(c), int(reset)), "eqvio_launch_count")
        return c.value

    def set_graphs(self, on=True):
        abi.check(self._L.eqvio_set_graphs(self._h, int(on)), "eqvio_set_graphs")

    def graph_stats(self):
        """(graph replays so far, instantiated graphs held)."""
        n, c = C.c_longlong(), C.c_int()
        abi.check(self._L.eqvio_graph_stats(self._h, C.byref(n), C.byref(c)), "eqvio_graph_stats")
        return n.value, c.value

    def profile_enable(self, on=True):
        abi.check(self._L.eqvio_profile_enable(self._h, int(on)), "eqvio_profile_enable")

    def profile_read(self, reset=True):
        n = C.c_longlong()
        ms, fl = C.c_double(), C.c_double()
        abi.check(self._L.eqvio_profile_read(self._h, C.byref(n), C.byref(ms), C.byref(fl), int(reset)), "eqvio_profile_read")
        return n.value, ms.value, fl.value

    PROFILE_CLASSES = ("riccati_gemm", "update_gemm", "schur_gemm", "schur_diag_lu", "small_kernels", "riccati_i8_gemm")
    PROFILE_LANES = ("main", "side", "lift", "main_helper", "lift_helper", "other")

    def profile_timeline(self) -> np.ndarray:
        """(k, 5) array: class, stream lane, start ms, end ms, flops of every bracketed launch since profile_enable."""
        cnt = C.c_size_t()
        abi.check(self._L.eqvio_profile_timeline(self._h, None, 0, C.byref(cnt)), "eqvio_profile_timeline")
        out = np.zeros((cnt.value, 5))
        if cnt.value:
            abi.check(self._L.eqvio_profile_timeline(self._h, _p(out), cnt.value, C.byref(cnt)), "eqvio_profile_timeline")
        return out

    def profile_read_classes(self, reset=True) -> dict:
        out = {}
        for cls, name in enumerate(self.PROFILE_CLASSES):
            n = C.c_longlong()
            ms, fl = C.c_double(), C.c_double()
            abi.check(self._L.eqvio_profile_read_class(self._h, cls, C.byref(n), C.byref(ms), C.byref(fl), int(reset)), "eqvio_profile_read_class")
            out[name] = {"launches": n.value, "ms": ms.value, "flops": fl.value}
        return out

    def riccati_int8_slices(self) -> int:
        """0: the Riccati products run on fp64 DMMA at the current N; S > 0: on the int8 tensor cores with S slices."""
        s = C.c_int()
        abi.check(self._L.eqvio_riccati_arith(self._h, C.byref(s)), "eqvio_riccati_arith")
        return s.value

    def stream_ptr(self) -> int:
        p = C.c_void_p()
        abi.check(self._L.eqvio_stream(self._h, C.byref(p)), "eqvio_stream")
        return p.value or 0


def getrf_block(A, device=0, reps=1):
    """Unpivoted LU of an nb x nb block (nb <= 64) and the 64 x 64 identity-padded triangular inverses on the
    library's diagonal-block kernel.  Returns (LU, Linv, Uinv, microseconds per launch)."""
    L = abi.lib()
    A = np.asfortranarray(A, dtype=np.float64)
    nb = A.shape[0]
    LU = np.zeros((nb, nb), order="F")
    Li, Ui = np.zeros((64, 64), order="F"), np.zeros((64, 64), order="F")
    us = C.c_float()
    abi.check(L.eqvio_getrf_block(int(device), nb, _p(A), nb, _p(LU), _p(Li), _p(Ui), int(reps), C.byref(us)), "eqvio_getrf_block")
    return LU, Li, Ui, us.value


def dgemm_pair(A1, B1, B2, transB2=True, alpha2=1.0, device=0, reps=1):
    """W = A1 @ B1, D = alpha2 * W @ op(B2) as ONE launch of the library's pair kernel (host arrays).  Returns (W, D, ms)."""
    L = abi.lib()
    A1 = np.asfortranarray(A1, dtype=np.float64)
    B1 = np.asfortranarray(B1, dtype=np.float64)
    B2 = np.asfortranarray(B2, dtype=np.float64)
    M, K1 = A1.shape
    N1 = B1.shape[1]
    assert B1.shape[0] == K1
    N2 = B2.shape[0] if transB2 else B2.shape[1]
    assert (B2.shape[1] if transB2 else B2.shape[0]) == N1
    W = np.zeros((M, N1), order="F")
    D = np.zeros((M, N2), order="F")
    ms = C.c_float()
    abi.check(
        L.eqvio_dgemm_pair(int(device), M, N1, K1, _p(A1), M, _p(B1), K1, int(transB2), N2, float(alpha2), _p(B2), B2.shape[0], _p(W), M, _p(D), M, int(reps), C.byref(ms)),
        "eqvio_dgemm_pair",
    )
    return W, D, ms.value


def dgemm(A, B, transB=False, alpha=1.0, beta=0.0, Cin=None, device=0, reps=1):
    """C = alpha * A @ op(B) + beta * Cin on the library's DMMA kernel (host arrays).  Returns (C, ms)."""
    L = abi.lib()
    A = np.asfortranarray(A, dtype=np.float64)
    B = np.asfortranarray(B, dtype=np.float64)
    M, K = A.shape
    N = B.shape[0] if transB else B.shape[1]
    assert (B.shape[1] if transB else B.shape[0]) == K
    Cm = np.zeros((M, N), order="F") if Cin is None else np.asfortranarray(Cin, dtype=np.float64).copy(order="F")
    ms = C.c_float()
    abi.check(
        L.eqvio_dgemm(int(device), int(transB), M, N, K, float(alpha), _p(A), max(A.shape[0], 1), _p(B), max(B.shape[0], 1), float(beta), _p(Cm), max(M, 1), int(reps), C.byref(ms)),
        "eqvio_dgemm",
    )
    return Cm, ms.value


def dgemm_ozaki(A, B, transB=False, slices=8, device=0, reps=1):
    """C = A @ op(B) assembled from int8 tensor-core products (tcgen05, Ozaki splitting; eqvio_dgemm_ozaki).
    Returns (C, ms per whole call, ms of the tcgen05 kernel alone)."""
    L = abi.lib()
    A = np.asfortranarray(A, dtype=np.float64)
    B = np.asfortranarray(B, dtype=np.float64)
    M, K = A.shape
    N = B.shape[0] if transB else B.shape[1]
    assert (B.shape[1] if transB else B.shape[0]) == K
    Cm = np.zeros((M, N), order="F")
    t_all, t_gemm = C.c_float(), C.c_float()
    abi.check(
        L.eqvio_dgemm_ozaki(int(device), int(transB), M, N, K, _p(A), M, _p(B), B.shape[0], _p(Cm), M, int(slices), int(reps), C.byref(t_all), C.byref(t_gemm)),
        "eqvio_dgemm_ozaki",
    )
    return Cm, t_all.value, t_gemm.value
