"""ctypes binding of the C ABI (include/eqvio.h) exported by csrc/libeqvio_b200.so.

There is no CPU fallback: if the CUDA library has not been built this module raises at import of the
library handle, and `eqvio_create` fails with EQVIO_ERR_NO_DEVICE on a box without a GPU."""
from __future__ import annotations

import ctypes as C
import os

from .settings import Settings

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libeqvio_b200.so")

OK, SKIPPED_DT, NOT_INITIALISED, EMPTY_MEASUREMENT = 0, 1, 2, 3
ERR_ARG, ERR_CUDA, ERR_NAN, ERR_SINGULAR_CHART, ERR_NOT_SPD, ERR_NO_DEVICE, ERR_UNSORTED = -1, -2, -3, -4, -5, -6, -7

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_h = C.c_void_p

# name -> (restype, argtypes); every symbol include/eqvio.h declares
SIGNATURES = {
    "eqvio_settings_default": (C.c_int, [C.POINTER(Settings)]),
    "eqvio_create": (C.c_int, [C.POINTER(Settings), C.c_int, C.POINTER(_h)]),
    "eqvio_destroy": (C.c_int, [_h]),
    "eqvio_reset": (C.c_int, [_h]),
    "eqvio_set_settings": (C.c_int, [_h, C.POINTER(Settings)]),
    "eqvio_set_auxiliary_data": (C.c_int, [_h, _dp, _dp, _dp]),
    "eqvio_initialise_from_imu": (C.c_int, [_h, _dp, _dp]),
    "eqvio_process_imu": (C.c_int, [_h, C.c_double, _dp, _dp]),
    "eqvio_process_vision": (C.c_int, [_h, C.c_double, C.c_int, _ip, _dp]),
    "eqvio_process_vision_dev": (C.c_int, [_h, C.c_double, C.c_int, _ip, C.c_void_p]),
    "eqvio_set_inertial_points": (C.c_int, [_h, C.c_int, _ip, _dp]),
    "eqvio_get_time": (C.c_int, [_h, _dp]),
    "eqvio_get_num_landmarks": (C.c_int, [_h, _ip]),
    "eqvio_get_state": (C.c_int, [_h, _dp, _dp, _dp, _ip, C.c_int, _ip, _dp]),
    "eqvio_get_pose_record": (C.c_int, [_h, _dp]),
    "eqvio_pose_record_dev": (C.c_int, [_h, C.POINTER(C.c_void_p)]),
    "eqvio_pose_publish": (C.c_int, [_h, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "eqvio_get_flags": (C.c_int, [_h, _ip, C.c_int]),
    "eqvio_get_covariance": (C.c_int, [_h, _dp, C.c_int]),
    "eqvio_get_bias": (C.c_int, [_h, _dp]),
    "eqvio_snapshot_size": (C.c_size_t, [C.c_int]),
    "eqvio_get_snapshot": (C.c_int, [_h, _dp, C.c_size_t]),
    "eqvio_set_snapshot": (C.c_int, [_h, _dp, C.c_size_t]),
    "eqvio_build_FB": (C.c_int, [_h, C.c_double, _dp, _dp, _dp]),
    "eqvio_riccati_propagate": (C.c_int, [_h, C.c_double, _dp]),
    "eqvio_build_C_delta": (C.c_int, [_h, _dp, _dp, _dp]),
    "eqvio_gain_update": (C.c_int, [_h, _dp, _dp, _dp]),
    "eqvio_bundle_lift": (C.c_int, [_h, _dp, _dp]),
    "eqvio_dgemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _dp, C.c_int, _dp, C.c_int, C.c_double, _dp, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "eqvio_dgemm_pair": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_double, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "eqvio_dgemm_ozaki": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "eqvio_schur_inverse": (C.c_int, [_h, C.c_int, _dp, C.c_int, _dp, C.c_int]),
    "eqvio_getrf_block": (C.c_int, [C.c_int, C.c_int, _dp, C.c_int, _dp, _dp, _dp, C.c_int, C.POINTER(C.c_float)]),
    "eqvio_synchronize": (C.c_int, [_h]),
    "eqvio_launch_count": (C.c_int, [_h, C.POINTER(C.c_longlong), C.c_int]),
    "eqvio_set_graphs": (C.c_int, [_h, C.c_int]),
    "eqvio_graph_stats": (C.c_int, [_h, C.POINTER(C.c_longlong), _ip]),
    "eqvio_profile_enable": (C.c_int, [_h, C.c_int]),
    "eqvio_profile_read": (C.c_int, [_h, C.POINTER(C.c_longlong), _dp, _dp, C.c_int]),
    "eqvio_profile_read_class": (C.c_int, [_h, C.c_int, C.POINTER(C.c_longlong), _dp, _dp, C.c_int]),
    "eqvio_profile_timeline": (C.c_int, [_h, _dp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "eqvio_riccati_arith": (C.c_int, [_h, _ip]),
    "eqvio_oz_stamps": (C.c_int, [_h, C.POINTER(C.c_longlong), C.c_size_t, C.POINTER(C.c_size_t)]),
    "eqvio_stream": (C.c_int, [_h, C.POINTER(C.c_void_p)]),
    "eqvio_status_string": (C.c_char_p, [C.c_int]),
    "eqvio_version": (C.c_char_p, []),
}

_lib = None


class EqvioError(RuntimeError):
    def __init__(self, status: int, where: str):
        msg = lib().eqvio_status_string(status).decode()
        super().__init__(f"{where}: status {status} ({msg})")
        self.status = status


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m eqf_vio_b200.build` (nvcc, sm_100a). "
                "The B200 filter path has no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int, where: str) -> int:
    if status < 0:
        raise EqvioError(status, where)
    return status
