"""eqf_vio_b200 — B200-native (sm_100a) EqF-VIO filter hot path behind the reference's VIOFilter API.

The compute path is the CUDA library `eqf_vio_b200/csrc/libeqvio_b200.so` reached through the C ABI in
include/eqvio.h; this package only holds the host-side mirror of the reference interface, the POD
settings, the synthetic-sequence generator and the build script.  There is no CPU fallback: importing
`eqf_vio_b200.filter` without the built library raises."""

__version__ = "0.1.0"
