"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the path
at two sizes, landmark churn included."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eqf_vio_b200.filter import VIOFilter, dgemm
from eqf_vio_b200.settings import template_settings, conditioned_settings
from eqf_vio_b200.synthetic import period_sequence

rng = np.random.default_rng(0)
for N, s in ((9, template_settings()), (70, conditioned_settings())):
    seq = period_sequence(N, 3, camera_offset=tuple(s.cameraOffset))
    f = VIOFilter(s)
    for kind, i in seq.events():
        if kind == "imu":
            f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else:
            sel = np.sort(rng.choice(N, size=max(3, N - 2), replace=False)) if i > 0 else np.arange(N - 1)
            f.processVisionData(seq.vision_stamps[i], seq.ids[sel], seq.bearings[i][sel])
    e = f.stateEstimate()
    S = f.stateCovariance()
    print("N", N, "landmarks", f.numLandmarks, "finite", np.isfinite(S).all(), "launches", f.launch_count())
# fixed landmark count: the update and the Riccati step run as replayed CUDA graphs from the third frame on
s = conditioned_settings(outlierThreshold=1e9)
seq = period_sequence(40, 5, camera_offset=tuple(s.cameraOffset))
f = VIOFilter(s)
for kind, i in seq.events():
    if kind == "imu":
        f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
    else:
        f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
print("N 40 fixed: graph replays", f.graph_stats(), "finite", np.isfinite(f.stateCovariance()).all())
# the narrow lift form of capacities > 256 at a size a sanitizer can afford (forced): k_fill_u64 + k_lift_rsolve (cp.async tile ring,
# self-validating hand-over of R between its CTAs) behind the elimination, 5 blocks
os.environ["EQVIO_LIFT_NARROW"] = "1"
s = conditioned_settings(outlierThreshold=1e9)
seq = period_sequence(90, 3, camera_offset=tuple(s.cameraOffset))
f = VIOFilter(s)
for kind, i in seq.events():
    if kind == "imu":
        f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
    else:
        f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
print("N 90 narrow lift: finite", np.isfinite(f.stateCovariance()).all(), "device flags", f.deviceFlags())
f.close()
del os.environ["EQVIO_LIFT_NARROW"]
# the pair kernel (ticket + row-block counters between CTAs) at a ragged multi-row-block size, both second-product layouts
from eqf_vio_b200.filter import dgemm_pair
for tB2 in (True, False):
    A1 = rng.standard_normal((75, 41)); B1 = rng.standard_normal((41, 70)); B2 = rng.standard_normal((50, 70) if tB2 else (70, 50))
    W, D, _ = dgemm_pair(A1, B1, B2, transB2=tB2)
    print("gemm pair", tB2, np.abs(D - (A1 @ B1) @ (B2.T if tB2 else B2)).max())
from eqf_vio_b200.filter import getrf_block
Bm = rng.standard_normal((48, 48))
LU, Li, Ui, _ = getrf_block(Bm @ Bm.T + 48 * np.eye(48))
print("chain block", np.isfinite(LU).all())
A = rng.standard_normal((45, 37)); B = rng.standard_normal((37, 50))
C, _ = dgemm(A, B)
print("gemm", np.abs(C - A @ B).max())
# the int8 tensor-core Riccati path at a size a sanitizer can afford (forced on: 4 tiles, border with landmark rows): fused kernel
# (tickets, tile-row barriers, emission), structural split of F, generic split kernels behind each update, and the unfused product
if os.environ.get("EQVIO_SANITIZE_INT8", "1") != "0":
    os.environ["EQVIO_OZAKI_MIN_TILES"] = "4"
    os.environ["EQVIO_OZ_SCT"] = "1"      # ... and the vision update's six products on the int8 path as well (structural splits of C,
    os.environ["EQVIO_OZ_UPDATE"] = "1"   # generic splits, k_oz_gemm, thin DMMA products on the helper stream)
    s = conditioned_settings(outlierThreshold=1e9)
    seq = period_sequence(90, 3, camera_offset=tuple(s.cameraOffset))
    f = VIOFilter(s)
    for kind, i in seq.events():
        if kind == "imu":
            f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else:
            f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
    print("N 90 int8 slices", f.riccati_int8_slices(), "finite", np.isfinite(f.stateCovariance()).all(), "graph replays", f.graph_stats())
    from eqf_vio_b200.filter import dgemm_ozaki
    A = rng.standard_normal((139, 150)); B = rng.standard_normal((150, 260))
    C, _, _ = dgemm_ozaki(A, B, slices=8)
    print("ozaki gemm", np.abs(C - A @ B).max())
