"""GPU bring-up diagnostics (not a test): GEMM correctness/timing, per-kernel parity vs the C oracle,
short sequences.  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eqf_vio_b200 import abi
from eqf_vio_b200.filter import VIOFilter, dgemm
from eqf_vio_b200.settings import template_settings
from eqf_vio_b200.synthetic import period_sequence
from oracle.c_oracle import COracleFilter

rng = np.random.default_rng(0)
what = sys.argv[1:] or ["gemm", "pieces", "seq"]

def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)

if "gemm" in what:
    for cfg in (None, 0, 1, 2, 3):
        if cfg is None: os.environ.pop("EQVIO_GEMM_CONFIG", None)
        else: os.environ["EQVIO_GEMM_CONFIG"] = str(cfg)
        for (M, N, K) in [(26, 26, 26), (64, 64, 64), (203, 203, 203), (128, 203, 77), (10, 26, 26), (300, 150, 6), (779, 779, 779)]:
            for tB in (False, True):
                A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K) if tB else (K, N)); C0 = rng.standard_normal((M, N))
                C1, _ = dgemm(A, B, transB=tB, alpha=-1.0, beta=1.0, Cin=C0)
                ref = C0 - A @ (B.T if tB else B)
                e = rel(C1, ref)
                print(f"cfg={cfg} M={M} N={N} K={K} tB={int(tB)} rel={e:.2e}", "OK" if e < 1e-13 else "FAIL", flush=True)
    os.environ.pop("EQVIO_GEMM_CONFIG", None)

if "gemmperf" in what:
    for n in (203, 779, 1547, 3083):
        for cfg in (None, 0, 1, 2):
            if cfg is None: os.environ.pop("EQVIO_GEMM_CONFIG", None)
            else: os.environ["EQVIO_GEMM_CONFIG"] = str(cfg)
            A = rng.standard_normal((n, n)); B = rng.standard_normal((n, n))
            for tB in (False, True):
                C1, ms = dgemm(A, B, transB=tB, reps=20)
                print(f"n={n} cfg={cfg} tB={int(tB)} {ms:.4f} ms {2*n**3/ms/1e9:.2f} TFLOP/s rel={rel(C1, A @ (B.T if tB else B)):.1e}", flush=True)
    os.environ.pop("EQVIO_GEMM_CONFIG", None)

def run_pair(N, periods, s=None, check_every=True, verbose=True):
    s = s or template_settings(outlierThreshold=1e9)
    seq = period_sequence(N, periods, camera_offset=tuple(s.cameraOffset))
    fg = VIOFilter(s); fc = COracleFilter(s)
    worst = 0
    for kind, i in seq.events():
        if kind == "imu":
            r1 = fg.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7]); r2 = fc.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else:
            r1 = fg.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i]); r2 = fc.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
        assert r1 == r2, (kind, i, r1, r2)
        if check_every or kind == "vision":
            a, b = fg.get_snapshot(), fc.get_snapshot()
            hn = 49 + 9 * fc.N
            es = rel(a[hn:], b[hn:]); eh = np.abs(a[:hn] - b[:hn]).max()
            worst = max(worst, es)
            if verbose and (kind == "vision" or i % 5 == 0):
                print(f"N={N} {kind}{i} st={r1} Sigma rel {es:.2e} state max {eh:.2e}", flush=True)
    return worst

if "pieces" in what:
    N = 16
    s = template_settings(outlierThreshold=1e9)
    seq = period_sequence(N, 2, camera_offset=tuple(s.cameraOffset))
    fc = COracleFilter(s)
    ev = list(seq.events())
    for kind, i in ev:
        if kind == "vision" and i == 1: break
        if kind == "imu": fc.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else: fc.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
    snap = fc.get_snapshot()
    fg = VIOFilter(s); fg.set_snapshot(snap)
    print("snapshot roundtrip", np.abs(fg.get_snapshot() - snap).max())
    om = np.array([0.1, -0.2, 0.05]); T = 0.005
    F1, B1 = fg.build_FB(T, om); F2, B2 = fc.build_FB(T, om)
    print("F", np.abs(F1 - F2).max(), "Bb", np.abs(B1 - B2).max())
    print("state untouched", np.abs(fg.get_snapshot() - snap).max())
    fg.riccati_propagate(T, om); fc.riccati_propagate(T, om)
    print("riccati Sigma rel", rel(fg.stateCovariance(), fc.stateCovariance()))
    fg.set_snapshot(snap); fc.set_snapshot(snap)
    y = seq.bearings[1]
    C1, d1 = fg.build_C_delta(y); C2, d2 = fc.build_C_delta(y)
    print("C", np.abs(C1 - C2).max(), "delta", np.abs(d1 - d2).max())
    g0 = rng.standard_normal(5 + 3 * N) * 1e-2
    G1 = fg.bundle_lift(g0); G2 = fc.bundle_lift(g0)
    print("bundleLift", np.abs(G1 - G2).max(), G2[:6])
    K1, g1 = fg.gain_update(y); K2, g2 = fc.gain_update(y)
    print("K", np.abs(K1 - K2).max() / np.abs(K2).max(), "gamma", np.abs(g1 - g2).max(), "Sigma rel", rel(fg.stateCovariance(), fc.stateCovariance()))

if "seq" in what:
    for N, P in ((5, 3), (16, 3), (64, 3)):
        t0 = time.time()
        w = run_pair(N, P)
        print(f"N={N} worst Sigma rel {w:.2e}  ({time.time()-t0:.1f}s)", flush=True)
    # bookkeeping: default outlier threshold, ids that come and go
    s = template_settings()
    seq = period_sequence(12, 6, camera_offset=tuple(s.cameraOffset))
    fg = VIOFilter(s); fc = COracleFilter(s)
    for kind, i in seq.events():
        if kind == "imu":
            fg.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7]); fc.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else:
            sel = np.sort(rng.choice(12, size=9, replace=False)) if i > 0 else np.arange(8)
            r1 = fg.processVisionData(seq.vision_stamps[i], seq.ids[sel], seq.bearings[i][sel]); r2 = fc.processVisionData(seq.vision_stamps[i], seq.ids[sel], seq.bearings[i][sel])
            a, b = fg.get_snapshot(), fc.get_snapshot()
            same = a.size == b.size
            print(f"bookkeeping frame {i} st {r1},{r2} N {fg.numLandmarks},{fc.N}", "Sigma rel %.2e" % rel(a[49+9*fc.N:], b[49+9*fc.N:]) if same else "SIZE MISMATCH", flush=True)
