"""Ranks the implementations against a 50-digit ground truth (mpmath) for one vision update from an
identical state: gain K / innovation gamma, updated Sigma and the lifted SE(3) increment.  Inputs to
the truth are the fp64 quantities every implementation starts from (Sigma, C, delta, and the bundle-lift
blocks D, M, obs), so the differences isolate the dense algebra: S^-1, K, Sigma - K C Sigma, Sigma_sub^-1.

    python tools/accuracy_vs_truth.py [N] [--gpu]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpmath as mp
import numpy as np
from eqf_vio_b200.settings import template_settings
from eqf_vio_b200.synthetic import period_sequence
from oracle import eqvio_numpy as onp
from oracle.c_oracle import COracleFilter

mp.mp.dps = 50
N = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 16
use_gpu = "--gpu" in sys.argv

def tomp(A): return mp.matrix(np.asarray(A, dtype=float).tolist())
def tonp(M): return np.array([[float(M[i, j]) for j in range(M.cols)] for i in range(M.rows)])
def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))

s = template_settings(outlierThreshold=1e9)
seq = period_sequence(N, 3, camera_offset=tuple(s.cameraOffset))
o = COracleFilter(s)
fn = onp.VIOFilter(onp.Settings(**s.as_dict()))
fg = None
if use_gpu:
    from eqf_vio_b200.filter import VIOFilter
    fg = VIOFilter(s)
for kind, i in seq.events():
    if kind == "imu":
        o.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        continue
    if i == 0:
        o.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
        continue
    # bring every implementation to the oracle's state at the frame time (propagate on the oracle only)
    y = seq.bearings[i]
    # emulate integrateUpToTime on the oracle through a zero-measurement-free path: use numpy filter for propagate
    fn.set_snapshot(o.get_snapshot())
    fn.integrateUpToTime(seq.vision_stamps[i])
    snap = fn.get_snapshot()
    o.set_snapshot(snap)
    C, delta = fn.build_C_delta(y)
    S0 = fn.Sigma.copy()
    n = S0.shape[0]
    # ---- truth ----
    Cm, Sm = tomp(C), tomp(S0)
    Smat = Cm * Sm * Cm.T + mp.mpf(s.measurementVariance) * mp.eye(2 * N)
    Km = Sm * Cm.T * mp.inverse(Smat)
    gm = Km * tomp(delta.reshape(-1, 1))
    Snew = Sm - Km * Cm * Sm
    K_t, g_t, Sn_t = tonp(Km), tonp(gm).reshape(-1), tonp(Snew)
    # ---- implementations ----
    res = {}
    oc = COracleFilter(s); oc.set_snapshot(snap)
    Kc, gc = oc.gain_update(y); res["C oracle (LU inverse)"] = (Kc, gc, oc.stateCovariance())
    Sn = C @ S0 @ C.T + s.measurementVariance * np.eye(2 * N)
    Kn = S0 @ C.T @ np.linalg.inv(Sn); res["numpy (LAPACK inverse)"] = (Kn, Kn @ delta, S0 - Kn @ C @ S0)
    if fg is not None:
        fg.set_snapshot(snap)
        Kg, gg = fg.gain_update(y); res["B200 (Cholesky, X X^T)"] = (Kg, gg, fg.stateCovariance())
    w = np.linalg.eigvalsh((S0 + S0.T) / 2); ws = np.linalg.eigvalsh((Sn + Sn.T) / 2)
    print(f"frame {i}: N={N} cond(Sigma)={w.max()/w.min():.1e} cond(S)={ws.max()/ws.min():.1e}")
    for name, (K, g, Sig) in res.items():
        print(f"   {name:26s} K rel {rel(K, K_t):.2e}  gamma rel {rel(g, g_t):.2e}  Sigma+ rel {rel(Sig, Sn_t):.2e}")
    # ---- bundle lift: Gamma[0:6] for the true gamma ----
    g_eqf = g_t[6:]
    Ssub = S0[6:, 6:]
    G_np = onp.bundle_lift(g_eqf, fn.xi0, fn.X, Ssub)
    oc.set_snapshot(snap); G_c = oc.bundle_lift(g_eqf)
    # truth through the normal equations with Sigma_sub^-1 in 50 digits: reuse numpy's blocks
    xiHat = onp.state_group_action(fn.X, fn.xi0)
    eta0 = onp.project_to_manifold(fn.xi0).gravityDir; eta0 = eta0 / np.linalg.norm(eta0)
    KPara = np.zeros((6, 4)); KPara[0:3, 0] = eta0; KPara[3:6, 1:4] = np.eye(3)
    KPerp = np.zeros((6, 6)); KPerp[0:3, 0:3] = np.eye(3) - np.outer(eta0, eta0)
    DU0 = np.zeros(6); DU0[0:3] = -onp.skew(eta0) @ onp.stereo_sphere_chart_inv_diff(np.zeros(2), eta0) @ g_eqf[0:2]
    DUF = KPerp @ DU0
    R_C = onp.q_mul(xiHat.pose.R, xiHat.cameraOffset.R); R_CT = onp.q_mat(onp.q_inv(R_C)); AdP0 = fn.xi0.pose.adjoint()
    PT = xiHat.pose * xiHat.cameraOffset
    coeff = np.zeros((3 * N, 4)); obs = np.zeros(3 * N); D = np.zeros((5 + 3 * N, 3 * N))
    for k in range(N):
        gq = g_eqf[5 + 3 * k: 8 + 3 * k]
        pHat = PT * xiHat.landmarks[k]
        alpha = -onp.q_rot(R_C, fn.X.Q[k].inverse() * gq)
        pm = np.hstack([-onp.skew(pHat), np.eye(3)])
        obs[3 * k:3 * k + 3] = alpha - pm @ AdP0 @ DUF
        coeff[3 * k:3 * k + 3] = pm @ AdP0 @ KPara
        D[5 + 3 * k:8 + 3 * k, 3 * k:3 * k + 3] = fn.X.Q[k].as_matrix3() @ R_CT
    Wm = tomp(D).T * mp.inverse(tomp(Ssub)) * tomp(D)
    A4 = tomp(coeff).T * Wm * tomp(coeff); b4 = tomp(coeff).T * Wm * tomp(obs.reshape(-1, 1))
    x = mp.lu_solve(A4, b4)
    DU_t = DUF + KPara @ np.array([float(v) for v in x])
    line = f"   bundleLift DeltaU abs err: C {np.abs(G_c[:6]-DU_t).max():.2e}  numpy {np.abs(G_np[:6]-DU_t).max():.2e}"
    if fg is not None:
        fg.set_snapshot(snap); G_g = fg.bundle_lift(g_eqf)
        line += f"  B200 {np.abs(G_g[:6]-DU_t).max():.2e}"
    print(line + f"   (|DeltaU| = {np.abs(DU_t).max():.2e})")
    # advance the oracle with the real update
    o.set_snapshot(snap)
    # process the frame on the C oracle from the pre-integration state is not possible after set_snapshot(time) -> apply update pieces:
    fn.processVisionData(seq.vision_stamps[i] + 1e-9, seq.ids, y)  # dt = 1e-9 propagate (negligible), then the update
    o.set_snapshot(fn.get_snapshot())
