"""Long free-running comparison of the int8 tensor-core Riccati path (default) with the fp64 DMMA path (EQVIO_OZAKI=0) on the same
inputs: rel-Frobenius difference of Sigma, the largest entry difference relative to sqrt(Sigma_ii Sigma_jj), and the largest state
difference, every `every` vision periods.      python tools/int8_vs_dmma_long.py [N=512] [periods=400] [every=50]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from eqf_vio_b200.filter import VIOFilter
from eqf_vio_b200.settings import conditioned_settings
from eqf_vio_b200.synthetic import period_sequence

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
periods = int(sys.argv[2]) if len(sys.argv) > 2 else 400
every = int(sys.argv[3]) if len(sys.argv) > 3 else 50
s = conditioned_settings(outlierThreshold=1e9)
seq = period_sequence(N, periods + 1, camera_offset=tuple(s.cameraOffset))
os.environ["EQVIO_OZAKI"] = "0"
d = VIOFilter(s, device=0)
del os.environ["EQVIO_OZAKI"]
i8 = VIOFilter(s, device=0)


def split(x):
    n_l = int(x[0]); hn = 49 + 9 * n_l; n = 11 + 3 * n_l
    return x[:hn], x[hn:hn + n * n].reshape(n, n, order="F")


print(f"N = {N} (n = {11 + 3 * N}), conditioned start-up, {periods} vision periods = {11 * periods} filter steps = {periods * 0.05:.0f} s of IMU 200 Hz / vision 20 Hz")
print("| period | filter steps | int8 slices | rel-Frobenius(Sigma_int8 - Sigma_dmma) | max entry diff / sqrt(S_ii S_jj) | max state diff |")
print("|---|---|---|---|---|---|")
for kind, i in seq.events():
    for f in (d, i8):
        if kind == "imu":
            f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
        else:
            f.processVisionData(seq.vision_stamps[i], seq.ids, seq.bearings[i])
    if kind == "vision" and i > 0 and (i % every == 0 or i == periods):
        h1, S1 = split(i8.get_snapshot()); h0, S0 = split(d.get_snapshot())
        dg = np.sqrt(np.abs(np.diag(S0)))
        print(f"| {i} | {11 * i} | {i8.riccati_int8_slices()} | {np.linalg.norm(S1 - S0) / np.linalg.norm(S0):.2e} | {(np.abs(S1 - S0) / (dg[:, None] * dg[None, :])).max():.2e} | {np.abs(h1 - h0).max():.2e} |", flush=True)
