"""How far apart do fp64 implementations of the reference's filter drift when free-running on the TEMPLATE start-up
(initialPointVariance 5000, initialSceneDepth 1 against a 3-15 m scene), over a full-length sequence?

Four fp64 evaluations of the same recursion on the same inputs are compared with the C restatement (oracle/eqvio_oracle.c):
  * the reference's own sources (oracle/_ref, Eigen stand-in),
  * the numpy restatement (LAPACK inverses, BLAS products: a different summation order),
  * the C restatement fed inputs perturbed by ONE ULP (random sign, every IMU and bearing value): the recursion's own
    sensitivity to round-off-sized perturbations — no implementation can be expected to agree with another more closely,
  * (--gpu) the B200 path.
Printed per frame: rel-Frobenius difference of Sigma and the max abs difference of the lifted state, for the template
start-up and for the conditioned start-up the benchmark uses.  Complements tools/accuracy_vs_truth.py (one update against a
50-digit ground truth).

    python tools/template_spread.py [N] [frames] [--gpu]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from eqf_vio_b200.settings import conditioned_settings, template_settings
from eqf_vio_b200.synthetic import period_sequence
from oracle import eqvio_numpy as onp
from oracle import ref_binding
from oracle.c_oracle import COracleFilter

args = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(args[0]) if args else 16
frames = int(args[1]) if len(args) > 1 else 60
use_gpu = "--gpu" in sys.argv
report = sorted(set([1, 2, 3, 5, 10, 20, 40, 60, 100, 200, frames]) & set(range(frames + 1)))


def split(d):
    n_l = int(d[0]); hn = 49 + 9 * n_l; n = 11 + 3 * n_l
    return d[:hn], d[hn:hn + n * n]


def ulp_perturb(a, rng):
    a = np.array(a, dtype=np.float64, copy=True)
    return np.where(rng.random(a.shape) < 0.5, np.nextafter(a, np.inf), np.nextafter(a, -np.inf))


for label, s in (("template start-up (variance 5000, depth 1 m)", template_settings(outlierThreshold=1e9)),
                 ("conditioned start-up (variance 100, depth 8 m) - the benchmark workload", conditioned_settings())):
    seq = period_sequence(N, frames, camera_offset=tuple(s.cameraOffset))
    rng = np.random.default_rng(1)
    imu_p, y_p = seq.imu.copy(), seq.bearings.copy()
    imu_p[:, 1:] = ulp_perturb(seq.imu[:, 1:], rng)
    y_p = ulp_perturb(seq.bearings, rng)
    impls = {"C oracle, inputs +-1 ulp": COracleFilter(s), "numpy restatement": onp.VIOFilter(onp.Settings(**s.as_dict()))}
    if ref_binding.available():
        impls["reference sources (shim)"] = ref_binding.ReferenceFilter(s)
    if use_gpu:
        from eqf_vio_b200.filter import VIOFilter
        impls["B200"] = VIOFilter(s)
    base = COracleFilter(s)
    print(f"\n### N = {N}, {label}\n")
    print("| frame | " + " | ".join(f"{k}: Sigma rel / state abs" for k in impls) + " |")
    print("|---|" + "---|" * len(impls))
    for kind, i in seq.events():
        for name, f in [("base", base)] + list(impls.items()):
            pert = name.startswith("C oracle, inputs")
            if kind == "imu":
                row = imu_p[i] if pert else seq.imu[i]
                f.processIMUData(seq.imu[i, 0], row[1:4], row[4:7])
            else:
                f.processVisionData(seq.vision_stamps[i], seq.ids, (y_p if pert else seq.bearings)[i])
        if kind == "vision" and i in report:
            hb, Sb = split(base.get_snapshot())
            cells = []
            for name, f in impls.items():
                h, S = split(f.get_snapshot())
                cells.append(f"{np.linalg.norm(S - Sb) / np.linalg.norm(Sb):.1e} / {np.abs(h - hb).max():.1e}")
            print(f"| {i} | " + " | ".join(cells) + " |")
