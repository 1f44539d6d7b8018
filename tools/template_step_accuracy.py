"""One vision frame (Riccati step + update) from an oracle state at N = 512 with the TEMPLATE settings (initialPointVariance 5000, depth 1 m: the
ill-conditioned start-up), B200 path against the C oracle: with the update's products on the int8 path (default) and on fp64 DMMA.
    python tools/template_step_accuracy.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from eqf_vio_b200.filter import VIOFilter
from eqf_vio_b200.settings import template_settings
from eqf_vio_b200.synthetic import period_sequence
from helpers import run, split_snapshot, rel
from oracle.c_oracle import COracleFilter

s = template_settings(outlierThreshold=1e9)
seq = period_sequence(512, 1, camera_offset=tuple(s.cameraOffset))
o = COracleFilter(s)
run(o, seq, ("vision", 1))
snap = o.get_snapshot()
o.processVisionData(seq.vision_stamps[1], seq.ids, seq.bearings[1])
h2, S2 = split_snapshot(o.get_snapshot())
print("| update's products | Riccati step | rel-Frobenius(Sigma) vs C oracle | max state error / max(1, |entry|) |")
print("|---|---|---|---|")
for label, env in (("int8, 8 slices (default)", {}), ("fp64 DMMA", {"EQVIO_OZ_PRE": "0", "EQVIO_OZ_SCT": "0", "EQVIO_OZ_UPDATE": "0"}), ("fp64 DMMA", {"EQVIO_OZAKI": "0"}),
                   ("int8, 7 slices", {"EQVIO_OZAKI": "7"}), ("int8, 9 slices", {"EQVIO_OZAKI": "9"})):
    for k in ("EQVIO_OZ_PRE", "EQVIO_OZ_SCT", "EQVIO_OZ_UPDATE", "EQVIO_OZAKI"):
        os.environ.pop(k, None)
    os.environ.update(env)
    f = VIOFilter(s, device=0)
    f.set_snapshot(snap)
    f.processVisionData(seq.vision_stamps[1], seq.ids, seq.bearings[1])
    h1, S1 = split_snapshot(f.get_snapshot())
    print(f"| {label} | {('int8, %d slices' % f.riccati_int8_slices()) if f.riccati_int8_slices() else 'fp64 DMMA'} | {rel(S1, S2):.2e} | {(np.abs(h1 - h2) / np.maximum(1.0, np.abs(h2))).max():.2e} |")
    f.close()
