"""Sigma - (K C) Sigma (the reference's evaluation order, VIOFilter.cpp:297) against Sigma - K (C Sigma) (what the B200 path evaluates:
C Sigma is already there from S = (C Sigma) C^T) and against Sigma - K (Sigma C^T)^T (NOT used: the reference never symmetrises
Sigma, so Sigma C^T is not (C Sigma)^T), on the numpy oracle's own K, C and prior Sigma with the template settings (CPU only).
    python tools/sigma_update_association.py [N]"""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import numpy as np
from eqf_vio_b200.settings import template_settings
from eqf_vio_b200.synthetic import period_sequence
from helpers import run
import oracle.eqvio_numpy as on

N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
s = template_settings(outlierThreshold=1e9)
seq = period_sequence(N, 3, camera_offset=tuple(s.cameraOffset))
cls = [c for c in vars(on).values() if isinstance(c, type) and hasattr(c, "processVisionData")][0]
o = cls(s)
run(o, seq, ("vision", 0))
f = lambda x, y: np.linalg.norm(x - y) / np.linalg.norm(y)
print("| frame | cond(S) | asymmetry of the prior Sigma | (KC)Sigma vs K(C Sigma) | vs K(Sigma C^T)^T | (KC)Sigma vs long double | K(C Sigma) vs long double |")
print("|---|---|---|---|---|---|---|")
for fr in (1, 2, 3):
    prior = {}
    orig = cls.build_C_delta
    def patched(self, m_y):
        prior["S"] = self.Sigma.copy()
        return orig(self, m_y)
    cls.build_C_delta = patched
    run(o, seq, ("vision", fr))
    cls.build_C_delta = orig
    K, C, Sg = o.last["K"], o.last["C"], prior["S"]
    a, b, c = Sg - (K @ C) @ Sg, Sg - K @ (C @ Sg), Sg - K @ (Sg @ C.T).T
    Kl, Cl, Sl = (x.astype(np.longdouble) for x in (K, C, Sg))
    ref = Sl - (Kl @ Cl) @ Sl
    print(f"| {fr} | {np.linalg.cond(o.last['S']):.1e} | {np.abs(Sg - Sg.T).max() / np.abs(Sg).max():.1e} | {f(b, a):.1e} | {f(c, a):.1e} | {f(a, ref):.1e} | {f(b, ref):.1e} |")
