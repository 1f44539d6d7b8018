"""The covariance update's evaluation order (host-side reasoning behind a kernel-side choice, CPU only).

The reference evaluates Sigma - K*C*Sigma left to right, (K C) Sigma (eqf_vio/src/VIOFilter.cpp:297).  The B200 path evaluates
Sigma - K (C Sigma) with the C Sigma it already formed for S = (C Sigma) C^T (:276): one product instead of two.  On the numpy
oracle's own K, C and prior Sigma the two orders agree to rounding; what must NOT be used is K (Sigma C^T)^T — the reference never
symmetrises Sigma, so Sigma C^T is not (C Sigma)^T once an update has run."""
import numpy as np

import oracle.eqvio_numpy as on
from eqf_vio_b200.settings import template_settings
from eqf_vio_b200.synthetic import period_sequence
from helpers import feed


def test_K_times_CSigma_is_the_reference_update_up_to_rounding():
    N = 24
    s = template_settings(outlierThreshold=1e9)
    seq = period_sequence(N, 4, camera_offset=tuple(s.cameraOffset))
    cls = [c for c in vars(on).values() if isinstance(c, type) and hasattr(c, "processVisionData")][0]
    o = cls(s)
    rel = lambda x, y: np.linalg.norm(x - y) / np.linalg.norm(y)
    prior = {}
    orig = cls.build_C_delta

    def patched(self, m_y):
        prior["S"] = self.Sigma.copy()
        return orig(self, m_y)

    cls.build_C_delta = patched
    asym = []
    try:
        for kind, i in seq.events():
            feed(o, seq, kind, i)
            if kind != "vision" or "S" not in prior:
                continue
            K, C, Sg = o.last["K"], o.last["C"], prior.pop("S")
            ref = Sg - (K @ C) @ Sg                      # the reference's order
            ours = Sg - K @ (C @ Sg)                     # the B200 path's order
            assert rel(ours, ref) < 1e-13, (i, rel(ours, ref))
            assert np.array_equal(ref, o.Sigma)          # (the oracle evaluates the reference's order)
            asym.append(np.abs(Sg - Sg.T).max() / np.abs(Sg).max())
    finally:
        cls.build_C_delta = orig
    # the prior Sigma of a later frame is no longer symmetric to machine precision: the transposed shortcut would see that
    assert len(asym) >= 3 and asym[0] < 1e-18 and asym[-1] > 1e-16, asym
