"""Library GEMM at the filter's sizes over a list of tile configs (run under gpurun): best of two passes per case."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eqf_vio_b200.filter import dgemm
cfgs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["3", "8", "9", "10"]
sizes = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["779"])]
rng = np.random.default_rng(0)
for n in sizes:
    A = rng.standard_normal((n, n)); B = rng.standard_normal((n, n))
    ref = {False: A @ B, True: A @ B.T}
    best = {}
    for rep in range(2):
        for cfg in cfgs:
            os.environ["EQVIO_GEMM_CONFIG"] = cfg
            for tB in (False, True):
                C, ms = dgemm(A, B, transB=tB, reps=30 if n > 1000 else 100)
                err = np.linalg.norm(C - ref[tB]) / np.linalg.norm(ref[tB])
                assert err < 1e-13, (n, cfg, tB, err)
                best[(cfg, tB)] = min(best.get((cfg, tB), 1e9), ms)
    for (cfg, tB), ms in best.items():
        print(f"n={n} cfg={cfg} tB={int(tB)} {ms*1e3:.1f} us {2*n**3/ms/1e9:.2f} TFLOP/s", flush=True)
