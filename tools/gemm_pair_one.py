"""Run the pair kernel (W = F S, D = W F^T in one launch) once or a few times at one size (for ncu captures and timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
from eqf_vio_b200.filter import dgemm_pair
rng = np.random.default_rng(0)
F = np.eye(n) + 1e-3 * rng.standard_normal((n, n)); S = rng.standard_normal((n, n))
W, D, ms = dgemm_pair(F, S, F, transB2=True, reps=reps)
if reps > 1:
    print(f"n={n} pair {ms:.4f} ms {4*n**3/ms/1e9:.2f} TFLOP/s", flush=True)
