"""One shape of the int8 tensor-core GEMM for ncu.   ncu ... python tools/ozaki_one.py [n] [slices]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eqf_vio_b200.filter import dgemm_ozaki
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
S = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rng = np.random.default_rng(0)
A = rng.standard_normal((n, n)); B = rng.standard_normal((n, n))
C, t_all, t_g = dgemm_ozaki(A, B, slices=S, reps=3)
print(n, S, float(np.linalg.norm(C - A @ B) / np.linalg.norm(A @ B)), t_all, t_g)
