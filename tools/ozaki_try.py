"""Bring-up / timing of the int8 tensor-core fp64 GEMM (eqvio_dgemm_ozaki).   python tools/ozaki_try.py [quick]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eqf_vio_b200.filter import dgemm, dgemm_ozaki

def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
rng = np.random.default_rng(0)
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
shapes = [(128, 128, 32), (128, 128, 64), (256, 128, 96), (256, 256, 256), (384, 256, 200), (139, 139, 139)]
if not quick:
    shapes += [(1536, 1536, 1536), (1547, 1547, 1547), (3083, 3083, 3083)]
for (M, N, K) in shapes:
    A = rng.standard_normal((M, K)); B = rng.standard_normal((K, N))
    for S in ((8,) if quick else (7, 8, 9)):
        C, t_all, t_g = dgemm_ozaki(A, B, slices=S, reps=(1 if quick else 10))
        ref = A @ B
        line = f"M={M} N={N} K={K} S={S}: rel {rel(C, ref):.2e}"
        if t_g > 0:
            line += f"  whole {t_all*1e3:.1f} us = {2.0*M*N*K/t_all/1e9:.1f} TF-equiv, tcgen05 kernel {t_g*1e3:.1f} us = {2.0*(M//128*128)*(N//128*128)*K/t_g/1e9:.1f} TF-equiv"
        print(line, flush=True)
    if not quick and M >= 1536:
        Bt = np.asfortranarray(B.T)
        C, t_all, t_g = dgemm_ozaki(A, Bt, transB=True, slices=8, reps=10)
        print(f"   transB: rel {rel(C, A @ B):.2e} whole {t_all*1e3:.1f} us", flush=True)
        _, ms = dgemm(A, B, reps=10)
        print(f"   DMMA kernel: {ms*1e3:.1f} us = {2.0*M*N*K/ms/1e9:.1f} TF", flush=True)
