"""Run the library GEMM once or a few times at one size / tile config (for ncu captures and A/B timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
n = int(sys.argv[1]); cfg = sys.argv[2]; tB = int(sys.argv[3]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
if cfg != "auto": os.environ["EQVIO_GEMM_CONFIG"] = cfg
from eqf_vio_b200.filter import dgemm
rng = np.random.default_rng(0)
A = rng.standard_normal((n, n)); B = rng.standard_normal((n, n))
C, ms = dgemm(A, B, transB=bool(tB), reps=reps)
if reps > 1:
    print(f"n={n} cfg={cfg} tB={tB} noedge={os.environ.get('EQVIO_GEMM_NOEDGE','0')} {ms:.4f} ms {2*n**3/ms/1e9:.2f} TFLOP/s", flush=True)
if "--cublas" in sys.argv:
    import torch
    a = torch.tensor(A, device="cuda"); b = torch.tensor(B, device="cuda")
    torch.matmul(a, b); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): torch.matmul(a, b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"n={n} cuBLAS {ms:.4f} ms {2*n**3/ms/1e9:.2f} TFLOP/s", flush=True)
