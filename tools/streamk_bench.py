"""Stream-K vs ticketed-pair vs two launches at the single-wave sizes (kernel-level entry point, L2-warm repeats).
    python tools/streamk_bench.py"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

if len(sys.argv) > 1:
    from eqf_vio_b200.filter import dgemm, dgemm_pair
    n = int(sys.argv[1])
    rng = np.random.default_rng(0)
    F = np.eye(n) + 1e-3 * rng.standard_normal((n, n)); S = rng.standard_normal((n, n))
    _, _, ms = dgemm_pair(F, S, F, transB2=True, reps=50)
    fl = 4.0 * n ** 3
    extra = ""
    if os.environ.get("EQVIO_STREAMK") == "0":
        _, m1 = dgemm(F, S, reps=50)
        _, m2 = dgemm(S, F, transB=True, reps=50)
        extra = f" | two launches {1e3*(m1+m2):.1f} us = {fl/(m1+m2)/1e9:.2f} TF"
    print(f"n={n} STREAMK={os.environ.get('EQVIO_STREAMK','-')} SPLITK={os.environ.get('EQVIO_SPLITK','-')} pair {ms*1e3:.1f} us = {fl/ms/1e9:.2f} TF{extra}")
else:
    for n in (395, 587, 779, 971, 1163):
        for mode, split in (("0", "0"), ("0", "2"), ("0", "3"), ("0", "4"), ("1", "0")):
            env = dict(os.environ, EQVIO_STREAMK=mode, EQVIO_PAIR_FORCE="1", EQVIO_SPLITK=split)
            subprocess.run([sys.executable, __file__, str(n)], env=env)
