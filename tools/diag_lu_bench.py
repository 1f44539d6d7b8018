#!/usr/bin/env python
"""Diagonal-block LU kernel of the Schur eliminations on its own: parity against an unpivoted numpy LU, time per
launch, and (with --clocks, against a -DEQVIO_DEBUG_CLOCKS build made by `--build-debug` on the CPU box) the
clock64 stamps between its phases.

    python tools/diag_lu_bench.py --build-debug        # here (no GPU): builds csrc/libeqvio_b200_dbg.so
    python tools/diag_lu_bench.py [--clocks]           # on the B200
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DBG = os.path.join(ROOT, "eqf_vio_b200", "csrc", "libeqvio_b200_dbg.so")


def lu_nopivot(A):
    A = A.copy()
    n = A.shape[0]
    for k in range(n - 1):
        A[k + 1 :, k] /= A[k, k]
        A[k + 1 :, k + 1 :] -= np.outer(A[k + 1 :, k], A[k, k + 1 :])
    return A


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build-debug", action="store_true")
    ap.add_argument("--clocks", action="store_true")
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    if args.build_debug:
        from eqf_vio_b200 import build

        print(build.build(force=True, extra_flags=["-DEQVIO_DEBUG_CLOCKS"], lib=DBG))
        return
    from eqf_vio_b200 import abi

    if args.clocks:
        abi.LIB_PATH = DBG
    from eqf_vio_b200.filter import getrf_block

    rng = np.random.default_rng(7)
    for nb in (64, 48, 5):
        B = rng.standard_normal((nb, nb))
        A = B @ B.T + nb * np.eye(nb) + 1e-3 * rng.standard_normal((nb, nb))  # SPD up to a small asymmetry
        LU, Li, Ui, us = getrf_block(A, reps=args.reps)
        ref = lu_nopivot(A)
        Lr, Ur = np.tril(ref, -1) + np.eye(nb), np.triu(ref)
        e_lu = np.abs(LU - ref).max() / np.abs(ref).max()
        e_li = np.abs(Li[:nb, :nb] - np.linalg.inv(Lr)).max()
        e_ui = np.abs(Ui[:nb, :nb] - np.linalg.inv(Ur)).max() / np.abs(np.linalg.inv(Ur)).max()
        print(f"nb={nb}: {us:.2f} us/launch, LU err {e_lu:.1e}, L^-1 err {e_li:.1e}, U^-1 err {e_ui:.1e}")
    if args.clocks:
        out = (C.c_longlong * 16)()
        abi.lib().eqvio_debug_clocks(out)
        c = list(out)
        print("clock64 phase marks 0-4 (load, LU, inverses, store):", [c[i + 1] - c[i] for i in range(4)])
        print("sub-step 0, marks 5-11 (LU8, bar, panels, bar, trailing, bar):", [c[i + 1] - c[i] for i in range(5, 11)])


if __name__ == "__main__":
    main()
