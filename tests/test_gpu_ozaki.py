"""fp64 GEMM assembled from int8 tensor-core products (tcgen05.mma kind::i8, TMEM accumulators; csrc/ozaki_sm100.cu) through
the opt-in C-ABI entry eqvio_dgemm_ozaki, against numpy.  The products are the reference's dense Sigma contractions
(eqf_vio/src/VIOFilter.cpp:188-189, 276-277, 297)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b))


@pytest.mark.parametrize("shape", [(128, 128, 32), (256, 128, 96), (384, 256, 200), (139, 139, 139), (300, 450, 77), (779, 779, 779)])
def test_ozaki_matches_numpy(shape):
    """Gaussian operands, 8 slices: as accurate as an fp64 GEMM (1e-14 relative Frobenius is asserted, ~5e-16 observed); ragged
    shapes exercise the DMMA strips in front of the 128-aligned core block and the zero padding of the slice arrays."""
    from eqf_vio_b200.filter import dgemm_ozaki

    M, N, K = shape
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)); B = rng.standard_normal((K, N))
    C, _, _ = dgemm_ozaki(A, B, slices=8)
    assert rel(C, A @ B) < 1e-14
    Ct, _, _ = dgemm_ozaki(A, np.asfortranarray(B.T), transB=True, slices=8)
    assert np.array_equal(C, Ct)          # same digits, same integer sums: the operand layout cannot matter


def test_ozaki_sigma_like_operands_at_the_headline_size():
    """n = 1547 (N = 512), F = I + T A and a Sigma-like SPD matrix whose entries span 1e-4 ... 5e3 (the dynamic range of the
    filter's covariance on the template start-up): the verdict's criterion, rel-Frobenius <= 1e-13 against numpy, with 8 and
    9 slices; 9 slices also hold entry by entry at the level of an fp64 GEMM."""
    from eqf_vio_b200.filter import dgemm_ozaki

    rng = np.random.default_rng(7)
    n = 1547
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    S = (Q * 10.0 ** rng.uniform(-4, 3.7, n)) @ Q.T
    d = 10.0 ** rng.uniform(-1, 1, n)
    S = S * d[:, None] * d[None, :]
    F = np.eye(n) + 0.005 * rng.standard_normal((n, n))
    ref = (F.astype(np.longdouble) @ S.astype(np.longdouble)).astype(np.float64)
    W8, _, _ = dgemm_ozaki(F, S, slices=8)
    W9, _, _ = dgemm_ozaki(F, S, slices=9)
    assert rel(W8, ref) < 1e-13 and rel(W9, ref) < 1e-14, (rel(W8, ref), rel(W9, ref))
    Wf = F @ S
    ent = lambda X: float(np.max(np.abs(X - ref) / np.maximum(np.abs(ref), 1e-300)))
    assert ent(W9) < 50 * max(ent(Wf), 1e-13), (ent(W9), ent(Wf))
    # second product of the Riccati step, N-major operand
    D8, _, _ = dgemm_ozaki(W8, F, transB=True, slices=8)
    assert rel(D8, W8 @ F.T) < 1e-13


def test_ozaki_is_deterministic_and_timed():
    from eqf_vio_b200.filter import dgemm_ozaki

    rng = np.random.default_rng(3)
    A = rng.standard_normal((512, 640)); B = rng.standard_normal((640, 384))
    C1, t_all, t_gemm = dgemm_ozaki(A, B, slices=8, reps=3)
    C2, _, _ = dgemm_ozaki(A, B, slices=8)
    assert np.array_equal(C1, C2) and t_all > 0 and t_gemm > 0


def test_ozaki_rejects_what_it_cannot_do():
    from eqf_vio_b200 import abi
    from eqf_vio_b200.filter import dgemm_ozaki

    A = np.ones((64, 64))
    with pytest.raises(abi.EqvioError):
        dgemm_ozaki(A, A, slices=8)        # smaller than one 128 x 128 tile
    A = np.ones((128, 128))
    with pytest.raises(abi.EqvioError):
        dgemm_ozaki(A, A, slices=12)
