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


# ---------------------------------------------------------------------------------------------------------------------------------
# The Riccati step of the filter on the int8 tensor cores (VIOFilter.cpp:188-189): fused form (two launches of k_oz_riccati, each
# emitting its result as the next product's int8 operand), unfused form (split kernels + DMMA strips between the products) and the
# fp64 DMMA path must agree to round-off with each other and with the CPU oracle.
# ---------------------------------------------------------------------------------------------------------------------------------
_MID = {}


def _events_after_vision(seq, j):
    ev = list(seq.events())
    k = max(i for i, (kind, idx) in enumerate(ev) if kind == "vision" and idx == j)
    return ev[: k + 1], ev[k + 1 :]


def _riccati_run(monkeypatch, N, env, ticks):
    """Seeds a filter with the state behind the first vision update (computed once per N on the fp64 DMMA path) and runs `ticks`
    IMU ticks with the environment `env`; returns (seed snapshot, final snapshot, int8 slices in use)."""
    from eqf_vio_b200.filter import VIOFilter
    from eqf_vio_b200.settings import conditioned_settings
    from eqf_vio_b200.synthetic import period_sequence
    from helpers import feed

    for k in ("EQVIO_OZAKI", "EQVIO_OZAKI_FUSED", "EQVIO_OZAKI_MIN_TILES"):
        monkeypatch.delenv(k, raising=False)
    s = conditioned_settings(outlierThreshold=1e9)
    seq = period_sequence(N, 3, camera_offset=tuple(s.cameraOffset))
    head, tail = _events_after_vision(seq, 1)
    if N not in _MID:
        monkeypatch.setenv("EQVIO_OZAKI", "0")
        f = VIOFilter(s, device=0)
        for kind, i in head:
            feed(f, seq, kind, i)
        _MID[N] = f.get_snapshot()
        f.close()
        monkeypatch.delenv("EQVIO_OZAKI")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    f = VIOFilter(s, device=0)
    f.set_snapshot(_MID[N])
    slices = None
    for kind, i in tail[:ticks]:
        assert kind == "imu"
        feed(f, seq, kind, i)
        slices = f.riccati_int8_slices()
    out = f.get_snapshot()
    f.close()
    return _MID[N], out, slices


@pytest.mark.parametrize("N", [512, 500, 470, 511])
def test_fused_riccati_matches_unfused_dmma_and_oracle(monkeypatch, N):
    """N = 512: n = 1547 = 11 + 12 x 128 (the border is the 11 base states).  N = 500: n = 1511 = 103 + 11 x 128 and N = 470:
    n = 1421 = 13 + 11 x 128 — the border holds landmark rows as well, several border jobs per tile row.  N = 511: n = 1544 =
    12 x 128 + 8, where n mod 128 is smaller than the 11 base states (the block is 128 floor(3N / 128) = 1408, the border 136).  Nine IMU ticks behind the
    first vision update: the first tick splits Sigma with the generic kernels, the others run on what the previous launch emitted."""
    from eqf_vio_b200.settings import conditioned_settings
    from eqf_vio_b200.synthetic import period_sequence
    from helpers import feed, split_snapshot
    from oracle.c_oracle import COracleFilter

    mid, out_f, sl_f = _riccati_run(monkeypatch, N, {}, 9)
    _, out_u, sl_u = _riccati_run(monkeypatch, N, {"EQVIO_OZAKI_FUSED": "0"}, 9)
    _, out_d, sl_d = _riccati_run(monkeypatch, N, {"EQVIO_OZAKI": "0"}, 9)
    assert sl_f == 8 and sl_u == 8 and sl_d == 0
    hf, Sf = split_snapshot(out_f); hu, Su = split_snapshot(out_u); hd, Sd = split_snapshot(out_d)
    assert np.array_equal(hf, hd) and np.array_equal(hu, hd)                  # the state propagate does not depend on Sigma
    assert rel(Sf, Sd) < 2e-14 and rel(Su, Sd) < 2e-14 and rel(Sf, Su) < 2e-14, (rel(Sf, Sd), rel(Su, Sd), rel(Sf, Su))
    # entry by entry, relative to sqrt(Sigma_ii Sigma_jj) — the scale the inner-dimension equilibration guarantees
    d = np.sqrt(np.abs(np.diag(Sd)))
    ent = np.abs(Sf - Sd) / (d[:, None] * d[None, :])
    assert ent.max() < 1e-12, ent.max()
    # and against the oracle seeded with the same state
    s = conditioned_settings(outlierThreshold=1e9)
    seq = period_sequence(N, 3, camera_offset=tuple(s.cameraOffset))
    _, tail = _events_after_vision(seq, 1)
    o = COracleFilter(s)
    o.set_snapshot(mid)
    for kind, i in tail[:9]:
        feed(o, seq, kind, i)
    ho, So = split_snapshot(o.get_snapshot())
    assert rel(Sf, So) < 1e-12 and np.abs(hf - ho).max() < 1e-9, (rel(Sf, So), np.abs(hf - ho).max())


def test_fused_riccati_is_deterministic(monkeypatch):
    """Atomic maxima, tickets and tile-row barriers order nothing that reaches the result: two runs agree bit for bit."""
    a = _riccati_run(monkeypatch, 512, {}, 10)[1]
    b = _riccati_run(monkeypatch, 512, {}, 10)[1]
    assert np.array_equal(a, b)


@pytest.mark.parametrize("N", [512, 470])
def test_int8_covariance_update_matches_dmma(monkeypatch, N):
    """The vision update's dense products — C Sigma, (C Sigma) C^T, Sigma C^T, (Sigma C^T) S^-1 and K (C Sigma) (VIOFilter.cpp:276-277,
    297) — on the int8 tensor cores, C Sigma against the slices of the prior Sigma left by the last Riccati launch, against the same
    update on fp64 DMMA: one whole vision period (10 IMU ticks + the frame's own propagate + the update) from the same state, Riccati
    steps on int8 in both runs."""
    from eqf_vio_b200.filter import VIOFilter
    from eqf_vio_b200.settings import conditioned_settings
    from eqf_vio_b200.synthetic import period_sequence
    from helpers import feed, split_snapshot

    mid = _riccati_run(monkeypatch, N, {}, 0)[0]
    s = conditioned_settings(outlierThreshold=1e9)
    seq = period_sequence(N, 3, camera_offset=tuple(s.cameraOffset))
    _, tail = _events_after_vision(seq, 1)
    outs = []
    for upd in ("1", "0"):
        monkeypatch.setenv("EQVIO_OZ_UPDATE", upd)
        monkeypatch.setenv("EQVIO_OZ_PRE", upd)
        monkeypatch.setenv("EQVIO_OZ_SCT", upd)
        f = VIOFilter(s, device=0)
        f.set_snapshot(mid)
        for kind, i in tail[:11]:
            feed(f, seq, kind, i)
        assert tail[10][0] == "vision"
        outs.append(f.get_snapshot())
        f.close()
    for k in ("EQVIO_OZ_UPDATE", "EQVIO_OZ_PRE", "EQVIO_OZ_SCT"):
        monkeypatch.delenv(k)
    (h1, S1), (h0, S0) = split_snapshot(outs[0]), split_snapshot(outs[1])
    assert np.abs(h1 - h0).max() < 1e-10, np.abs(h1 - h0).max()   # S, K and gamma come from int8 products in one run, from DMMA in the other
    assert not np.array_equal(S1, S0)                   # (the int8 path was taken)
    d = np.sqrt(np.abs(np.diag(S0)))
    assert rel(S1, S0) < 1e-11 and (np.abs(S1 - S0) / (d[:, None] * d[None, :])).max() < 1e-10, (rel(S1, S0), (np.abs(S1 - S0) / (d[:, None] * d[None, :])).max())


def test_int8_paths_when_landmarks_come_and_go(monkeypatch):
    """N = 500 features; after five stable frames (their launch sequences are captured as CUDA graphs) ten ids drop out for one frame and
    come back: Sigma is compacted and grown, the slice layouts of every int8 operand follow n, the update splits Sigma itself because
    the Riccati launch's slices are stale.  (i) Graph replay and direct launches must agree bit for bit — what a captured sequence
    depends on (which slice arrays must be cleared first, which exponent array is current ...) is part of its key; (ii) the int8 paths
    must agree with the fp64 DMMA paths to round-off."""
    from eqf_vio_b200.filter import VIOFilter
    from eqf_vio_b200.settings import conditioned_settings
    from eqf_vio_b200.synthetic import period_sequence
    from helpers import split_snapshot

    N = 500
    s = conditioned_settings(outlierThreshold=1e9)
    seq = period_sequence(N, 12, camera_offset=tuple(s.cameraOffset))
    drop = np.arange(40, 50)

    def run_it(env):
        for k in ("EQVIO_OZAKI", "EQVIO_GRAPHS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        f = VIOFilter(s, device=0)
        for kind, i in seq.events():
            if kind == "imu":
                f.processIMUData(seq.imu[i, 0], seq.imu[i, 1:4], seq.imu[i, 4:7])
            else:
                sel = np.setdiff1d(np.arange(N), drop) if i in (6, 9) else np.arange(N)
                f.processVisionData(seq.vision_stamps[i], seq.ids[sel], seq.bearings[i][sel])
        out = f.get_snapshot(), f.graph_stats(), f.riccati_int8_slices()
        f.close()
        return out

    g, gstats, sl = run_it({})
    d, dstats, _ = run_it({"EQVIO_GRAPHS": "0"})
    x, _, slx = run_it({"EQVIO_OZAKI": "0"})
    assert sl == 8 and slx == 0 and gstats[0] > 0 and dstats[0] == 0
    assert np.array_equal(g, d)
    (hg, Sg), (hx, Sx) = split_snapshot(g), split_snapshot(x)
    assert int(g[0]) == N
    assert rel(Sg, Sx) < 1e-10 and np.abs(hg - hx).max() < 1e-9, (rel(Sg, Sx), np.abs(hg - hx).max())


@pytest.mark.parametrize("N", [512, 64])
def test_covariance_update_associations_agree(monkeypatch, N):
    """Sigma - K (C Sigma) (what the path evaluates: one product, C Sigma from the S formation) against the reference's evaluation order
    Sigma - (K C) Sigma (VIOFilter.cpp:297; EQVIO_SIGMA_KCS=0), on the int8 path (N = 512) and on DMMA (N = 64): the same matrix up
    to rounding, the lifted state (which does not depend on the covariance update) bit for bit."""
    from eqf_vio_b200.filter import VIOFilter
    from eqf_vio_b200.settings import conditioned_settings
    from eqf_vio_b200.synthetic import period_sequence
    from helpers import feed, split_snapshot

    s = conditioned_settings(outlierThreshold=1e9)
    seq = period_sequence(N, 2, camera_offset=tuple(s.cameraOffset))
    outs = []
    for kcs in ("1", "0"):
        monkeypatch.setenv("EQVIO_SIGMA_KCS", kcs)
        f = VIOFilter(s, device=0)
        for kind, i in seq.events():
            feed(f, seq, kind, i)
        outs.append(f.get_snapshot())
        f.close()
    monkeypatch.delenv("EQVIO_SIGMA_KCS")
    (h1, S1), (h0, S0) = split_snapshot(outs[0]), split_snapshot(outs[1])
    assert not np.array_equal(S1, S0)                   # (two different evaluation orders ran)
    d = np.sqrt(np.abs(np.diag(S0)))
    assert rel(S1, S0) < 1e-12 and (np.abs(S1 - S0) / (d[:, None] * d[None, :])).max() < 1e-10, (rel(S1, S0), (np.abs(S1 - S0) / (d[:, None] * d[None, :])).max())
    assert np.abs(h1 - h0).max() < 1e-10, np.abs(h1 - h0).max()


def test_narrow_lift_back_substitution_matches_wide_form(monkeypatch):
    """bundleLift's R^T = Ym^T Sigma_sub^-1 (EqFMatrices.cpp:239-242) by k_lift_rsolve (the form of capacities > 256, forced here at
    N = 90: five CTAs with their cp.async tile rings handing R over through self-validating entries) against the identity-bordered
    elimination of capacities <= 256, two vision periods."""
    from eqf_vio_b200.filter import VIOFilter
    from eqf_vio_b200.settings import conditioned_settings
    from eqf_vio_b200.synthetic import period_sequence
    from helpers import feed, split_snapshot

    s = conditioned_settings(outlierThreshold=1e9)
    seq = period_sequence(90, 2, camera_offset=tuple(s.cameraOffset))
    outs = []
    for narrow in (True, False):
        if narrow:
            monkeypatch.setenv("EQVIO_LIFT_NARROW", "1")
        f = VIOFilter(s, device=0)
        for kind, i in seq.events():
            feed(f, seq, kind, i)
        assert f.deviceFlags() == 0
        outs.append(f.get_snapshot())
        f.close()
        if narrow:
            monkeypatch.delenv("EQVIO_LIFT_NARROW")
    (h1, S1), (h0, S0) = split_snapshot(outs[0]), split_snapshot(outs[1])
    assert not np.array_equal(h1, h0)                   # (two different forms ran)
    assert np.abs(h1 - h0).max() < 1e-9 and rel(S1, S0) < 1e-11, (np.abs(h1 - h0).max(), rel(S1, S0))
