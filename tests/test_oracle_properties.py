"""The reference's own property tests (eqf_vio/test/*.cpp, googletest) re-stated against the numpy
restatement of the reference (oracle/eqvio_numpy.py).  These are the only pins the reference holds for
this path: finite-difference convergence of A0/B/C0, lift consistency, group axioms,
chart round trips, exp/log.  Same sizes as the reference: N = 5 ids, TEST_REPS = 25, NEAR_ZERO = 1e-12
(test/CMakeLists.txt:30-31)."""
import numpy as np
import pytest
import scipy.linalg

from oracle import eqvio_numpy as onp
from helpers import log_norm, manifold_distance, random_group, random_state, random_unit_quat, state_vec_diff

IDS = [0, 1, 2, 3, 4]
N = len(IDS)
TEST_REPS = 25
NEAR_ZERO = 1e-12
# "monotone non-increasing" in the reference is EXPECT_LE(dist, previousDist); once the truncation error
# is below round-off/dt the FD error floors, so allow the same floor the reference's B test uses (1e-8)
FLOOR = 1e-8


def check_monotone(dists, floor=0.0):
    prev = 1e8
    for d in dists:
        if d > floor:
            assert d <= prev * (1 + 1e-9), (dists,)
        prev = d


# ---- test/test_EqFMatrices.cpp ----
def test_state_matrix_A():  # :28-91
    rng = np.random.default_rng(1)
    X = random_group(rng, IDS)
    xi0 = onp.project_to_manifold(random_state(rng, IDS))
    omega, accel = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
    A0 = onp.state_matrix_A(X, xi0, omega)

    def a0(eps):
        xi_hat = onp.state_group_action(X, xi0)
        xi_e = onp.euclid_coordinate_chart_inv(eps, xi0)
        xi = onp.state_group_action(X, xi_e)
        Lt = onp.lift_velocity(xi, omega, accel) - onp.lift_velocity(xi_hat, omega, accel)
        xi_hat1 = onp.state_group_action(onp.vio_exp(Lt), xi_hat)
        xi_e1 = onp.state_group_action(X.inverse(), xi_hat1)
        return onp.euclid_coordinate_chart(xi_e1, xi0)

    assert np.linalg.norm(a0(np.zeros(5 + 3 * N))) <= NEAR_ZERO
    dirs = [np.eye(5 + 3 * N)[j] for j in range(5 + 3 * N)] + [rng.uniform(-1, 1, 5 + 3 * N) for _ in range(TEST_REPS)]
    for e in dirs:
        comp = A0 @ e
        dists = [np.linalg.norm(a0(10.0**-i * e) / 10.0**-i - comp) for i in range(1, 8)]
        check_monotone(dists, FLOOR)
        assert dists[3] < 1e-2 * max(1.0, np.linalg.norm(comp))  # first-order: error ~ dt


def test_input_matrix_B():  # :93-157
    rng = np.random.default_rng(2)
    X = random_group(rng, IDS)
    xi0 = onp.project_to_manifold(random_state(rng, IDS))
    Bt = onp.input_matrix_B(X, xi0)
    omega, accel = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)

    def b0(v):
        xi_hat = onp.state_group_action(X, xi0)
        Lt = onp.lift_velocity(xi_hat, omega + v[0:3], accel + v[3:6]) - onp.lift_velocity(xi_hat, omega, accel)
        xi_hat1 = onp.state_group_action(onp.vio_exp(Lt), xi_hat)
        xi_e1 = onp.state_group_action(X.inverse(), xi_hat1)
        return onp.euclid_coordinate_chart(xi_e1, xi0)

    assert np.linalg.norm(b0(np.zeros(6))) <= NEAR_ZERO
    dirs = [np.eye(6)[j] for j in range(6)] + [rng.uniform(-1, 1, 6) for _ in range(TEST_REPS)]
    for e in dirs:
        comp = Bt @ e
        dists = [np.linalg.norm(b0(10.0**-i * e) / 10.0**-i - comp) for i in range(1, 6)]
        check_monotone(dists, 1e-8)
        assert dists[3] < 1e-2 * max(1.0, np.linalg.norm(comp))


def test_output_matrix_C():  # :159-217
    rng = np.random.default_rng(3)
    xi0 = onp.project_to_manifold(random_state(rng, IDS))
    C0 = onp.output_matrix_C(xi0)
    y0 = onp.measure_system_state(xi0)

    def c0(eps):
        xi = onp.euclid_coordinate_chart_inv(eps, xi0)
        return onp.output_coordinate_chart(onp.measure_system_state(xi), y0)

    assert np.linalg.norm(c0(np.zeros(5 + 3 * N))) <= NEAR_ZERO
    dirs = [np.eye(5 + 3 * N)[j] for j in range(5 + 3 * N)] + [rng.uniform(-1, 1, 5 + 3 * N) for _ in range(TEST_REPS)]
    for e in dirs:
        comp = C0 @ e
        dists = [np.linalg.norm(c0(10.0**-i * e) / 10.0**-i - comp) for i in range(1, 8)]
        check_monotone(dists, FLOOR)
        assert dists[3] < 1e-2 * max(1.0, np.linalg.norm(comp))


# ---- test/test_VIOLift.cpp ----
def test_lift():  # :28-55
    rng = np.random.default_rng(4)
    for _ in range(TEST_REPS):
        xi0 = random_state(rng, IDS)
        omega, accel = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
        dists = []
        for i in range(8):
            dt = 10.0**-i
            xi1 = onp.integrate_system_function(xi0, omega, accel, dt)
            lam = onp.lift_velocity(onp.project_to_manifold(xi0), omega, accel)
            xi2 = onp.state_group_action(onp.vio_exp(lam * dt), xi0)
            dists.append(np.linalg.norm(state_vec_diff(xi0, xi1) / dt - state_vec_diff(xi0, xi2) / dt))
        check_monotone(dists, FLOOR)


def test_discrete_lift():  # :57-76
    rng = np.random.default_rng(5)
    for _ in range(TEST_REPS):
        Xi0 = random_state(rng, IDS)
        omega, accel = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
        Xi1 = onp.integrate_system_function(Xi0, omega, accel, 0.1)
        xi0, xi1 = onp.project_to_manifold(Xi0), onp.project_to_manifold(Xi1)
        X = onp.lift_velocity_discrete(xi0, omega, accel, 0.1)
        xi2 = onp.state_group_action(X, xi0)
        assert manifold_distance(xi1, xi2) <= NEAR_ZERO


def _lift_converges(lifted, xi0, base, first=0, last=8):
    dists = []
    for i in range(first, last):
        dt = 10.0**-i
        xi1 = onp.state_group_action(onp.vio_exp(lifted * dt), xi0)
        d = np.linalg.norm(onp.euclid_coordinate_chart(xi1, xi0) / dt - base)
        if d < NEAR_ZERO:
            break
        dists.append(d)
    check_monotone(dists, FLOOR)
    return dists


def test_innovation_lift():  # :78-101, :103-130
    rng = np.random.default_rng(6)
    for rep in range(TEST_REPS):
        xi0 = onp.project_to_manifold(random_state(rng, IDS))
        base = rng.uniform(-1, 1, 5 + 3 * N)
        _lift_converges(onp.lift_innovation(base, xi0), xi0, base)
        if rep < 5:
            for j in range(5 + 3 * N):
                e = np.eye(5 + 3 * N)[j]
                _lift_converges(onp.lift_innovation(e, xi0), xi0, e, 1, 5)


def _random_spd(rng, p):
    S = rng.uniform(-1, 1, (p, p))
    return S @ S.T


def test_full_innovation_lift_unit_dirs():  # :132-163
    rng = np.random.default_rng(7)
    for _ in range(5):
        Xi0 = random_state(rng, IDS)
        xi0 = onp.project_to_manifold(Xi0)
        X = random_group(rng, IDS)
        Sigma = _random_spd(rng, 5 + 3 * N)
        for j in range(5 + 3 * N):
            e = np.eye(5 + 3 * N)[j]
            _lift_converges(onp.lift_innovation_wls(e, Xi0, X, Sigma), xi0, e, 1, 8)


def test_total_innovation_lift_and_comparisons():  # :165-252
    rng = np.random.default_rng(8)
    for _ in range(5):
        Xi0 = random_state(rng, IDS)
        xi0 = onp.project_to_manifold(Xi0)
        X = random_group(rng, IDS)
        Sigma = _random_spd(rng, 5 + 3 * N)
        for j in range(5 + 3 * N):
            e = np.eye(5 + 3 * N)[j]
            Gamma = onp.bundle_lift(e, Xi0, X, Sigma)
            lam1 = onp.lift_total_space_innovation(Gamma, Xi0)
            _lift_converges(lam1, xi0, e, 1, 8)
            # the two WLS variants agree (:199-219, 1e-10)
            lam2 = onp.lift_innovation_wls(e, Xi0, X, Sigma)
            assert np.linalg.norm(lam1.U - lam2.U) + np.linalg.norm(lam1.u - lam2.u) + np.linalg.norm(lam1.W - lam2.W) < 1e-8
            # discrete vs exponential lift agree to first order (:221-252)
            dists = []
            for i in range(1, 6):
                dt = 10.0**-i
                G = onp.bundle_lift(dt * e, Xi0, X, Sigma)
                D1 = onp.lift_total_space_innovation_discrete(G, Xi0)
                D2 = onp.vio_exp(onp.lift_total_space_innovation(G, Xi0))
                x1 = onp.state_group_action(D1, xi0)
                x2 = onp.state_group_action(D2, xi0)
                dists.append(manifold_distance(x1, x2) / dt)
            check_monotone(dists, FLOOR)


# ---- test/test_VIOGroup.cpp, test_VIOGroupActions.cpp ----
def test_group_axioms():  # test_VIOGroup.cpp:26
    rng = np.random.default_rng(9)
    for _ in range(TEST_REPS):
        X1, X2, X3 = (random_group(rng, IDS) for _ in range(3))
        I = onp.VIOGroup.identity(IDS)
        assert log_norm(X1 * X1.inverse()) <= 1e-10
        assert log_norm(X1.inverse() * X1) <= 1e-10
        assert log_norm(((X1 * X2) * X3) * (X1 * (X2 * X3)).inverse()) <= 1e-10
        assert log_norm((X1 * I) * X1.inverse()) <= 1e-10


def test_group_actions():  # test_VIOGroupActions.cpp:28-92
    rng = np.random.default_rng(10)
    for _ in range(TEST_REPS):
        X1, X2 = random_group(rng, IDS), random_group(rng, IDS)
        xi0 = random_state(rng, IDS)
        a = onp.state_group_action(X2, onp.state_group_action(X1, xi0))
        b = onp.state_group_action(X1 * X2, xi0)
        assert np.linalg.norm(state_vec_diff(a, b)) <= 1e-10
        y = rng.standard_normal((N, 3))
        y /= np.linalg.norm(y, axis=1, keepdims=True)
        ya = onp.output_group_action(X2, onp.output_group_action(X1, y))
        yb = onp.output_group_action(X1 * X2, y)
        assert np.abs(ya - yb).max() <= 1e-12
        # output equivariance h(phi_X xi) = rho_X h(xi)
        m = onp.project_to_manifold(xi0)
        h1 = onp.measure_system_state(onp.state_group_action(X1, m))
        h2 = onp.output_group_action(X1, onp.measure_system_state(m))
        assert np.abs(h1 - h2).max() <= 1e-12


# ---- test/test_CoordinateCharts.cpp ----
def test_sphere_charts():  # :26-140
    rng = np.random.default_rng(11)
    for _ in range(TEST_REPS):
        eta = rng.standard_normal(3); eta /= np.linalg.norm(eta)
        pole = rng.standard_normal(3); pole /= np.linalg.norm(pole)
        if eta[2] > 0.99 or pole @ eta < -0.99 or pole[2] > 0.99:
            continue
        y = onp.e3_project_sphere(eta)
        assert np.linalg.norm(onp.e3_project_sphere_inv(y) - eta) <= 1e-10
        y = onp.stereo_sphere_chart(eta, pole)
        assert np.linalg.norm(onp.stereo_sphere_chart_inv(y, pole) - eta) <= 1e-10
        assert np.linalg.norm(onp.stereo_sphere_chart(pole, pole)) <= 1e-12
        # differentials vs finite differences
        D = onp.stereo_sphere_chart_diff(eta, pole)
        Di = onp.stereo_sphere_chart_inv_diff(y, pole)
        h = 1e-6
        for k in range(2):
            e = np.eye(2)[k]
            fd = (onp.stereo_sphere_chart_inv(y + h * e, pole) - onp.stereo_sphere_chart_inv(y - h * e, pole)) / (2 * h)
            assert np.linalg.norm(fd - Di[:, k]) < 1e-6
        assert np.linalg.norm(D @ Di - np.eye(2)) < 1e-9


def test_vio_charts():  # :142-220
    rng = np.random.default_rng(12)
    for _ in range(TEST_REPS):
        xi0 = onp.project_to_manifold(random_state(rng, IDS))
        xi1 = onp.project_to_manifold(random_state(rng, IDS))
        xi1.cameraOffset = xi0.cameraOffset
        if xi0.gravityDir @ xi1.gravityDir < -0.9:
            continue
        eps = onp.euclid_coordinate_chart(xi1, xi0)
        xi2 = onp.euclid_coordinate_chart_inv(eps, xi0)
        assert manifold_distance(xi1, xi2) <= 1e-9
        y0 = onp.measure_system_state(xi0)
        y1 = onp.measure_system_state(xi1)
        if np.min(np.sum(y0 * y1, axis=1)) < -0.9:
            continue
        d = onp.output_coordinate_chart(y1, y0)
        assert np.abs(onp.output_coordinate_chart_inv(d, y0) - y1).max() <= 1e-9


# ---- test/test_common.cpp ----
def test_exp_log():  # :27-159
    rng = np.random.default_rng(13)
    for _ in range(TEST_REPS):
        w = rng.uniform(-1, 1, 3)
        R = onp.q_mat(onp.so3_exp(w))
        assert np.linalg.norm(R - scipy.linalg.expm(onp.skew(w))) <= 1e-8
        assert np.linalg.norm(onp.so3_log(onp.so3_exp(w)) - w) <= 1e-8
        u = rng.uniform(-1, 1, 6)
        P = onp.SE3.exp(u)
        U = np.zeros((4, 4)); U[:3, :3] = onp.skew(u[:3]); U[:3, 3] = u[3:]
        E = scipy.linalg.expm(U)
        assert np.linalg.norm(onp.q_mat(P.R) - E[:3, :3]) <= 1e-8 and np.linalg.norm(P.x - E[:3, 3]) <= 1e-8
        assert np.linalg.norm(onp.SE3.log(P) - u) <= 1e-8
    # orthogonality drift over 1000 products
    q = onp.q_identity()
    for _ in range(1000):
        q = onp.q_mul(q, onp.so3_exp(rng.uniform(-1, 1, 3)))
    R = onp.q_mat(q)
    assert np.linalg.norm(R @ R.T - np.eye(3)) < 1e-9
    v1 = rng.standard_normal(3); v2 = rng.standard_normal(3)
    Rv = onp.so3_from_vectors(v1, v2)
    assert np.linalg.norm(onp.q_rot(Rv, v1 / np.linalg.norm(v1)) - v2 / np.linalg.norm(v2)) < 1e-12
    with pytest.raises(onp.SingularChart):
        onp.so3_from_vectors(v1, -v1)
