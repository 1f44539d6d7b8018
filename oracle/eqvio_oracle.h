/*
 * eqvio_oracle.h — CPU restatement (plain C, fp64) of the reference's EqF-VIO filter hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (eqf_vio_b200/, include/) may include,
 * link or call this; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference
 * arm do.  It restates the arithmetic of pvangoor/eqf_vio @ 0b1334ec in the reference's operation
 * order (dense products associated left-to-right as written, explicit LU inverses) with no Eigen.
 *
 * PARITY PIN STATUS: pinned against the reference's own code, not against reference-shipped vectors.
 * The reference ships no golden vectors or known-answer tests for the filter recursion (SURVEY.md §4,
 * §8c) and cannot be built as shipped (Eigen3, yaml-cpp, googletest absent, no network).  Pins:
 *   (1) the reference's UNMODIFIED translation units (every .cpp under eqf_vio/src and libs/core/src) compiled in
 *       place against a minimal Eigen-API stand-in (oracle/refshim/ -> oracle/_ref/libeqvio_ref.so) and
 *       compared with this restatement on identical inputs: free functions to <=1e-13, whole sequences
 *       in every Settings mode to <=1e-12, landmark bookkeeping to identical id sets
 *       (tests/test_oracle_vs_reference.py); the committed golden vectors (tests/golden/<name>.npz) are
 *       recorded from that build.  What this cannot pin is Eigen's internal kernels (the stand-in uses
 *       plain loops, LU with partial pivoting for dynamic inverse(), cofactors for 3x3).
 *   (2) the reference's own property tests (test/test_EqFMatrices.cpp, test_VIOLift.cpp,
 *       test_CoordinateCharts.cpp, test_VIOGroup*.cpp, test_common.cpp) re-stated in
 *       tests/test_oracle_properties.py and passing;
 *   (3) an independent numpy restatement (oracle/eqvio_numpy.py) agreeing with this file.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef EQVIO_ORACLE_H
#define EQVIO_ORACLE_H

#include <stddef.h>
#include "../include/eqvio.h" /* POD settings + snapshot layout shared with the C ABI */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct eqo_filter eqo_filter;

/* ---- filter (eqf_vio/src/VIOFilter.cpp) ---- */
eqo_filter* eqo_create(const eqvio_settings_t* s);
void eqo_destroy(eqo_filter* f);
int eqo_process_imu(eqo_filter* f, double stamp, const double omega[3], const double accel[3]);
int eqo_process_vision(eqo_filter* f, double stamp, int n, const int* ids, const double* bearings);
int eqo_set_inertial_points(eqo_filter* f, int n, const int* ids, const double* points);
double eqo_get_time(const eqo_filter* f);
int eqo_get_num_landmarks(const eqo_filter* f);
int eqo_get_state(const eqo_filter* f, double pose[7], double velocity[3], double cam_offset[7], int* n, int cap,
                  int* ids, double* landmarks);
int eqo_get_covariance(const eqo_filter* f, double* dst, int ld);
int eqo_get_bias(const eqo_filter* f, double bias[6]);
size_t eqo_snapshot_size(int n_landmarks);
int eqo_get_snapshot(const eqo_filter* f, double* dst, size_t cap);
int eqo_set_snapshot(eqo_filter* f, const double* src, size_t len);
/* number of worker threads used by the dense products (1 = the reference's single-threaded Eigen) */
void eqo_set_threads(int nthreads);

/* ---- pieces, on the filter's current (xi0, X, Sigma) ---- */
/* A0 (p x p), Bt (p x 6), col-major, p = 5+3N  (EqFMatrices.cpp:277-317, 346-382) */
int eqo_state_matrix_A(const eqo_filter* f, const double omega[3], double* A0);
int eqo_input_matrix_B(const eqo_filter* f, double* Bt);
/* C0 (2N x p) (EqFMatrices.cpp:319-344) */
int eqo_output_matrix_C(const eqo_filter* f, double* C0);
/* F = I + T*A_b (n x n), B_b (n x 6)  (VIOFilter.cpp:177-185) */
int eqo_build_FB(const eqo_filter* f, double T, const double omega[3], double* F, double* Bb);
/* Sigma <- T (P + B_b R B_b^T) + F Sigma F^T  (VIOFilter.cpp:162-189) */
int eqo_riccati_propagate(eqo_filter* f, double T, const double omega[3]);
/* C = [0, C0] (m x n) and delta (m)  (VIOFilter.cpp:264-273) */
int eqo_build_C_delta(const eqo_filter* f, const double* bearings, double* C, double* delta);
/* S, K, gamma, Sigma <- Sigma - K C Sigma  (VIOFilter.cpp:276-279,297) */
int eqo_gain_update(eqo_filter* f, const double* bearings, double* K, double* gamma);
/* bundleLift (EqFMatrices.cpp:173-252) with Sigma[6:,6:] of the filter */
int eqo_bundle_lift(const eqo_filter* f, const double* gamma_eqf, double* Gamma);
/* liftInnovation(gamma, xi0, X, Sigma) WLS variant (EqFMatrices.cpp:98-171): out = U(6), u(3), W(4N) */
int eqo_lift_innovation_wls(const eqo_filter* f, const double* gamma_eqf, double* alg);
/* liftInnovation(gamma, xi0) (EqFMatrices.cpp:35-67): out = U(6), u(3), W(4N) */
int eqo_lift_innovation(const eqo_filter* f, const double* gamma_eqf, double* alg);

/* sphere charts (VIOState.cpp:230-240); nonzero return = singular */
int eqo_stereo_sphere_chart(double* y2, const double* eta3, const double* pole3);
int eqo_stereo_sphere_chart_inv(double* eta3, const double* y2, const double* pole3);

/* ---- dense helpers (stand-ins for Eigen's GEMM / PartialPivLU inverse) ---- */
void eqo_dgemm(int transA, int transB, int M, int N, int K, double alpha, const double* A, int lda, const double* B,
               int ldb, double beta, double* C, int ldc);
int eqo_inverse(int n, const double* A, int lda, double* Ainv, int ldi);

#ifdef __cplusplus
}
#endif
#endif
