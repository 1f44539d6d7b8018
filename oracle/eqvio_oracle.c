/*
 * eqvio_oracle.c — CPU restatement (plain C99, fp64, no Eigen) of the reference EqF-VIO filter.
 * TEST INFRASTRUCTURE ONLY — see eqvio_oracle.h for the rules and the parity-pin status.
 *
 * Reference paths are relative to /root/reference (pvangoor/eqf_vio @ 0b1334ec).
 * Matrices are column-major (Eigen's default).  3x3 matrices are double[9] column-major too:
 * M(r,c) = m[r + 3*c].
 */
#include "eqvio_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * small fixed-size helpers
 * ---------------------------------------------------------------------------------------- */
typedef struct { double w, x, y, z; } quat;            /* Eigen::Quaterniond, SO3.cpp:28 */
typedef struct { quat R; double x[3]; } se3;           /* SE3.cpp:25-27 */
typedef struct { quat R; double a; } sot3;             /* SOT3.cpp:25-27 */

#define M3(m, r, c) ((m)[(r) + 3 * (c)])

static void v3_set(double* o, double a, double b, double c) { o[0] = a; o[1] = b; o[2] = c; }
static void v3_copy(double* o, const double* a) { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
static double v3_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double v3_norm(const double* a) { return sqrt(v3_dot(a, a)); }
static void v3_cross(double* o, const double* a, const double* b) {
    double t0 = a[1] * b[2] - a[2] * b[1], t1 = a[2] * b[0] - a[0] * b[2], t2 = a[0] * b[1] - a[1] * b[0];
    o[0] = t0; o[1] = t1; o[2] = t2;
}
static void v3_normalized(double* o, const double* a) {
    double n = v3_norm(a);
    o[0] = a[0] / n; o[1] = a[1] / n; o[2] = a[2] / n;
}
static void m3_identity(double* m) { memset(m, 0, 9 * sizeof(double)); m[0] = m[4] = m[8] = 1.0; }
static void m3_mul(double* o, const double* a, const double* b) {
    double t[9];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += M3(a, r, k) * M3(b, k, c);
            M3(t, r, c) = s;
        }
    memcpy(o, t, sizeof t);
}
static void m3_transpose(double* o, const double* a) {
    double t[9];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) M3(t, r, c) = M3(a, c, r);
    memcpy(o, t, sizeof t);
}
static void m3_vec(double* o, const double* m, const double* v) {
    double t[3];
    for (int r = 0; r < 3; ++r) t[r] = M3(m, r, 0) * v[0] + M3(m, r, 1) * v[1] + M3(m, r, 2) * v[2];
    v3_copy(o, t);
}
static void m3_scale(double* o, const double* a, double s) { for (int i = 0; i < 9; ++i) o[i] = a[i] * s; }
/* SO3::skew, libs/core/src/SO3.cpp:110-114 */
static void skew(double* m, const double* v) {
    M3(m, 0, 0) = 0;     M3(m, 0, 1) = -v[2]; M3(m, 0, 2) = v[1];
    M3(m, 1, 0) = v[2];  M3(m, 1, 1) = 0;     M3(m, 1, 2) = -v[0];
    M3(m, 2, 0) = -v[1]; M3(m, 2, 1) = v[0];  M3(m, 2, 2) = 0;
}
/* general 3x3 inverse by cofactors (Eigen's fixed-size inverse, used at EqFMatrices.cpp:310) */
static void m3_inverse(double* o, const double* a) {
    double c00 = M3(a, 1, 1) * M3(a, 2, 2) - M3(a, 1, 2) * M3(a, 2, 1);
    double c10 = M3(a, 1, 2) * M3(a, 2, 0) - M3(a, 1, 0) * M3(a, 2, 2);
    double c20 = M3(a, 1, 0) * M3(a, 2, 1) - M3(a, 1, 1) * M3(a, 2, 0);
    double det = M3(a, 0, 0) * c00 + M3(a, 0, 1) * c10 + M3(a, 0, 2) * c20;
    double id = 1.0 / det;
    double t[9];
    M3(t, 0, 0) = c00 * id;
    M3(t, 1, 0) = c10 * id;
    M3(t, 2, 0) = c20 * id;
    M3(t, 0, 1) = (M3(a, 0, 2) * M3(a, 2, 1) - M3(a, 0, 1) * M3(a, 2, 2)) * id;
    M3(t, 1, 1) = (M3(a, 0, 0) * M3(a, 2, 2) - M3(a, 0, 2) * M3(a, 2, 0)) * id;
    M3(t, 2, 1) = (M3(a, 0, 1) * M3(a, 2, 0) - M3(a, 0, 0) * M3(a, 2, 1)) * id;
    M3(t, 0, 2) = (M3(a, 0, 1) * M3(a, 1, 2) - M3(a, 0, 2) * M3(a, 1, 1)) * id;
    M3(t, 1, 2) = (M3(a, 0, 2) * M3(a, 1, 0) - M3(a, 0, 0) * M3(a, 1, 2)) * id;
    M3(t, 2, 2) = (M3(a, 0, 0) * M3(a, 1, 1) - M3(a, 0, 1) * M3(a, 1, 0)) * id;
    memcpy(o, t, sizeof t);
}

/* ---- quaternion-backed SO3 (libs/core/src/SO3.cpp; Eigen::Quaterniond semantics) ---- */
static quat q_identity(void) { quat q = {1, 0, 0, 0}; return q; }
/* Eigen quaternion product, no renormalisation (SO3.cpp:66-70) */
static quat q_mul(quat a, quat b) {
    quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
/* Eigen Quaternion::inverse(): conjugate / squaredNorm (SO3.cpp:74) */
static quat q_inverse(quat a) {
    double n2 = a.w * a.w + a.x * a.x + a.y * a.y + a.z * a.z;
    quat r = {a.w / n2, -a.x / n2, -a.y / n2, -a.z / n2};
    return r;
}
/* Eigen Quaternion::_transformVector: v + w*uv + u x uv with uv = 2 u x v (SO3.cpp:58) */
static void q_rotate(double* o, quat q, const double* v) {
    double u[3] = {q.x, q.y, q.z}, uv[3], uuv[3];
    v3_cross(uv, u, v);
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    v3_cross(uuv, u, uv);
    double t[3] = {v[0] + q.w * uv[0] + uuv[0], v[1] + q.w * uv[1] + uuv[1], v[2] + q.w * uv[2] + uuv[2]};
    v3_copy(o, t);
}
/* SO3::applyInverse: quat.inverse() * point (SO3.cpp:108) */
static void q_rotate_inv(double* o, quat q, const double* v) { q_rotate(o, q_inverse(q), v); }
/* Eigen Quaternion::toRotationMatrix (SO3.cpp:92) */
static void q_to_mat(double* m, quat q) {
    double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    M3(m, 0, 0) = 1 - (tyy + tzz); M3(m, 0, 1) = txy - twz;       M3(m, 0, 2) = txz + twy;
    M3(m, 1, 0) = txy + twz;       M3(m, 1, 1) = 1 - (txx + tzz); M3(m, 1, 2) = tyz - twx;
    M3(m, 2, 0) = txz - twy;       M3(m, 2, 1) = tyz + twx;       M3(m, 2, 2) = 1 - (txx + tyy);
}
/* Eigen Quaterniond(Matrix3d) — Shepperd's method as in Eigen's quaternionbase_assign_impl (SO3.cpp:100) */
static quat q_from_mat(const double* m) {
    quat q;
    double t = M3(m, 0, 0) + M3(m, 1, 1) + M3(m, 2, 2);
    if (t > 0) {
        t = sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (M3(m, 2, 1) - M3(m, 1, 2)) * t;
        q.y = (M3(m, 0, 2) - M3(m, 2, 0)) * t;
        q.z = (M3(m, 1, 0) - M3(m, 0, 1)) * t;
    } else {
        int i = 0;
        if (M3(m, 1, 1) > M3(m, 0, 0)) i = 1;
        if (M3(m, 2, 2) > M3(m, i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        double v[3];
        t = sqrt(M3(m, i, i) - M3(m, j, j) - M3(m, k, k) + 1.0);
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (M3(m, k, j) - M3(m, j, k)) * t;
        v[j] = (M3(m, j, i) + M3(m, i, j)) * t;
        v[k] = (M3(m, k, i) + M3(m, i, k)) * t;
        q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
    return q;
}
/* SO3::SO3Exp, SO3.cpp:122-140 */
static quat so3_exp(const double* w) {
    double th = v3_norm(w), A, B;
    if (fabs(th) >= 1e-8) { A = sin(th) / th; B = (1 - cos(th)) / (th * th); }
    else { A = 1.0; B = 0.5; }
    double wx[9], wx2[9], R[9];
    skew(wx, w);
    m3_mul(wx2, wx, wx);
    m3_identity(R);
    for (int i = 0; i < 9; ++i) R[i] += A * wx[i] + B * wx2[i];
    return q_from_mat(R);
}
/* SO3::SO3FromVectors, SO3.cpp:155-167.  Returns 0 on success, nonzero if the vectors oppose. */
static int so3_from_vectors(quat* out, const double* origin, const double* dest) {
    double a[3], b[3], v[3];
    v3_normalized(a, origin);
    v3_normalized(b, dest);
    v3_cross(v, a, b);
    double c = v3_dot(a, b);
    double sv[9], sv2[9], mat[9];
    skew(sv, v);
    m3_mul(sv2, sv, sv);
    m3_identity(mat);
    for (int i = 0; i < 9; ++i) mat[i] += sv[i] + 1 / (1 + c) * sv2[i];
    if (fabs(1 + c) <= 1e-8) return 1;
    *out = q_from_mat(mat);
    return 0;
}

/* ---- SE3 (libs/core/src/SE3.cpp) ---- */
static se3 se3_identity(void) { se3 P; P.R = q_identity(); v3_set(P.x, 0, 0, 0); return P; }
/* SE3 * point, SE3.cpp:65 */
static void se3_apply(double* o, const se3* P, const double* p) {
    double t[3];
    q_rotate(t, P->R, p);
    o[0] = t[0] + P->x[0]; o[1] = t[1] + P->x[1]; o[2] = t[2] + P->x[2];
}
/* SE3 product, SE3.cpp:73-78 */
static se3 se3_mul(const se3* a, const se3* b) {
    se3 r;
    r.R = q_mul(a->R, b->R);
    double t[3];
    q_rotate(t, a->R, b->x);
    r.x[0] = a->x[0] + t[0]; r.x[1] = a->x[1] + t[1]; r.x[2] = a->x[2] + t[2];
    return r;
}
/* SE3::invert, SE3.cpp:82-85: R.invert(); x = -(R * x) */
static se3 se3_inverse(const se3* a) {
    se3 r;
    r.R = q_inverse(a->R);
    double t[3];
    q_rotate(t, r.R, a->x);
    v3_set(r.x, -t[0], -t[1], -t[2]);
    return r;
}
/* SE3::Adjoint, SE3.cpp:95-103; 6x6 col-major */
static void se3_adjoint(double* Ad, const se3* P) {
    double R[9], sx[9], sxR[9];
    q_to_mat(R, P->R);
    skew(sx, P->x);
    m3_mul(sxR, sx, R);
    memset(Ad, 0, 36 * sizeof(double));
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            Ad[r + 6 * c] = M3(R, r, c);
            Ad[(r + 3) + 6 * c] = M3(sxR, r, c);
            Ad[(r + 3) + 6 * (c + 3)] = M3(R, r, c);
        }
}
/* SE3::SE3Exp, SE3.cpp:139-164; u = (omega, v) */
static se3 se3_exp(const double* u) {
    const double* w = u;
    const double* v = u + 3;
    double th = v3_norm(w), A, B, C;
    if (fabs(th) >= 1e-12) { A = sin(th) / th; B = (1 - cos(th)) / (th * th); C = (1 - A) / (th * th); }
    else { A = 1.0; B = 0.5; C = 1.0 / 6.0; }
    double wx[9], wx2[9], R[9], V[9];
    skew(wx, w);
    m3_mul(wx2, wx, wx);
    m3_identity(R);
    m3_identity(V);
    for (int i = 0; i < 9; ++i) { R[i] += A * wx[i] + B * wx2[i]; V[i] += B * wx[i] + C * wx2[i]; }
    se3 P;
    P.R = q_from_mat(R);
    m3_vec(P.x, V, v);
    return P;
}

/* ---- SOT3 (libs/core/src/SOT3.cpp) ---- */
static sot3 sot3_identity(void) { sot3 Q; Q.R = q_identity(); Q.a = 1.0; return Q; }
static sot3 sot3_mul(sot3 a, sot3 b) { sot3 r; r.R = q_mul(a.R, b.R); r.a = a.a * b.a; return r; }          /* :69-74 */
static sot3 sot3_inverse(sot3 a) { sot3 r; r.R = q_inverse(a.R); r.a = 1.0 / a.a; return r; }               /* :78-81 */
static void sot3_apply(double* o, sot3 Q, const double* p) {                                                 /* :62 */
    double t[3]; q_rotate(t, Q.R, p); v3_set(o, Q.a * t[0], Q.a * t[1], Q.a * t[2]);
}
static void sot3_as_matrix3(double* m, sot3 Q) { double R[9]; q_to_mat(R, Q.R); m3_scale(m, R, Q.a); }      /* :107-110 */
/* SOT3::SOT3Exp, SOT3.cpp:127-132 */
static sot3 sot3_exp(const double* w4) { sot3 r; r.R = so3_exp(w4); r.a = exp(w4[3]); return r; }

/* ---- sphere charts (eqf_vio/src/VIOState.cpp:199-251) ---- */
/* e3ProjectSphere :199-204 */
static void e3_project_sphere(double* y, const double* eta) {
    double d = 1 - eta[2];
    y[0] = eta[0] / d; y[1] = eta[1] / d;
}
/* e3ProjectSphereInv :206-211 */
static void e3_project_sphere_inv(double* eta, const double* y) {
    double s = 2.0 / (y[0] * y[0] + y[1] * y[1] + 1);
    eta[0] = s * y[0]; eta[1] = s * y[1]; eta[2] = 1 + s * (0 - 1);
}
/* e3ProjectSphereDiff :213-220 ; 2x3 col-major (D[r + 2c]) */
static void e3_project_sphere_diff(double* D, const double* eta) {
    double omz = 1 - eta[2];
    double full[9];
    m3_identity(full);
    for (int i = 0; i < 9; ++i) full[i] *= omz;
    /* + (eta - e3) e3^T : third column */
    M3(full, 0, 2) += eta[0]; M3(full, 1, 2) += eta[1]; M3(full, 2, 2) += eta[2] - 1;
    double s = pow(1 - eta[2], -2.0);
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 2; ++r) D[r + 2 * c] = s * M3(full, r, c);
}
/* e3ProjectSphereInvDiff :222-228 ; 3x2 col-major (D[r + 3c]) */
static void e3_project_sphere_inv_diff(double* D, const double* y) {
    double n2 = y[0] * y[0] + y[1] * y[1];
    double t[6];
    t[0 + 3 * 0] = (n2 + 1.0) - 2 * y[0] * y[0];
    t[1 + 3 * 0] = -2 * y[1] * y[0];
    t[0 + 3 * 1] = -2 * y[0] * y[1];
    t[1 + 3 * 1] = (n2 + 1.0) - 2 * y[1] * y[1];
    t[2 + 3 * 0] = 2 * y[0];
    t[2 + 3 * 1] = 2 * y[1];
    double s = 2.0 * pow(n2 + 1.0, -2.0);
    for (int i = 0; i < 6; ++i) D[i] = s * t[i];
}
static int sphere_rot(quat* q, const double* pole) {
    double mp[3] = {-pole[0], -pole[1], -pole[2]}, e3[3] = {0, 0, 1};
    return so3_from_vectors(q, mp, e3);
}
/* stereoSphereChart :230-234 */
static int stereo_sphere_chart(double* y, const double* eta, const double* pole);
int eqo_stereo_sphere_chart(double* y, const double* eta, const double* pole) { return stereo_sphere_chart(y, eta, pole); }
static int stereo_sphere_chart(double* y, const double* eta, const double* pole) {
    quat q; if (sphere_rot(&q, pole)) return 1;
    double er[3]; q_rotate(er, q, eta);
    e3_project_sphere(y, er);
    return 0;
}
/* stereoSphereChartInv :236-240 (exported for the chart round-trip tests) */
int eqo_stereo_sphere_chart_inv(double* eta, const double* y, const double* pole) {
    double er[3]; e3_project_sphere_inv(er, y);
    quat q; if (sphere_rot(&q, pole)) return 1;
    q_rotate(eta, q_inverse(q), er);
    return 0;
}
/* stereoSphereChartDiff :242-246 ; 2x3 */
static int stereo_sphere_chart_diff(double* D, const double* eta, const double* pole) {
    quat q; if (sphere_rot(&q, pole)) return 1;
    double er[3]; q_rotate(er, q, eta);
    double d23[6], R[9];
    e3_project_sphere_diff(d23, er);
    q_to_mat(R, q);
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 2; ++r) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += d23[r + 2 * k] * M3(R, k, c);
            D[r + 2 * c] = s;
        }
    return 0;
}
/* stereoSphereChartInvDiff :248-251 ; 3x2 */
static int stereo_sphere_chart_inv_diff(double* D, const double* y, const double* pole) {
    quat q; if (sphere_rot(&q, pole)) return 1;
    double Rinv[9], d32[6];
    q_to_mat(Rinv, q_inverse(q));
    e3_project_sphere_inv_diff(d32, y);
    for (int c = 0; c < 2; ++c)
        for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += M3(Rinv, r, k) * d32[k + 3 * c];
            D[r + 3 * c] = s;
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * dense helpers: stand-ins for Eigen's dynamic GEMM, PartialPivLU inverse, 4x4 Householder QR
 * ---------------------------------------------------------------------------------------- */
void eqo_set_threads(int n) { (void)n; /* single-threaded; kept for ABI stability */ }

/* C <- alpha op(A) op(B) + beta C, column-major.  Cache-blocked, inner loop is an axpy down a
 * column of C so gcc vectorises it.  Single-threaded like the reference's Eigen build
 * (no -fopenmp in eqf_vio/CMakeLists.txt). */
void eqo_dgemm(int transA, int transB, int M, int N, int K, double alpha, const double* A, int lda, const double* B,
               int ldb, double beta, double* C, int ldc) {
    /* materialise op(A) as M x K col-major if transposed: keeps the kernel simple */
    double* At = NULL;
    if (transA) {
        At = (double*)malloc(sizeof(double) * (size_t)M * K);
        for (int k = 0; k < K; ++k)
            for (int m = 0; m < M; ++m) At[m + (size_t)M * k] = A[k + (size_t)lda * m];
        A = At; lda = M;
    }
    const int KB = 256, MB = 512;
    for (int n = 0; n < N; ++n) {
        double* c = C + (size_t)ldc * n;
        if (beta == 0.0) for (int m = 0; m < M; ++m) c[m] = 0.0;
        else if (beta != 1.0) for (int m = 0; m < M; ++m) c[m] *= beta;
        for (int k0 = 0; k0 < K; k0 += KB) {
            int k1 = k0 + KB < K ? k0 + KB : K;
            for (int m0 = 0; m0 < M; m0 += MB) {
                int m1 = m0 + MB < M ? m0 + MB : M;
                for (int k = k0; k < k1; ++k) {
                    double b = alpha * (transB ? B[n + (size_t)ldb * k] : B[k + (size_t)ldb * n]);
                    if (b == 0.0) continue; /* exact: adding 0*a changes nothing for finite a */
                    const double* a = A + (size_t)lda * k;
                    for (int m = m0; m < m1; ++m) c[m] += a[m] * b;
                }
            }
        }
    }
    free(At);
}

/* Full inverse by LU with partial pivoting (what Eigen's MatrixXd::inverse() does for dynamic sizes:
 * PartialPivLU then solve against the identity; VIOFilter.cpp:277, EqFMatrices.cpp:239). */
int eqo_inverse(int n, const double* A, int lda, double* Ainv, int ldi) {
    double* LU = (double*)malloc(sizeof(double) * (size_t)n * n);
    int* piv = (int*)malloc(sizeof(int) * n);
    for (int c = 0; c < n; ++c) memcpy(LU + (size_t)n * c, A + (size_t)lda * c, sizeof(double) * n);
    int status = 0;
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(LU[k + (size_t)n * k]);
        for (int r = k + 1; r < n; ++r) {
            double v = fabs(LU[r + (size_t)n * k]);
            if (v > best) { best = v; p = r; }
        }
        piv[k] = p;
        if (best == 0.0) { status = 1; continue; }
        if (p != k)
            for (int c = 0; c < n; ++c) {
                double t = LU[k + (size_t)n * c]; LU[k + (size_t)n * c] = LU[p + (size_t)n * c]; LU[p + (size_t)n * c] = t;
            }
        double inv = 1.0 / LU[k + (size_t)n * k];
        for (int r = k + 1; r < n; ++r) LU[r + (size_t)n * k] *= inv;
        for (int c = k + 1; c < n; ++c) {
            double f = LU[k + (size_t)n * c];
            if (f == 0.0) continue;
            double* col = LU + (size_t)n * c;
            const double* l = LU + (size_t)n * k;
            for (int r = k + 1; r < n; ++r) col[r] -= l[r] * f;
        }
    }
    /* solve L U X = P I, column by column */
    for (int c = 0; c < n; ++c) {
        double* x = Ainv + (size_t)ldi * c;
        for (int r = 0; r < n; ++r) x[r] = 0.0;
        x[c] = 1.0;
        for (int k = 0; k < n; ++k) /* apply the row interchanges to e_c */
            if (piv[k] != k) { double t = x[k]; x[k] = x[piv[k]]; x[piv[k]] = t; }
        for (int k = 0; k < n; ++k) { /* forward, unit lower */
            double f = x[k];
            if (f == 0.0) continue;
            const double* l = LU + (size_t)n * k;
            for (int r = k + 1; r < n; ++r) x[r] -= l[r] * f;
        }
        for (int k = n - 1; k >= 0; --k) { /* backward */
            x[k] /= LU[k + (size_t)n * k];
            double f = x[k];
            const double* u = LU + (size_t)n * k;
            for (int r = 0; r < k; ++r) x[r] -= u[r] * f;
        }
    }
    free(LU);
    free(piv);
    return status;
}

/* 4x4 Householder QR solve (EqFMatrices.cpp:240-242 `.householderQr().solve`).  A col-major 4x4. */
static void qr_solve4(const double* Ain, const double* bin, double* x) {
    double A[16], b[4];
    memcpy(A, Ain, sizeof A);
    memcpy(b, bin, sizeof b);
    for (int k = 0; k < 4; ++k) {
        double tail2 = 0;
        for (int r = k + 1; r < 4; ++r) tail2 += A[r + 4 * k] * A[r + 4 * k];
        double c0 = A[k + 4 * k];
        if (tail2 == 0.0) continue; /* already triangular in this column (tau = 0) */
        double beta = sqrt(c0 * c0 + tail2);
        if (c0 >= 0) beta = -beta;
        double v[4] = {0, 0, 0, 0};
        for (int r = k + 1; r < 4; ++r) v[r] = A[r + 4 * k] / (c0 - beta);
        v[k] = 1.0;
        double tau = (beta - c0) / beta;
        for (int c = k; c < 4; ++c) {
            double s = 0;
            for (int r = k; r < 4; ++r) s += v[r] * A[r + 4 * c];
            s *= tau;
            for (int r = k; r < 4; ++r) A[r + 4 * c] -= s * v[r];
        }
        double s = 0;
        for (int r = k; r < 4; ++r) s += v[r] * b[r];
        s *= tau;
        for (int r = k; r < 4; ++r) b[r] -= s * v[r];
    }
    for (int k = 3; k >= 0; --k) {
        double s = b[k];
        for (int c = k + 1; c < 4; ++c) s -= A[k + 4 * c] * x[c];
        x[k] = s / A[k + 4 * k];
    }
}

/* ------------------------------------------------------------------------------------------
 * filter state (eqf_vio/include/eqf_vio/VIOFilter.h:43-55)
 * ---------------------------------------------------------------------------------------- */
struct eqo_filter {
    eqvio_settings_t s;
    double inputBias[6];
    /* xi0 : VIOState (VIOState.h:51-60) */
    se3 pose0;
    double vel0[3];
    se3 camOffset;
    /* X : VIOGroup (VIOGroup.h:24-33) */
    se3 XA;
    double Xw[3];
    int N, cap;
    int* id;
    double* q0; /* 3 per landmark: xi0.bodyLandmarks[i].p */
    sot3* Q;
    double* Sigma; /* n x n col-major, ld = n */
    int initialised;
    double currentTime;
    double curOmega[3], curAccel[3]; /* currentVelocity */
    double accOmega[3], accAccel[3]; /* accumulatedVelocity */
    double accTime;
};

#define NS(f) (EQVIO_SIGMA_BASE_SIZE + 3 * (f)->N)

static void ensure_cap(eqo_filter* f, int N) {
    if (N <= f->cap) return;
    int cap = f->cap ? f->cap : 16;
    while (cap < N) cap *= 2;
    f->id = (int*)realloc(f->id, sizeof(int) * cap);
    f->q0 = (double*)realloc(f->q0, sizeof(double) * 3 * cap);
    f->Q = (sot3*)realloc(f->Q, sizeof(sot3) * cap);
    f->cap = cap;
}

static void se3_from_pose7(se3* P, const double* p) { /* x y z qw qx qy qz */
    v3_set(P->x, p[0], p[1], p[2]);
    P->R.w = p[3]; P->R.x = p[4]; P->R.y = p[5]; P->R.z = p[6];
}
static void se3_to_pose7(double* p, const se3* P) {
    p[0] = P->x[0]; p[1] = P->x[1]; p[2] = P->x[2];
    p[3] = P->R.w; p[4] = P->R.x; p[5] = P->R.y; p[6] = P->R.z;
}

/* VIOFilter::VIOFilter(const Settings&), VIOFilter.cpp:60-73 (+ member defaults VIOFilter.h:46-55) */
eqo_filter* eqo_create(const eqvio_settings_t* s) {
    eqo_filter* f = (eqo_filter*)calloc(1, sizeof *f);
    f->s = *s;
    f->Sigma = (double*)calloc(121, sizeof(double));
    for (int i = 0; i < 11; ++i) f->Sigma[i + 11 * i] = 1.0;
    for (int i = 0; i < 3; ++i) {
        f->Sigma[i + 11 * i] = s->initialBiasOmegaVariance;
        f->Sigma[(3 + i) + 11 * (3 + i)] = s->initialBiasAccelVariance;
        f->Sigma[(8 + i) + 11 * (8 + i)] = s->initialVelocityVariance;
    }
    for (int i = 0; i < 2; ++i) f->Sigma[(6 + i) + 11 * (6 + i)] = s->initialGravityVariance;
    f->pose0 = se3_identity();
    se3_from_pose7(&f->camOffset, s->cameraOffset);
    f->XA = se3_identity();
    for (int i = 0; i < 3; ++i) { f->inputBias[i] = s->initialOmegaBias[i]; f->inputBias[3 + i] = s->initialAccelBias[i]; }
    f->currentTime = -1;
    return f;
}
void eqo_destroy(eqo_filter* f) {
    if (!f) return;
    free(f->id); free(f->q0); free(f->Q); free(f->Sigma); free(f);
}

/* derived quantities of xi_hat = stateGroupAction(X, xi0) (VIOGroup.cpp:23-69) */
typedef struct {
    se3 pose;        /* P0 * A */
    double eta0[3];  /* projectToManifold(xi0).gravityDir = R_P0^-1 e3, VIOState.cpp:90 */
    double eta[3];   /* R_A^-1 eta0 */
    double vel[3];   /* R_A^-1 (v0 - w) */
} hat_base;
static void compute_hat_base(const eqo_filter* f, hat_base* h) {
    h->pose = se3_mul(&f->pose0, &f->XA);
    double e3[3] = {0, 0, 1};
    q_rotate_inv(h->eta0, f->pose0.R, e3);
    q_rotate_inv(h->eta, f->XA.R, h->eta0);
    double d[3] = {f->vel0[0] - f->Xw[0], f->vel0[1] - f->Xw[1], f->vel0[2] - f->Xw[2]};
    q_rotate_inv(h->vel, f->XA.R, d);
}
/* qhat_i = Q_i^-1 * q0_i  (VIOGroup.cpp:39-44; SOT3::inverse then operator*) */
static void qhat_i(const eqo_filter* f, int i, double* out) { sot3_apply(out, sot3_inverse(f->Q[i]), f->q0 + 3 * i); }

/* ---- EqF matrices (eqf_vio/src/EqFMatrices.cpp) ---- */
/* EqFStateMatrixA_euclid_impl, EqFMatrices.cpp:277-317.  A0 is p x p col-major. */
int eqo_state_matrix_A(const eqo_filter* f, const double omega[3], double* A0) {
    const int N = f->N, p = 5 + 3 * N;
    memset(A0, 0, sizeof(double) * (size_t)p * p);
    hat_base h;
    compute_hat_base(f, &h);
    /* :289 */
    double y0[2] = {0, 0}, D32[6];
    if (stereo_sphere_chart_inv_diff(D32, y0, h.eta0)) return EQVIO_ERR_SINGULAR_CHART;
    for (int c = 0; c < 2; ++c)
        for (int r = 0; r < 3; ++r) A0[(2 + r) + (size_t)p * c] = -D32[r + 3 * c] * EQVIO_GRAVITY_CONSTANT;
    /* :292-297 */
    double R_IC[9], R_A[9], R_ICt[9], R_At[9], RR[9];
    q_to_mat(R_IC, f->camOffset.R);
    q_to_mat(R_A, f->XA.R);
    m3_transpose(R_ICt, R_IC);
    m3_transpose(R_At, R_A);
    for (int i = 0; i < N; ++i) {
        double Qhat[9], t[9];
        sot3_as_matrix3(Qhat, f->Q[i]); /* X.Q[i].R().asMatrix() * X.Q[i].a() */
        m3_mul(t, Qhat, R_ICt);
        m3_mul(RR, t, R_At);
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) A0[(5 + 3 * i + r) + (size_t)p * (2 + c)] = -M3(RR, r, c);
    }
    /* :302-304: U_C = Ad(T_IC^-1) (omega; v_hat) */
    se3 Tinv = se3_inverse(&f->camOffset);
    double Ad[36], UI[6] = {omega[0], omega[1], omega[2], h.vel[0], h.vel[1], h.vel[2]}, vC[3];
    se3_adjoint(Ad, &Tinv);
    for (int r = 0; r < 3; ++r) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += Ad[(3 + r) + 6 * k] * UI[k];
        vC[r] = s;
    }
    /* :305-312 */
    for (int i = 0; i < N; ++i) {
        double Qhat[9], Qinv[9], qh[3], sq[9], sv[9], inner[9], t[9], Aq[9];
        sot3_as_matrix3(Qhat, f->Q[i]);
        qhat_i(f, i, qh);
        skew(sq, qh);
        skew(sv, vC);
        m3_mul(inner, sq, sv);
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) M3(inner, r, c) += -2 * vC[r] * qh[c] + qh[r] * vC[c];
        m3_inverse(Qinv, Qhat);
        m3_mul(t, Qhat, inner);
        m3_mul(Aq, t, Qinv);
        double sc = 1 / v3_dot(qh, qh);
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) A0[(5 + 3 * i + r) + (size_t)p * (5 + 3 * i + c)] = -M3(Aq, r, c) * sc;
    }
    return EQVIO_OK;
}

/* EqFInputMatrixB_euclid_impl, EqFMatrices.cpp:346-382.  Bt is p x 6 col-major. */
int eqo_input_matrix_B(const eqo_filter* f, double* Bt) {
    const int N = f->N, p = 5 + 3 * N;
    memset(Bt, 0, sizeof(double) * (size_t)p * 6);
    hat_base h;
    compute_hat_base(f, &h);
    double R_A[9], D23[6], se[9], t[9];
    q_to_mat(R_A, f->XA.R);
    /* :364 */
    if (stereo_sphere_chart_diff(D23, h.eta0, h.eta0)) return EQVIO_ERR_SINGULAR_CHART;
    skew(se, h.eta);
    m3_mul(t, R_A, se);
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 2; ++r) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += D23[r + 2 * k] * M3(t, k, c);
            Bt[r + (size_t)p * c] = s;
        }
    /* :367-368 */
    double sv[9];
    skew(sv, h.vel);
    m3_mul(t, R_A, sv);
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            Bt[(2 + r) + (size_t)p * c] = M3(t, r, c);
            Bt[(2 + r) + (size_t)p * (3 + c)] = M3(R_A, r, c);
        }
    /* :371-377 */
    double RT_IC[9], sx[9], RTsx[9];
    q_to_mat(RT_IC, q_inverse(f->camOffset.R));
    skew(sx, f->camOffset.x);
    m3_mul(RTsx, RT_IC, sx);
    for (int i = 0; i < N; ++i) {
        double Qhat[9], qh[3], sq[9], inner[9], blk[9];
        sot3_as_matrix3(Qhat, f->Q[i]);
        qhat_i(f, i, qh);
        skew(sq, qh);
        m3_mul(inner, sq, RT_IC);
        for (int k = 0; k < 9; ++k) inner[k] += RTsx[k];
        m3_mul(blk, Qhat, inner);
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) Bt[(5 + 3 * i + r) + (size_t)p * c] = M3(blk, r, c);
    }
    return EQVIO_OK;
}

/* EqFOutputMatrixC_euclid_impl, EqFMatrices.cpp:319-344.  C0 is 2N x p col-major. */
int eqo_output_matrix_C(const eqo_filter* f, double* C0) {
    const int N = f->N, p = 5 + 3 * N, m = 2 * N;
    memset(C0, 0, sizeof(double) * (size_t)m * p);
    for (int i = 0; i < N; ++i) {
        const double* qi0 = f->q0 + 3 * i;
        double yi0[3], D23[6], proj[9];
        v3_normalized(yi0, qi0);
        if (stereo_sphere_chart_diff(D23, yi0, yi0)) return EQVIO_ERR_SINGULAR_CHART;
        m3_identity(proj);
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) M3(proj, r, c) -= yi0[r] * yi0[c];
        double sc = 1 / v3_norm(qi0);
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 2; ++r) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += (sc * D23[r + 2 * k]) * M3(proj, k, c);
                C0[(2 * i + r) + (size_t)m * (5 + 3 * i + c)] = s;
            }
    }
    return EQVIO_OK;
}

/* Biased assembly, VIOFilter.cpp:177-185: F = I + A_b T, B_b = [0; Bt] */
int eqo_build_FB(const eqo_filter* f, double T, const double omega[3], double* F, double* Bb) {
    const int N = f->N, p = 5 + 3 * N, n = NS(f);
    double* A0 = (double*)malloc(sizeof(double) * (size_t)p * p);
    double* Bt = (double*)malloc(sizeof(double) * (size_t)p * 6);
    int st = eqo_state_matrix_A(f, omega, A0);
    if (!st) st = eqo_input_matrix_B(f, Bt);
    if (!st) {
        memset(F, 0, sizeof(double) * (size_t)n * n);
        memset(Bb, 0, sizeof(double) * (size_t)n * 6);
        for (int c = 0; c < p; ++c)
            for (int r = 0; r < p; ++r) F[(6 + r) + (size_t)n * (6 + c)] = A0[r + (size_t)p * c] * T;
        for (int c = 0; c < 6; ++c)
            for (int r = 0; r < p; ++r) {
                F[(6 + r) + (size_t)n * c] = -Bt[r + (size_t)p * c] * T;
                Bb[(6 + r) + (size_t)n * c] = Bt[r + (size_t)p * c];
            }
        for (int i = 0; i < n; ++i) F[i + (size_t)n * i] += 1.0;
    }
    free(A0);
    free(Bt);
    return st;
}

/* Riccati step, VIOFilter.cpp:162-189 */
int eqo_riccati_propagate(eqo_filter* f, double T, const double omega[3]) {
    const int N = f->N, n = NS(f);
    double* F = (double*)malloc(sizeof(double) * (size_t)n * n);
    double* Bb = (double*)malloc(sizeof(double) * (size_t)n * 6);
    int st = eqo_build_FB(f, T, omega, F, Bb);
    if (st) { free(F); free(Bb); return st; }
    double* W = (double*)malloc(sizeof(double) * (size_t)n * n);
    double* S2 = (double*)malloc(sizeof(double) * (size_t)n * n);
    /* (F * Sigma) * F^T */
    eqo_dgemm(0, 0, n, n, n, 1.0, F, n, f->Sigma, n, 0.0, W, n);
    eqo_dgemm(0, 1, n, n, n, 1.0, W, n, F, n, 0.0, S2, n);
    /* T * (PMat + B_b R B_b^T) */
    double Rd[6] = {f->s.velOmegaVariance, f->s.velOmegaVariance, f->s.velOmegaVariance,
                    f->s.velAccelVariance, f->s.velAccelVariance, f->s.velAccelVariance};
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < n; ++r) {
            double brb = 0;
            for (int k = 0; k < 6; ++k) brb += (Bb[r + (size_t)n * k] * Rd[k]) * Bb[c + (size_t)n * k];
            double pm = 0;
            if (r == c) {
                if (r < 3) pm = f->s.biasOmegaProcessVariance;
                else if (r < 6) pm = f->s.biasAccelProcessVariance;
                else if (r < 8) pm = f->s.gravityProcessVariance;
                else if (r < 11) pm = f->s.velocityProcessVariance;
                else pm = f->s.pointProcessVariance;
            }
            f->Sigma[r + (size_t)n * c] = T * (pm + brb) + S2[r + (size_t)n * c];
        }
    (void)N;
    free(F); free(Bb); free(W); free(S2);
    return EQVIO_OK;
}

/* ---- velocity lift and state propagate (eqf_vio/src/VIOGroup.cpp) ---- */
/* liftVelocityDiscrete, VIOGroup.cpp:209-243; X <- X * lift, VIOGroup.cpp:92-110 */
static int propagate_state_discrete(eqo_filter* f, const double* omega, const double* accel, double dt) {
    hat_base h;
    compute_hat_base(f, &h);
    double AVel[6] = {omega[0], omega[1], omega[2], h.vel[0], h.vel[1], h.vel[2]}, u[6];
    for (int i = 0; i < 6; ++i) u[i] = dt * AVel[i];
    se3 LA = se3_exp(u);
    /* lift.w = v - R_LA (v + dt(-omega x v + accel - eta g)) */
    double so[9], sov[3], inner[3], rot[3], Lw[3];
    skew(so, omega);
    m3_vec(sov, so, h.vel);
    for (int i = 0; i < 3; ++i)
        inner[i] = h.vel[i] + dt * (-sov[i] + accel[i] - h.eta[i] * EQVIO_GRAVITY_CONSTANT);
    q_rotate(rot, LA.R, inner);
    for (int i = 0; i < 3; ++i) Lw[i] = h.vel[i] - rot[i];
    /* camera-frame velocity and its inverse pose change */
    se3 Tinv = se3_inverse(&f->camOffset);
    double Ad[36], UC[6], mUC[6];
    se3_adjoint(Ad, &Tinv);
    for (int r = 0; r < 6; ++r) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += Ad[r + 6 * k] * AVel[k];
        UC[r] = s;
    }
    for (int i = 0; i < 6; ++i) mUC[i] = -dt * UC[i];
    se3 camInv = se3_exp(mUC);
    for (int i = 0; i < f->N; ++i) {
        double p0[3], p1[3], n1[3], n0[3];
        qhat_i(f, i, p0);
        se3_apply(p1, &camInv, p0);
        v3_normalized(n1, p1);
        v3_normalized(n0, p0);
        sot3 L;
        if (so3_from_vectors(&L.R, n1, n0)) return EQVIO_ERR_SINGULAR_CHART;
        L.a = v3_norm(p0) / v3_norm(p1);
        f->Q[i] = sot3_mul(f->Q[i], L);
    }
    /* X = X * lift: w <- w + R_A * Lw; A <- A * LA */
    double t[3];
    q_rotate(t, f->XA.R, Lw);
    for (int i = 0; i < 3; ++i) f->Xw[i] += t[i];
    f->XA = se3_mul(&f->XA, &LA);
    return EQVIO_OK;
}
/* liftVelocity (VIOGroup.cpp:178-207) then X <- X * VIOExp(dt * lift) (VIOGroup.cpp:245-256) */
static int propagate_state_continuous(eqo_filter* f, const double* omega, const double* accel, double dt) {
    hat_base h;
    compute_hat_base(f, &h);
    double U[6] = {omega[0], omega[1], omega[2], h.vel[0], h.vel[1], h.vel[2]}, u[3];
    for (int i = 0; i < 3; ++i) u[i] = -accel[i] + h.eta[i] * EQVIO_GRAVITY_CONSTANT;
    se3 Tinv = se3_inverse(&f->camOffset);
    double Ad[36], UC[6];
    se3_adjoint(Ad, &Tinv);
    for (int r = 0; r < 6; ++r) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += Ad[r + 6 * k] * U[k];
        UC[r] = s;
    }
    for (int i = 0; i < f->N; ++i) {
        double p[3], sp[9], spv[3], W[4];
        qhat_i(f, i, p);
        double n2 = v3_dot(p, p);
        skew(sp, p);
        m3_vec(spv, sp, UC + 3);
        for (int k = 0; k < 3; ++k) W[k] = dt * (UC[k] + spv[k] / n2);
        W[3] = dt * (v3_dot(p, UC + 3) / n2);
        f->Q[i] = sot3_mul(f->Q[i], sot3_exp(W));
    }
    double dU[6], t[3];
    for (int i = 0; i < 6; ++i) dU[i] = dt * U[i];
    se3 EA = se3_exp(dU);
    double du[3] = {dt * u[0], dt * u[1], dt * u[2]};
    q_rotate(t, f->XA.R, du);
    for (int i = 0; i < 3; ++i) f->Xw[i] += t[i];
    f->XA = se3_mul(&f->XA, &EA);
    return EQVIO_OK;
}

/* VIOFilter::integrateUpToTime, VIOFilter.cpp:146-209.  Returns 1 if integrated, 0 if skipped, <0 on error. */
static int integrate_up_to_time(eqo_filter* f, double newTime, int doRiccati) {
    if (f->currentTime < 0) return 0;
    double dt = newTime - f->currentTime;
    if (dt <= 0) return 0;
    f->accTime += dt;
    for (int i = 0; i < 3; ++i) { f->accOmega[i] += f->curOmega[i] * dt; f->accAccel[i] += f->curAccel[i] * dt; }
    if (doRiccati) {
        double inv = 1.0 / f->accTime;
        double om[3] = {f->accOmega[0] * inv, f->accOmega[1] * inv, f->accOmega[2] * inv};
        int st = eqo_riccati_propagate(f, f->accTime, om);
        if (st) return st;
        for (int i = 0; i < 3; ++i) f->accOmega[i] = f->accAccel[i] = 0;
        f->accTime = 0.0;
    }
    int st = f->s.useDiscreteVelocityLift ? propagate_state_discrete(f, f->curOmega, f->curAccel, dt)
                                          : propagate_state_continuous(f, f->curOmega, f->curAccel, dt);
    if (st) return st;
    f->currentTime = newTime;
    return 1;
}

/* VIOFilter::processIMUData, VIOFilter.cpp:120-131 (+ initialiseFromIMUData :133-144) */
int eqo_process_imu(eqo_filter* f, double stamp, const double omega[3], const double accel[3]) {
    double uo[3], ua[3];
    for (int i = 0; i < 3; ++i) { uo[i] = omega[i] - f->inputBias[i]; ua[i] = accel[i] - f->inputBias[3 + i]; }
    if (!f->initialised) {
        f->pose0 = se3_identity();
        v3_set(f->vel0, 0, 0, 0);
        f->initialised = 1;
        double g[3], e3[3] = {0, 0, 1};
        v3_normalized(g, ua);
        if (so3_from_vectors(&f->pose0.R, g, e3)) return EQVIO_ERR_SINGULAR_CHART;
    }
    int r = integrate_up_to_time(f, stamp, !f->s.fastRiccati);
    if (r < 0) return r;
    v3_copy(f->curOmega, uo);
    v3_copy(f->curAccel, ua);
    f->currentTime = stamp;
    return r == 1 ? EQVIO_OK : EQVIO_SKIPPED_DT;
}

/* removeRows/removeCols + removeLandmarkAtIndex, VIOFilter.cpp:29-47, 421-427 */
static void remove_landmark_at(eqo_filter* f, int idx) {
    int n = NS(f), s0 = EQVIO_SIGMA_BASE_SIZE + 3 * idx, nn = n - 3;
    double* S = (double*)malloc(sizeof(double) * (size_t)nn * nn);
    for (int c = 0, cc = 0; c < n; ++c) {
        if (c >= s0 && c < s0 + 3) continue;
        for (int r = 0, rr = 0; r < n; ++r) {
            if (r >= s0 && r < s0 + 3) continue;
            S[rr + (size_t)nn * cc] = f->Sigma[r + (size_t)n * c];
            ++rr;
        }
        ++cc;
    }
    free(f->Sigma);
    f->Sigma = S;
    for (int i = idx; i < f->N - 1; ++i) {
        f->id[i] = f->id[i + 1];
        f->Q[i] = f->Q[i + 1];
        v3_copy(f->q0 + 3 * i, f->q0 + 3 * (i + 1));
    }
    f->N -= 1;
}

static int cmp_double(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* C = [0, C0], delta; VIOFilter.cpp:264-273, VIOGroup.cpp:71-90, VisionMeasurement.cpp:24-34, VIOState.cpp:58-70 */
int eqo_build_C_delta(const eqo_filter* f, const double* bearings, double* C, double* delta) {
    const int N = f->N, m = 2 * N, n = NS(f), p = 5 + 3 * N;
    if (delta) {
        for (int i = 0; i < N; ++i) {
            double y0[3], ye[3];
            v3_normalized(y0, f->q0 + 3 * i);                       /* measureSystemState(xi0) */
            q_rotate(ye, q_inverse(q_inverse(f->Q[i].R)), bearings + 3 * i); /* X.inverse().Q.R().inverse() * y */
            if (stereo_sphere_chart(delta + 2 * i, ye, y0)) return EQVIO_ERR_SINGULAR_CHART;
        }
    }
    if (C) {
        double* C0 = (double*)malloc(sizeof(double) * (size_t)m * p);
        int st = eqo_output_matrix_C(f, C0);
        if (st) { free(C0); return st; }
        memset(C, 0, sizeof(double) * (size_t)m * n);
        memcpy(C + (size_t)m * 6, C0, sizeof(double) * (size_t)m * p);
        free(C0);
    }
    return EQVIO_OK;
}

/* S, K, gamma and Sigma update; VIOFilter.cpp:269-279, 297.  K (n x m) and gamma (n) are outputs. */
static int gain_core(eqo_filter* f, const double* C, const double* delta, double* K, double* gamma, int update_sigma) {
    const int N = f->N, m = 2 * N, n = NS(f);
    double* CS = (double*)malloc(sizeof(double) * (size_t)m * n);
    double* S = (double*)malloc(sizeof(double) * (size_t)m * m);
    double* Sinv = (double*)malloc(sizeof(double) * (size_t)m * m);
    double* SCt = (double*)malloc(sizeof(double) * (size_t)n * m);
    eqo_dgemm(0, 0, m, n, n, 1.0, C, m, f->Sigma, n, 0.0, CS, m);      /* C * Sigma */
    eqo_dgemm(0, 1, m, m, n, 1.0, CS, m, C, m, 0.0, S, m);             /* (C Sigma) * C^T */
    for (int i = 0; i < m; ++i) S[i + (size_t)m * i] += f->s.measurementVariance;
    int st = eqo_inverse(m, S, m, Sinv, m);
    eqo_dgemm(0, 1, n, m, n, 1.0, f->Sigma, n, C, m, 0.0, SCt, n);     /* Sigma * C^T */
    eqo_dgemm(0, 0, n, m, m, 1.0, SCt, n, Sinv, m, 0.0, K, n);         /* (Sigma C^T) * S^-1 */
    for (int r = 0; r < n; ++r) {
        double s = 0;
        for (int c = 0; c < m; ++c) s += K[r + (size_t)n * c] * delta[c];
        gamma[r] = s;
    }
    if (update_sigma) {
        double* KC = (double*)malloc(sizeof(double) * (size_t)n * n);
        double* KCS = (double*)malloc(sizeof(double) * (size_t)n * n);
        eqo_dgemm(0, 0, n, n, m, 1.0, K, n, C, m, 0.0, KC, n);         /* K * C */
        eqo_dgemm(0, 0, n, n, n, 1.0, KC, n, f->Sigma, n, 0.0, KCS, n);/* (K C) * Sigma */
        for (size_t i = 0; i < (size_t)n * n; ++i) f->Sigma[i] -= KCS[i];
        free(KC); free(KCS);
    }
    free(CS); free(S); free(Sinv); free(SCt);
    return st ? EQVIO_ERR_NOT_SPD : EQVIO_OK;
}

int eqo_gain_update(eqo_filter* f, const double* bearings, double* K, double* gamma) {
    const int N = f->N, m = 2 * N, n = NS(f);
    double* C = (double*)malloc(sizeof(double) * (size_t)m * n);
    double* delta = (double*)malloc(sizeof(double) * m);
    double* Kl = K ? K : (double*)malloc(sizeof(double) * (size_t)n * m);
    double* gl = gamma ? gamma : (double*)malloc(sizeof(double) * n);
    int st = eqo_build_C_delta(f, bearings, C, delta);
    if (!st) st = gain_core(f, C, delta, Kl, gl, 1);
    if (!K) free(Kl);
    if (!gamma) free(gl);
    free(C); free(delta);
    return st;
}

/* liftInnovation(baseInnovation, xi0 manifold), EqFMatrices.cpp:35-67.  alg = U(6), u(3), W(4N) */
int eqo_lift_innovation(const eqo_filter* f, const double* g, double* alg) {
    hat_base h;
    compute_hat_base(f, &h);
    double y0[2] = {0, 0}, D32[6], se[9], t[3], Om[3];
    if (stereo_sphere_chart_inv_diff(D32, y0, h.eta0)) return EQVIO_ERR_SINGULAR_CHART;
    for (int r = 0; r < 3; ++r) t[r] = D32[r] * g[0] + D32[r + 3] * g[1];
    skew(se, h.eta0);
    m3_vec(Om, se, t);
    for (int r = 0; r < 3; ++r) { alg[r] = -Om[r]; alg[3 + r] = 0; }
    double sO[9], sOv[3];
    skew(sO, alg);
    m3_vec(sOv, sO, f->vel0);
    for (int r = 0; r < 3; ++r) alg[6 + r] = -g[2 + r] - sOv[r];
    for (int i = 0; i < f->N; ++i) {
        const double* q = f->q0 + 3 * i;
        const double* gq = g + 5 + 3 * i;
        double c[3], n2 = v3_dot(q, q);
        v3_cross(c, q, gq);
        for (int r = 0; r < 3; ++r) alg[9 + 4 * i + r] = -c[r] / n2;
        alg[9 + 4 * i + 3] = -v3_dot(q, gq) / n2;
    }
    return EQVIO_OK;
}

/* shared WLS core of bundleLift (EqFMatrices.cpp:173-252) and liftInnovation/4-arg (:98-171):
 * given DeltaU (6) with the default Omega part, returns DeltaU <- KPerp DeltaU + KPara x */
static int wls_core(const eqo_filter* f, const double* g, double* DeltaU) {
    const int N = f->N, p = 5 + 3 * N, l = 3 * N, n = NS(f);
    hat_base h;
    compute_hat_base(f, &h);
    double eta0[3];
    v3_normalized(eta0, h.eta0);
    /* KPara (6x4), KPerp (6x6) :195-206 */
    double KPara[24], KPerp[36];
    memset(KPara, 0, sizeof KPara);
    memset(KPerp, 0, sizeof KPerp);
    for (int r = 0; r < 3; ++r) KPara[r + 6 * 0] = eta0[r];
    for (int r = 0; r < 3; ++r) KPara[(3 + r) + 6 * (1 + r)] = 1.0;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) KPerp[r + 6 * c] = (r == c ? 1.0 : 0.0) - eta0[r] * eta0[c];
    /* R_C = R_Phat * R_IC, its inverse as a matrix :209-210 */
    quat RC = q_mul(h.pose.R, f->camOffset.R);
    double RCtm[9];
    q_to_mat(RCtm, q_inverse(RC));
    double AdP0[36];
    se3_adjoint(AdP0, &f->pose0);
    double DUF[6];
    for (int r = 0; r < 6; ++r) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += KPerp[r + 6 * k] * DeltaU[k];
        DUF[r] = s;
    }
    double* M = (double*)malloc(sizeof(double) * (size_t)l * 4);   /* coeffMat, l x 4 col-major */
    double* obs = (double*)malloc(sizeof(double) * l);
    double* D = (double*)calloc((size_t)p * l, sizeof(double));     /* weightingTransferD, p x l */
    se3 PT = se3_mul(&h.pose, &f->camOffset);
    for (int i = 0; i < N; ++i) {
        const double* gq = g + 5 + 3 * i;
        double qh[3], pH[3], tq[3], alpha[3];
        qhat_i(f, i, qh);
        se3_apply(pH, &PT, qh); /* xiHat.pose * xiHat.cameraOffset * p (SE3 product first, left-assoc) */
        sot3_apply(tq, sot3_inverse(f->Q[i]), gq);
        q_rotate(alpha, RC, tq);
        for (int r = 0; r < 3; ++r) alpha[r] = -alpha[r];
        /* pHatMat (3x6) = [-skew(pHat), I] ; pHatMat * AdP0 (3x6) */
        double pm[18], sp[9], pmAd[18];
        skew(sp, pH);
        memset(pm, 0, sizeof pm);
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) { pm[r + 3 * c] = -M3(sp, r, c); pm[r + 3 * (3 + c)] = (r == c) ? 1.0 : 0.0; }
        for (int c = 0; c < 6; ++c)
            for (int r = 0; r < 3; ++r) {
                double s = 0;
                for (int k = 0; k < 6; ++k) s += pm[r + 3 * k] * AdP0[k + 6 * c];
                pmAd[r + 3 * c] = s;
            }
        for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int k = 0; k < 6; ++k) s += pmAd[r + 3 * k] * DUF[k];
            obs[3 * i + r] = alpha[r] - s;
        }
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 3; ++r) {
                double s = 0;
                for (int k = 0; k < 6; ++k) s += pmAd[r + 3 * k] * KPara[k + 6 * c];
                M[(3 * i + r) + (size_t)l * c] = s;
            }
        double Qm[9], blk[9];
        sot3_as_matrix3(Qm, f->Q[i]);
        m3_mul(blk, Qm, RCtm);
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) D[(5 + 3 * i + r) + (size_t)p * (3 * i + c)] = M3(blk, r, c);
    }
    /* weightMat = D^T * Sigma_sub^-1 * D  :239 */
    double* Ssub = (double*)malloc(sizeof(double) * (size_t)p * p);
    double* Sinv = (double*)malloc(sizeof(double) * (size_t)p * p);
    for (int c = 0; c < p; ++c)
        memcpy(Ssub + (size_t)p * c, f->Sigma + 6 + (size_t)n * (6 + c), sizeof(double) * p);
    int st = eqo_inverse(p, Ssub, p, Sinv, p);
    double* DtSi = (double*)malloc(sizeof(double) * (size_t)l * p);
    double* Wm = (double*)malloc(sizeof(double) * (size_t)l * l);
    eqo_dgemm(1, 0, l, p, p, 1.0, D, p, Sinv, p, 0.0, DtSi, l);
    eqo_dgemm(0, 0, l, l, p, 1.0, DtSi, l, D, p, 0.0, Wm, l);
    /* (M^T W M) x = M^T W obs, 4x4 Householder QR :240-242 */
    double* MtW = (double*)malloc(sizeof(double) * (size_t)4 * l);
    eqo_dgemm(1, 0, 4, l, l, 1.0, M, l, Wm, l, 0.0, MtW, 4);
    double G[16], rhs[4], x[4];
    eqo_dgemm(0, 0, 4, 4, l, 1.0, MtW, 4, M, l, 0.0, G, 4);
    eqo_dgemm(0, 0, 4, 1, l, 1.0, MtW, 4, obs, l, 0.0, rhs, 4);
    qr_solve4(G, rhs, x);
    for (int r = 0; r < 6; ++r) {
        double s = 0;
        for (int k = 0; k < 4; ++k) s += KPara[r + 6 * k] * x[k];
        DeltaU[r] = DUF[r] + s;
    }
    free(M); free(obs); free(D); free(Ssub); free(Sinv); free(DtSi); free(Wm); free(MtW);
    return st ? EQVIO_ERR_NOT_SPD : EQVIO_OK;
}

/* bundleLift, EqFMatrices.cpp:173-252.  Gamma has 9+3N entries. */
int eqo_bundle_lift(const eqo_filter* f, const double* g, double* Gamma) {
    const int N = f->N;
    hat_base h;
    compute_hat_base(f, &h);
    double eta0[3], y0[2] = {0, 0}, D32[6], se[9], t[3], Om[3], DU[6];
    v3_normalized(eta0, h.eta0);
    if (stereo_sphere_chart_inv_diff(D32, y0, eta0)) return EQVIO_ERR_SINGULAR_CHART;
    for (int r = 0; r < 3; ++r) t[r] = D32[r] * g[0] + D32[r + 3] * g[1];
    skew(se, eta0);
    m3_vec(Om, se, t);
    for (int r = 0; r < 3; ++r) { DU[r] = -Om[r]; DU[3 + r] = 0; }
    int st = wls_core(f, g, DU);
    if (st) return st;
    for (int r = 0; r < 6; ++r) Gamma[r] = DU[r];
    for (int r = 0; r < 3 + 3 * N; ++r) Gamma[6 + r] = g[2 + r];
    return EQVIO_OK;
}

/* liftInnovation(gamma, xi0, X, Sigma), EqFMatrices.cpp:98-171 */
int eqo_lift_innovation_wls(const eqo_filter* f, const double* g, double* alg) {
    int st = eqo_lift_innovation(f, g, alg);
    if (st) return st;
    double DU[6];
    memcpy(DU, alg, sizeof DU);
    st = wls_core(f, g, DU);
    if (st) return st;
    memcpy(alg, DU, sizeof DU);
    double sO[9], sOv[3];
    skew(sO, alg);
    m3_vec(sOv, sO, f->vel0);
    for (int r = 0; r < 3; ++r) alg[6 + r] = -g[2 + r] - sOv[r];
    return EQVIO_OK;
}

/* Delta * X for a group element given as (DA, Dw, DQ[]), VIOGroup.cpp:92-110 */
static void left_multiply(eqo_filter* f, const se3* DA, const double* Dw, const sot3* DQ) {
    double t[3];
    q_rotate(t, DA->R, f->Xw);
    for (int i = 0; i < 3; ++i) f->Xw[i] = Dw[i] + t[i];
    f->XA = se3_mul(DA, &f->XA);
    for (int i = 0; i < f->N; ++i) f->Q[i] = sot3_mul(DQ[i], f->Q[i]);
}

/* VIOFilter::processVisionData, VIOFilter.cpp:232-302 */
int eqo_process_vision(eqo_filter* f, double stamp, int nmeas, const int* ids, const double* bearings) {
    int r = integrate_up_to_time(f, stamp, 1);
    if (r < 0) return r;
    if (r == 0) return EQVIO_SKIPPED_DT;
    if (!f->initialised) return EQVIO_NOT_INITIALISED;
    for (int i = 1; i < nmeas; ++i)
        if (ids[i] < ids[i - 1]) return EQVIO_ERR_UNSORTED;

    /* removeOldLandmarks :393-419 — descending index order */
    for (int li = f->N - 1; li >= 0; --li) {
        int found = 0;
        for (int j = 0; j < nmeas; ++j)
            if (ids[j] == f->id[li]) { found = 1; break; }
        if (!found) remove_landmark_at(f, li);
    }
    /* matchMeasurementsToState :211-230 */
    int* mid = (int*)malloc(sizeof(int) * (nmeas > 0 ? nmeas : 1));
    double* my = (double*)malloc(sizeof(double) * 3 * (nmeas > 0 ? nmeas : 1));
    int newPos = f->N - 1;
    for (int j = 0; j < nmeas; ++j) {
        int idx = -1;
        for (int i = 0; i < f->N; ++i)
            if (f->id[i] == ids[j]) { idx = i; break; }
        if (idx < 0) idx = ++newPos;
        mid[idx] = ids[j];
        v3_copy(my + 3 * idx, bearings + 3 * j);
    }
    int M = nmeas;
    /* removeOutliers :429-443 — yHat is computed once, before any removal */
    {
        int N0 = f->N;
        double* yh = (double*)malloc(sizeof(double) * 3 * (N0 > 0 ? N0 : 1));
        for (int i = 0; i < N0; ++i) {
            double qh[3];
            qhat_i(f, i, qh);
            v3_normalized(yh + 3 * i, qh);
        }
        for (int i = N0 - 1; i >= 0; --i) {
            double d[3] = {my[3 * i] - yh[3 * i], my[3 * i + 1] - yh[3 * i + 1], my[3 * i + 2] - yh[3 * i + 2]};
            if (v3_norm(d) > f->s.outlierThreshold) {
                remove_landmark_at(f, i);
                for (int j = i; j < M - 1; ++j) { mid[j] = mid[j + 1]; v3_copy(my + 3 * j, my + 3 * (j + 1)); }
                --M;
            }
        }
        free(yh);
    }
    /* addNewLandmarks :345-391 */
    if (M > f->N) {
        int oldN = f->N, newN = M - oldN;
        double median = f->s.initialSceneDepth;
        if (oldN > 0) {
            double* d2 = (double*)malloc(sizeof(double) * oldN);
            for (int i = 0; i < oldN; ++i) { double qh[3]; qhat_i(f, i, qh); d2[i] = v3_dot(qh, qh); }
            qsort(d2, oldN, sizeof(double), cmp_double); /* nth_element(size/2) == sorted[size/2] */
            median = pow(d2[oldN / 2], 0.5);
            free(d2);
        }
        ensure_cap(f, M);
        int n0 = NS(f), n1 = n0 + 3 * newN;
        double* S = (double*)calloc((size_t)n1 * n1, sizeof(double));
        for (int c = 0; c < n0; ++c) memcpy(S + (size_t)n1 * c, f->Sigma + (size_t)n0 * c, sizeof(double) * n0);
        for (int i = n0; i < n1; ++i) S[i + (size_t)n1 * i] = f->s.initialPointVariance;
        free(f->Sigma);
        f->Sigma = S;
        for (int j = 0; j < newN; ++j) {
            int i = oldN + j;
            f->id[i] = mid[i];
            for (int k = 0; k < 3; ++k) f->q0[3 * i + k] = my[3 * i + k] * median;
            f->Q[i] = sot3_identity();
        }
        f->N = M;
    }
    if (M == 0) { free(mid); free(my); return EQVIO_EMPTY_MEASUREMENT; }

    const int N = f->N, m = 2 * N, n = NS(f), p = 5 + 3 * N;
    double* C = (double*)malloc(sizeof(double) * (size_t)m * n);
    double* delta = (double*)malloc(sizeof(double) * m);
    double* K = (double*)malloc(sizeof(double) * (size_t)n * m);
    double* gamma = (double*)malloc(sizeof(double) * n);
    int st = eqo_build_C_delta(f, my, C, delta);
    if (!st) st = gain_core(f, C, delta, K, gamma, 0);
    se3 DA;
    double Dw[3];
    sot3* DQ = (sot3*)malloc(sizeof(sot3) * N);
    if (!st) {
        const double* ge = gamma + 6;
        if (f->s.useInnovationLift) {
            double* Gamma = (double*)malloc(sizeof(double) * (9 + 3 * N));
            st = eqo_bundle_lift(f, ge, Gamma); /* prior Sigma block, :285 */
            if (!st && f->s.useDiscreteInnovationLift) {
                /* liftTotalSpaceInnovationDiscrete, EqFMatrices.cpp:254-275 */
                DA = se3_exp(Gamma);
                double t[3] = {f->vel0[0] + Gamma[6], f->vel0[1] + Gamma[7], f->vel0[2] + Gamma[8]}, rt[3];
                q_rotate(rt, DA.R, t);
                for (int k = 0; k < 3; ++k) Dw[k] = f->vel0[k] - rt[k];
                for (int i = 0; i < N && !st; ++i) {
                    const double* qi = f->q0 + 3 * i;
                    double q1[3] = {qi[0] + Gamma[9 + 3 * i], qi[1] + Gamma[10 + 3 * i], qi[2] + Gamma[11 + 3 * i]};
                    double n1[3], n0[3];
                    v3_normalized(n1, q1);
                    v3_normalized(n0, qi);
                    if (so3_from_vectors(&DQ[i].R, n1, n0)) st = EQVIO_ERR_SINGULAR_CHART;
                    DQ[i].a = v3_norm(qi) / v3_norm(q1);
                }
            } else if (!st) {
                /* VIOExp(liftTotalSpaceInnovation(Gamma, xi0)), EqFMatrices.cpp:69-96, VIOGroup.cpp:245-256 */
                DA = se3_exp(Gamma);
                double sO[9], sOv[3];
                skew(sO, Gamma);
                m3_vec(sOv, sO, f->vel0);
                for (int k = 0; k < 3; ++k) Dw[k] = -Gamma[6 + k] - sOv[k];
                for (int i = 0; i < N; ++i) {
                    const double* qi = f->q0 + 3 * i;
                    const double* gq = Gamma + 9 + 3 * i;
                    double c[3], W[4], n2 = v3_dot(qi, qi);
                    v3_cross(c, qi, gq);
                    for (int k = 0; k < 3; ++k) W[k] = -c[k] / n2;
                    W[3] = -v3_dot(qi, gq) / n2;
                    DQ[i] = sot3_exp(W);
                }
            }
            free(Gamma);
        } else {
            /* VIOExp(liftInnovation(gamma_eqf, xi0)), VIOFilter.cpp:292 */
            double* alg = (double*)malloc(sizeof(double) * (9 + 4 * N));
            st = eqo_lift_innovation(f, ge, alg);
            if (!st) {
                DA = se3_exp(alg);
                for (int k = 0; k < 3; ++k) Dw[k] = alg[6 + k];
                for (int i = 0; i < N; ++i) DQ[i] = sot3_exp(alg + 9 + 4 * i);
            }
            free(alg);
        }
    }
    if (!st) {
        for (int k = 0; k < 6; ++k) f->inputBias[k] += gamma[k];    /* :295 */
        left_multiply(f, &DA, Dw, DQ);                                /* :296 */
        /* Sigma <- Sigma - (K C) Sigma  :297 */
        double* KC = (double*)malloc(sizeof(double) * (size_t)n * n);
        double* KCS = (double*)malloc(sizeof(double) * (size_t)n * n);
        eqo_dgemm(0, 0, n, n, m, 1.0, K, n, C, m, 0.0, KC, n);
        eqo_dgemm(0, 0, n, n, n, 1.0, KC, n, f->Sigma, n, 0.0, KCS, n);
        for (size_t i = 0; i < (size_t)n * n; ++i) f->Sigma[i] -= KCS[i];
        free(KC); free(KCS);
    }
    (void)p;
    free(C); free(delta); free(K); free(gamma); free(DQ); free(mid); free(my);
    return st;
}

/* VIOFilter::setInertialPoints, VIOFilter.cpp:93-118 */
int eqo_set_inertial_points(eqo_filter* f, int n, const int* ids, const double* points) {
    ensure_cap(f, n);
    se3 PT = se3_mul(&f->pose0, &f->camOffset);
    se3 inv = se3_inverse(&PT);
    for (int i = 0; i < n; ++i) {
        f->id[i] = ids[i];
        f->Q[i] = sot3_identity();
        se3_apply(f->q0 + 3 * i, &inv, points + 3 * i);
    }
    int nn = EQVIO_SIGMA_BASE_SIZE + 3 * n, n0 = NS(f);
    double* S = (double*)calloc((size_t)nn * nn, sizeof(double));
    for (int i = 0; i < nn; ++i) S[i + (size_t)nn * i] = f->s.initialPointVariance;
    for (int c = 0; c < 11; ++c)
        for (int r = 0; r < 11; ++r) S[r + (size_t)nn * c] = f->Sigma[r + (size_t)n0 * c];
    free(f->Sigma);
    f->Sigma = S;
    f->N = n;
    return EQVIO_OK;
}

/* ---- outputs ---- */
double eqo_get_time(const eqo_filter* f) { return f->currentTime; }
int eqo_get_num_landmarks(const eqo_filter* f) { return f->N; }
/* VIOFilter::stateEstimate, VIOFilter.cpp:304; stateGroupAction VIOGroup.cpp:23-45 */
int eqo_get_state(const eqo_filter* f, double pose[7], double velocity[3], double cam_offset[7], int* n, int cap,
                  int* ids, double* landmarks) {
    hat_base h;
    compute_hat_base(f, &h);
    if (pose) se3_to_pose7(pose, &h.pose);
    if (velocity) v3_copy(velocity, h.vel);
    if (cam_offset) se3_to_pose7(cam_offset, &f->camOffset);
    if (n) *n = f->N;
    for (int i = 0; i < f->N && i < cap; ++i) {
        if (ids) ids[i] = f->id[i];
        if (landmarks) qhat_i(f, i, landmarks + 3 * i);
    }
    return EQVIO_OK;
}
int eqo_get_covariance(const eqo_filter* f, double* dst, int ld) {
    int n = NS(f);
    for (int c = 0; c < n; ++c) memcpy(dst + (size_t)ld * c, f->Sigma + (size_t)n * c, sizeof(double) * n);
    return EQVIO_OK;
}
int eqo_get_bias(const eqo_filter* f, double bias[6]) { memcpy(bias, f->inputBias, 6 * sizeof(double)); return EQVIO_OK; }

size_t eqo_snapshot_size(int N) {
    size_t n = EQVIO_SIGMA_BASE_SIZE + 3 * (size_t)N;
    return EQVIO_SNAPSHOT_HEADER + EQVIO_SNAPSHOT_PER_LANDMARK * (size_t)N + n * n;
}
static void put_se3(double* d, const se3* P) {
    d[0] = P->R.w; d[1] = P->R.x; d[2] = P->R.y; d[3] = P->R.z; d[4] = P->x[0]; d[5] = P->x[1]; d[6] = P->x[2];
}
static void take_se3(se3* P, const double* d) {
    P->R.w = d[0]; P->R.x = d[1]; P->R.y = d[2]; P->R.z = d[3]; P->x[0] = d[4]; P->x[1] = d[5]; P->x[2] = d[6];
}
int eqo_get_snapshot(const eqo_filter* f, double* d, size_t cap) {
    if (cap < eqo_snapshot_size(f->N)) return EQVIO_ERR_ARG;
    d[0] = f->N; d[1] = f->currentTime; d[2] = f->initialised; d[3] = f->accTime;
    memcpy(d + 4, f->inputBias, 6 * sizeof(double));
    memcpy(d + 10, f->curOmega, 3 * sizeof(double)); memcpy(d + 13, f->curAccel, 3 * sizeof(double));
    memcpy(d + 16, f->accOmega, 3 * sizeof(double)); memcpy(d + 19, f->accAccel, 3 * sizeof(double));
    put_se3(d + 22, &f->pose0);
    memcpy(d + 29, f->vel0, 3 * sizeof(double));
    put_se3(d + 32, &f->camOffset);
    put_se3(d + 39, &f->XA);
    memcpy(d + 46, f->Xw, 3 * sizeof(double));
    double* L = d + EQVIO_SNAPSHOT_HEADER;
    for (int i = 0; i < f->N; ++i, L += EQVIO_SNAPSHOT_PER_LANDMARK) {
        L[0] = f->id[i];
        v3_copy(L + 1, f->q0 + 3 * i);
        L[4] = f->Q[i].R.w; L[5] = f->Q[i].R.x; L[6] = f->Q[i].R.y; L[7] = f->Q[i].R.z; L[8] = f->Q[i].a;
    }
    size_t n = NS(f);
    memcpy(L, f->Sigma, sizeof(double) * n * n);
    return EQVIO_OK;
}
int eqo_set_snapshot(eqo_filter* f, const double* d, size_t len) {
    if (len < EQVIO_SNAPSHOT_HEADER) return EQVIO_ERR_ARG;
    int N = (int)d[0];
    if (N < 0 || len < eqo_snapshot_size(N)) return EQVIO_ERR_ARG;
    ensure_cap(f, N);
    f->N = N; f->currentTime = d[1]; f->initialised = (int)d[2]; f->accTime = d[3];
    memcpy(f->inputBias, d + 4, 6 * sizeof(double));
    memcpy(f->curOmega, d + 10, 3 * sizeof(double)); memcpy(f->curAccel, d + 13, 3 * sizeof(double));
    memcpy(f->accOmega, d + 16, 3 * sizeof(double)); memcpy(f->accAccel, d + 19, 3 * sizeof(double));
    take_se3(&f->pose0, d + 22);
    memcpy(f->vel0, d + 29, 3 * sizeof(double));
    take_se3(&f->camOffset, d + 32);
    take_se3(&f->XA, d + 39);
    memcpy(f->Xw, d + 46, 3 * sizeof(double));
    const double* L = d + EQVIO_SNAPSHOT_HEADER;
    for (int i = 0; i < N; ++i, L += EQVIO_SNAPSHOT_PER_LANDMARK) {
        f->id[i] = (int)L[0];
        v3_copy(f->q0 + 3 * i, L + 1);
        f->Q[i].R.w = L[4]; f->Q[i].R.x = L[5]; f->Q[i].R.y = L[6]; f->Q[i].R.z = L[7]; f->Q[i].a = L[8];
    }
    size_t n = NS(f);
    free(f->Sigma);
    f->Sigma = (double*)malloc(sizeof(double) * n * n);
    memcpy(f->Sigma, L, sizeof(double) * n * n);
    return EQVIO_OK;
}
