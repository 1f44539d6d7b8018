// eqvio_math.cuh — POD Lie-group and chart math for device code (and the few host uses).
// Device-side counterparts of the reference's pimpl classes:
//   SO3  (quaternion-backed)  eqf_vio/libs/core/src/SO3.cpp
//   SE3                       eqf_vio/libs/core/src/SE3.cpp
//   SOT3                      eqf_vio/libs/core/src/SOT3.cpp
//   sphere charts             eqf_vio/src/VIOState.cpp:199-251
// Rotations are kept as quaternions (w,x,y,z) and composed/applied with Eigen::Quaterniond's
// formulas so round-off follows the reference; 3x3 matrices are row-major structs here because
// they only ever live in registers.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define EQ_HD __host__ __device__ __forceinline__

namespace eqvio {

struct V3 { double x, y, z; };
struct Quat { double w, x, y, z; };
struct M3 { double m[3][3]; };  // m[r][c]
struct Se3 { Quat R; V3 x; };
struct Sot3 { Quat R; double a; };

EQ_HD V3 v3(double x, double y, double z) { V3 r = {x, y, z}; return r; }
EQ_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
EQ_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
EQ_HD V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
EQ_HD V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
EQ_HD V3 operator*(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
EQ_HD V3 operator/(V3 a, double s) { return v3(a.x / s, a.y / s, a.z / s); }
EQ_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
EQ_HD double norm(V3 a) { return sqrt(dot(a, a)); }
EQ_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
EQ_HD V3 normalized(V3 a) { double n = norm(a); return v3(a.x / n, a.y / n, a.z / n); }
EQ_HD double v3_get(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

EQ_HD M3 m3_identity() { M3 r = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; return r; }
EQ_HD M3 m3_zero() { M3 r = {{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}}; return r; }
EQ_HD M3 operator*(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
EQ_HD V3 operator*(const M3& a, V3 v) {
    return v3(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
              a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}
EQ_HD M3 operator+(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j];
    return r;
}
EQ_HD M3 operator*(const M3& a, double s) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] * s;
    return r;
}
EQ_HD M3 transpose(const M3& a) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
    return r;
}
EQ_HD M3 outer(V3 a, V3 b) {
    M3 r = {{{a.x * b.x, a.x * b.y, a.x * b.z}, {a.y * b.x, a.y * b.y, a.y * b.z}, {a.z * b.x, a.z * b.y, a.z * b.z}}};
    return r;
}
// SO3::skew, SO3.cpp:110-114
EQ_HD M3 skew(V3 v) { M3 r = {{{0, -v.z, v.y}, {v.z, 0, -v.x}, {-v.y, v.x, 0}}}; return r; }
// 3x3 inverse by cofactors (Eigen fixed-size inverse; EqFMatrices.cpp:310)
EQ_HD M3 inverse(const M3& a) {
    double c00 = a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1];
    double c10 = a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2];
    double c20 = a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0];
    double id = 1.0 / (a.m[0][0] * c00 + a.m[0][1] * c10 + a.m[0][2] * c20);
    M3 r;
    r.m[0][0] = c00 * id; r.m[1][0] = c10 * id; r.m[2][0] = c20 * id;
    r.m[0][1] = (a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2]) * id;
    r.m[1][1] = (a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0]) * id;
    r.m[2][1] = (a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1]) * id;
    r.m[0][2] = (a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1]) * id;
    r.m[1][2] = (a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2]) * id;
    r.m[2][2] = (a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0]) * id;
    return r;
}

// ---- SO3 as Eigen::Quaterniond (SO3.cpp) ----
EQ_HD Quat q_identity() { Quat q = {1, 0, 0, 0}; return q; }
EQ_HD Quat operator*(Quat a, Quat b) {  // SO3.cpp:66-70, no renormalisation
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
EQ_HD Quat q_inverse(Quat a) {  // Eigen inverse(): conjugate / squaredNorm (SO3.cpp:74)
    double n2 = a.w * a.w + a.x * a.x + a.y * a.y + a.z * a.z;
    Quat r = {a.w / n2, -a.x / n2, -a.y / n2, -a.z / n2};
    return r;
}
EQ_HD V3 rotate(Quat q, V3 v) {  // Eigen _transformVector (SO3.cpp:58)
    V3 u = v3(q.x, q.y, q.z);
    V3 uv = cross(u, v);
    uv = uv + uv;
    return v + q.w * uv + cross(u, uv);
}
EQ_HD V3 rotate_inv(Quat q, V3 v) { return rotate(q_inverse(q), v); }  // SO3::applyInverse, SO3.cpp:108
EQ_HD M3 to_matrix(Quat q) {  // Eigen toRotationMatrix (SO3.cpp:92)
    double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    M3 r = {{{1 - (tyy + tzz), txy - twz, txz + twy}, {txy + twz, 1 - (txx + tzz), tyz - twx}, {txz - twy, tyz + twx, 1 - (txx + tyy)}}};
    return r;
}
EQ_HD Quat from_matrix(const M3& m) {  // Eigen Quaterniond(Matrix3d) (SO3.cpp:100)
    Quat q;
    double t = m.m[0][0] + m.m[1][1] + m.m[2][2];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (m.m[2][1] - m.m[1][2]) * t;
        q.y = (m.m[0][2] - m.m[2][0]) * t;
        q.z = (m.m[1][0] - m.m[0][1]) * t;
    } else if (m.m[0][0] >= m.m[1][1] && m.m[0][0] >= m.m[2][2]) {  // i = 0, j = 1, k = 2
        t = sqrt(m.m[0][0] - m.m[1][1] - m.m[2][2] + 1.0);
        q.x = 0.5 * t;
        t = 0.5 / t;
        q.w = (m.m[2][1] - m.m[1][2]) * t;
        q.y = (m.m[1][0] + m.m[0][1]) * t;
        q.z = (m.m[2][0] + m.m[0][2]) * t;
    } else if (m.m[1][1] > m.m[0][0] && m.m[1][1] >= m.m[2][2]) {  // i = 1, j = 2, k = 0
        t = sqrt(m.m[1][1] - m.m[2][2] - m.m[0][0] + 1.0);
        q.y = 0.5 * t;
        t = 0.5 / t;
        q.w = (m.m[0][2] - m.m[2][0]) * t;
        q.z = (m.m[2][1] + m.m[1][2]) * t;
        q.x = (m.m[0][1] + m.m[1][0]) * t;
    } else {  // i = 2, j = 0, k = 1
        t = sqrt(m.m[2][2] - m.m[0][0] - m.m[1][1] + 1.0);
        q.z = 0.5 * t;
        t = 0.5 / t;
        q.w = (m.m[1][0] - m.m[0][1]) * t;
        q.x = (m.m[0][2] + m.m[2][0]) * t;
        q.y = (m.m[1][2] + m.m[2][1]) * t;
    }
    return q;
}
// SO3::SO3Exp, SO3.cpp:122-140
EQ_HD Quat so3_exp(V3 w) {
    double th = norm(w), A, B;
    if (fabs(th) >= 1e-8) { A = sin(th) / th; B = (1 - cos(th)) / (th * th); }
    else { A = 1.0; B = 0.5; }
    M3 wx = skew(w);
    return from_matrix(m3_identity() + wx * A + (wx * wx) * B);
}
// SO3::SO3FromVectors, SO3.cpp:155-167; *singular set when |1+c| <= 1e-8 (the reference throws)
EQ_HD Quat so3_from_vectors(V3 origin, V3 dest, int* singular) {
    V3 a = normalized(origin), b = normalized(dest);
    V3 v = cross(a, b);
    double c = dot(a, b);
    if (fabs(1 + c) <= 1e-8) *singular = 1;
    M3 sv = skew(v);
    return from_matrix(m3_identity() + (sv + (sv * sv) * (1 / (1 + c))));
}

// ---- SE3 (SE3.cpp) ----
EQ_HD Se3 se3_identity() { Se3 P; P.R = q_identity(); P.x = v3(0, 0, 0); return P; }
EQ_HD V3 operator*(const Se3& P, V3 p) { return rotate(P.R, p) + P.x; }                               // :65
EQ_HD Se3 operator*(const Se3& a, const Se3& b) { Se3 r; r.R = a.R * b.R; r.x = a.x + rotate(a.R, b.x); return r; }  // :73-78
EQ_HD Se3 inverse(const Se3& a) { Se3 r; r.R = q_inverse(a.R); r.x = -rotate(r.R, a.x); return r; }   // :82-85
// SE3::SE3Exp, SE3.cpp:139-164
EQ_HD Se3 se3_exp(V3 w, V3 v) {
    double th = norm(w), A, B, C;
    if (fabs(th) >= 1e-12) { A = sin(th) / th; B = (1 - cos(th)) / (th * th); C = (1 - A) / (th * th); }
    else { A = 1.0; B = 0.5; C = 1.0 / 6.0; }
    M3 wx = skew(w), wx2 = wx * wx;
    Se3 P;
    P.R = from_matrix(m3_identity() + wx * A + wx2 * B);
    P.x = (m3_identity() + wx * B + wx2 * C) * v;
    return P;
}
// Adjoint(P) applied to a twist (omega, v): (R omega, [x]x R omega + R v), SE3.cpp:95-103 (matrix form)
EQ_HD void se3_adjoint_apply(const Se3& P, V3 om, V3 v, V3* om_out, V3* v_out) {
    M3 R = to_matrix(P.R);
    M3 sxR = skew(P.x) * R;
    *om_out = R * om;
    *v_out = sxR * om + R * v;
}

// ---- SOT3 (SOT3.cpp) ----
EQ_HD Sot3 sot3_identity() { Sot3 Q; Q.R = q_identity(); Q.a = 1.0; return Q; }
EQ_HD Sot3 operator*(Sot3 a, Sot3 b) { Sot3 r; r.R = a.R * b.R; r.a = a.a * b.a; return r; }  // :69-74
EQ_HD Sot3 inverse(Sot3 a) { Sot3 r; r.R = q_inverse(a.R); r.a = 1.0 / a.a; return r; }       // :78-81
EQ_HD V3 operator*(Sot3 Q, V3 p) { return Q.a * rotate(Q.R, p); }                             // :62
EQ_HD M3 as_matrix3(Sot3 Q) { return to_matrix(Q.R) * Q.a; }                                  // :107-110
EQ_HD Sot3 sot3_exp(V3 w, double s) { Sot3 r; r.R = so3_exp(w); r.a = exp(s); return r; }     // :127-132

// ---- sphere charts (VIOState.cpp:199-251) ----
struct M23 { double m[2][3]; };
struct M32 { double m[3][2]; };
EQ_HD void e3_project_sphere(V3 eta, double* y) {  // :199-204
    double d = 1 - eta.z;
    y[0] = eta.x / d;
    y[1] = eta.y / d;
}
EQ_HD M23 e3_project_sphere_diff(V3 eta) {  // :213-220
    double omz = 1 - eta.z;
    double s = 1.0 / (omz * omz);  // pow(1 - e3.eta, -2.0)
    M23 D = {{{s * omz, 0.0, s * eta.x}, {0.0, s * omz, s * eta.y}}};
    return D;
}
EQ_HD M32 e3_project_sphere_inv_diff(double y0, double y1) {  // :222-228
    double n2 = y0 * y0 + y1 * y1;
    double s = 2.0 / ((n2 + 1.0) * (n2 + 1.0));
    M32 D = {{{s * ((n2 + 1.0) - 2 * y0 * y0), s * (-2 * y0 * y1)}, {s * (-2 * y1 * y0), s * ((n2 + 1.0) - 2 * y1 * y1)}, {s * 2 * y0, s * 2 * y1}}};
    return D;
}
EQ_HD Quat sphere_rot(V3 pole, int* singular) { return so3_from_vectors(-pole, v3(0, 0, 1), singular); }
EQ_HD void stereo_sphere_chart(V3 eta, V3 pole, double* y, int* singular) {  // :230-234
    e3_project_sphere(rotate(sphere_rot(pole, singular), eta), y);
}
EQ_HD M23 stereo_sphere_chart_diff(V3 eta, V3 pole, int* singular) {  // :242-246
    Quat q = sphere_rot(pole, singular);
    M23 d = e3_project_sphere_diff(rotate(q, eta));
    M3 R = to_matrix(q);
    M23 r;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[i][j] = d.m[i][0] * R.m[0][j] + d.m[i][1] * R.m[1][j] + d.m[i][2] * R.m[2][j];
    return r;
}
EQ_HD M32 stereo_sphere_chart_inv_diff(double y0, double y1, V3 pole, int* singular) {  // :248-251
    Quat q = sphere_rot(pole, singular);
    M3 Ri = to_matrix(q_inverse(q));
    M32 d = e3_project_sphere_inv_diff(y0, y1);
    M32 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) r.m[i][j] = Ri.m[i][0] * d.m[0][j] + Ri.m[i][1] * d.m[1][j] + Ri.m[i][2] * d.m[2][j];
    return r;
}

}  // namespace eqvio
