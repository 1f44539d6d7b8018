// ref_capi.cpp — C wrapper around the REFERENCE's own `class VIOFilter`, compiled from the unmodified
// sources under /root/reference against the Eigen stand-in in this directory.  TEST INFRASTRUCTURE ONLY:
// output goes to oracle/_ref/ (git-ignored), used to pin the restated oracle (oracle/eqvio_oracle.c)
// against the reference's source-level behaviour.  Nothing here is copied from the reference; the
// reference's files are compiled where they lie.
#include <cstring>
#include <vector>

#include "eqf_vio/EqFMatrices.h"
#include "eqf_vio/VIOFilter.h"
#include "eqf_vio/VIOFilterSettings.h"

#include "../../include/eqvio.h"

using namespace Eigen;

namespace {
struct RefFilter : public VIOFilter {
    using VIOFilter::VIOFilter;
    using VIOFilter::accumulatedTime;
    using VIOFilter::accumulatedVelocity;
    using VIOFilter::currentTime;
    using VIOFilter::currentVelocity;
    using VIOFilter::initialisedFlag;
    using VIOFilter::inputBias;
    using VIOFilter::Sigma;
    using VIOFilter::X;
    using VIOFilter::xi0;
};

VIOFilter::Settings to_settings(const eqvio_settings_t* s) {
    VIOFilter::Settings o;
    o.biasOmegaProcessVariance = s->biasOmegaProcessVariance;
    o.biasAccelProcessVariance = s->biasAccelProcessVariance;
    o.gravityProcessVariance = s->gravityProcessVariance;
    o.velocityProcessVariance = s->velocityProcessVariance;
    o.pointProcessVariance = s->pointProcessVariance;
    o.velOmegaVariance = s->velOmegaVariance;
    o.velAccelVariance = s->velAccelVariance;
    o.measurementVariance = s->measurementVariance;
    o.initialGravityVariance = s->initialGravityVariance;
    o.initialVelocityVariance = s->initialVelocityVariance;
    o.initialPointVariance = s->initialPointVariance;
    o.initialBiasOmegaVariance = s->initialBiasOmegaVariance;
    o.initialBiasAccelVariance = s->initialBiasAccelVariance;
    o.initialSceneDepth = s->initialSceneDepth;
    o.outlierThreshold = s->outlierThreshold;
    o.useInnovationLift = s->useInnovationLift != 0;
    o.useDiscreteInnovationLift = s->useDiscreteInnovationLift != 0;
    o.useDiscreteVelocityLift = s->useDiscreteVelocityLift != 0;
    o.fastRiccati = s->fastRiccati != 0;
    o.initialAccelBias = Vector3d(s->initialAccelBias[0], s->initialAccelBias[1], s->initialAccelBias[2]);
    o.initialOmegaBias = Vector3d(s->initialOmegaBias[0], s->initialOmegaBias[1], s->initialOmegaBias[2]);
    o.cameraOffset.x() = Vector3d(s->cameraOffset[0], s->cameraOffset[1], s->cameraOffset[2]);
    o.cameraOffset.R().fromQuaternion(Quaterniond(s->cameraOffset[3], s->cameraOffset[4], s->cameraOffset[5], s->cameraOffset[6]));
    return o;
}
void put_se3(double* d, const SE3& P) {
    const Quaterniond q = P.R().asQuaternion();
    d[0] = q.w(); d[1] = q.x(); d[2] = q.y(); d[3] = q.z();
    d[4] = P.x()(0); d[5] = P.x()(1); d[6] = P.x()(2);
}
void take_se3(SE3& P, const double* d) {
    P.R().fromQuaternion(Quaterniond(d[0], d[1], d[2], d[3]));
    P.x() = Vector3d(d[4], d[5], d[6]);
}
VisionMeasurement make_meas(double stamp, int n, const int* ids, const double* y) {
    VisionMeasurement m;
    m.stamp = stamp;
    m.numberOfBearings = n;
    m.bearings.resize(n);
    for (int i = 0; i < n; ++i) { m.bearings[i].id = ids[i]; m.bearings[i].p = Vector3d(y[3 * i], y[3 * i + 1], y[3 * i + 2]); }
    return m;
}
}  // namespace

extern "C" {

void* ref_create(const eqvio_settings_t* s) { return new RefFilter(to_settings(s)); }
void ref_destroy(void* h) { delete static_cast<RefFilter*>(h); }
int ref_process_imu(void* h, double stamp, const double* omega, const double* accel) {
    RefFilter* f = static_cast<RefFilter*>(h);
    IMUVelocity v;
    v.stamp = stamp;
    v.omega = Vector3d(omega[0], omega[1], omega[2]);
    v.accel = Vector3d(accel[0], accel[1], accel[2]);
    try { f->processIMUData(v); } catch (const std::domain_error&) { return EQVIO_ERR_SINGULAR_CHART; }
    return 0;
}
int ref_process_vision(void* h, double stamp, int n, const int* ids, const double* y) {
    RefFilter* f = static_cast<RefFilter*>(h);
    try { f->processVisionData(make_meas(stamp, n, ids, y)); } catch (const std::domain_error&) { return EQVIO_ERR_SINGULAR_CHART; }
    return 0;
}
int ref_num_landmarks(void* h) { return (int)static_cast<RefFilter*>(h)->X.id.size(); }
double ref_get_time(void* h) { return static_cast<RefFilter*>(h)->getTime(); }

size_t ref_snapshot_size(int N) { size_t n = 11 + 3 * (size_t)N; return EQVIO_SNAPSHOT_HEADER + EQVIO_SNAPSHOT_PER_LANDMARK * (size_t)N + n * n; }
int ref_get_snapshot(void* h, double* d) {
    RefFilter* f = static_cast<RefFilter*>(h);
    const int N = (int)f->X.id.size();
    d[0] = N; d[1] = f->currentTime; d[2] = f->initialisedFlag ? 1.0 : 0.0; d[3] = f->accumulatedTime;
    for (int i = 0; i < 6; ++i) d[4 + i] = f->inputBias(i);
    for (int i = 0; i < 3; ++i) {
        d[10 + i] = f->currentVelocity.omega(i); d[13 + i] = f->currentVelocity.accel(i);
        d[16 + i] = f->accumulatedVelocity.omega(i); d[19 + i] = f->accumulatedVelocity.accel(i);
        d[29 + i] = f->xi0.velocity(i); d[46 + i] = f->X.w(i);
    }
    put_se3(d + 22, f->xi0.pose);
    put_se3(d + 32, f->xi0.cameraOffset);
    put_se3(d + 39, f->X.A);
    double* L = d + EQVIO_SNAPSHOT_HEADER;
    for (int i = 0; i < N; ++i, L += EQVIO_SNAPSHOT_PER_LANDMARK) {
        L[0] = f->X.id[i];
        for (int k = 0; k < 3; ++k) L[1 + k] = f->xi0.bodyLandmarks[i].p(k);
        const Quaterniond q = f->X.Q[i].R().asQuaternion();
        L[4] = q.w(); L[5] = q.x(); L[6] = q.y(); L[7] = q.z(); L[8] = f->X.Q[i].a();
    }
    const int n = 11 + 3 * N;
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < n; ++r) L[r + (size_t)n * c] = f->Sigma(r, c);
    return 0;
}
int ref_set_snapshot(void* h, const double* d) {
    RefFilter* f = static_cast<RefFilter*>(h);
    const int N = (int)d[0];
    f->currentTime = d[1]; f->initialisedFlag = d[2] != 0.0; f->accumulatedTime = d[3];
    for (int i = 0; i < 6; ++i) f->inputBias(i) = d[4 + i];
    for (int i = 0; i < 3; ++i) {
        f->currentVelocity.omega(i) = d[10 + i]; f->currentVelocity.accel(i) = d[13 + i];
        f->accumulatedVelocity.omega(i) = d[16 + i]; f->accumulatedVelocity.accel(i) = d[19 + i];
        f->xi0.velocity(i) = d[29 + i]; f->X.w(i) = d[46 + i];
    }
    take_se3(f->xi0.pose, d + 22);
    take_se3(f->xi0.cameraOffset, d + 32);
    take_se3(f->X.A, d + 39);
    f->xi0.bodyLandmarks.resize(N); f->X.Q.resize(N); f->X.id.resize(N);
    const double* L = d + EQVIO_SNAPSHOT_HEADER;
    for (int i = 0; i < N; ++i, L += EQVIO_SNAPSHOT_PER_LANDMARK) {
        f->X.id[i] = (int)L[0]; f->xi0.bodyLandmarks[i].id = (int)L[0];
        f->xi0.bodyLandmarks[i].p = Vector3d(L[1], L[2], L[3]);
        f->X.Q[i].R().fromQuaternion(Quaterniond(L[4], L[5], L[6], L[7]));
        f->X.Q[i].a() = L[8];
    }
    const int n = 11 + 3 * N;
    f->Sigma = MatrixXd(n, n);
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < n; ++r) f->Sigma(r, c) = L[r + (size_t)n * c];
    return 0;
}

// pieces: the reference's free functions on the filter's current (xi0, X, Sigma)
int ref_state_matrix_A(void* h, const double* omega, double* A0) {
    RefFilter* f = static_cast<RefFilter*>(h);
    IMUVelocity v; v.stamp = 0; v.omega = Vector3d(omega[0], omega[1], omega[2]); v.accel = Vector3d(0, 0, 0);
    const MatrixXd A = EqFStateMatrixA_euclid(f->X, f->xi0, v);
    std::memcpy(A0, A.data(), sizeof(double) * A.size());
    return 0;
}
int ref_input_matrix_B(void* h, double* Bt) {
    RefFilter* f = static_cast<RefFilter*>(h);
    const MatrixXd B = EqFInputMatrixB_euclid(f->X, f->xi0);
    std::memcpy(Bt, B.data(), sizeof(double) * B.size());
    return 0;
}
int ref_output_matrix_C(void* h, double* C0) {
    RefFilter* f = static_cast<RefFilter*>(h);
    const MatrixXd C = EqFOutputMatrixC_euclid(f->xi0);
    std::memcpy(C0, C.data(), sizeof(double) * C.size());
    return 0;
}
int ref_bundle_lift(void* h, const double* gamma_eqf, double* Gamma) {
    RefFilter* f = static_cast<RefFilter*>(h);
    const int N = (int)f->X.id.size(), p = 5 + 3 * N;
    VectorXd g(p);
    for (int i = 0; i < p; ++i) g(i) = gamma_eqf[i];
    const MatrixXd Ssub = static_cast<const MatrixXd&>(f->Sigma).block(6, 6, p, p);
    const VectorXd G = bundleLift(g, f->xi0, f->X, Ssub);
    for (int i = 0; i < 9 + 3 * N; ++i) Gamma[i] = G(i);
    return 0;
}
int ref_delta(void* h, const double* y, double* delta) {
    RefFilter* f = static_cast<RefFilter*>(h);
    const int N = (int)f->X.id.size();
    std::vector<int> ids(f->X.id);
    const VisionMeasurement m = make_meas(0.0, N, ids.data(), y);
    const VectorXd d = outputCoordinateChart(outputGroupAction(f->X.inverse(), m), measureSystemState(f->xi0));
    for (int i = 0; i < 2 * N; ++i) delta[i] = d(i);
    return 0;
}
}  // extern "C"
