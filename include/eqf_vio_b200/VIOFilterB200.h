// VIOFilterB200.h — the reference's `class VIOFilter`, on the reference's OWN types, backed by the B200 library.
//
// Drop-in for eqf_vio/include/eqf_vio/VIOFilter.h:64-88: the same constructors, the same public `settings`
// member, the same nine methods and the same operator<<, taking and returning the reference's IMUVelocity,
// VisionMeasurement, VIOState, AuxiliaryFilterData, VIOFilter::Settings and Eigen::MatrixXd.  A caller such as
// eqf_vio/src/main.cpp changes ONE line — `VIOFilter filter(filterSettings);` (main.cpp:86) becomes
// `VIOFilterB200 filter(filterSettings);` — and the calls at main.cpp:116,129,134-140 compile and run unchanged.
//
// Build: add `-I<this repo>/include` next to the reference's own include paths and link
// `eqf_vio_b200/csrc/libeqvio_b200.so`.  This header needs the reference's headers (and therefore
// Eigen); the library itself does not.
//
// Behaviour kept from the reference (file:line = paths under the reference tree, eqf_vio/):
//   * move-only (the reference owns a unique_ptr<Settings>; the ROS node move-assigns, eqf_vio_ros_node.cpp:59);
//   * `settings` is public and read at use time (src/VIOFilter.cpp:126,163-175): a change made through the pointer
//     between two calls takes effect at the next call here too (it is re-sent when it differs from what the device
//     side holds).  Constructor-time effects (initial variances, initial biases, camera offset) are not re-applied,
//     exactly as in the reference;
//   * the Settings-less constructors (VIOFilter.h:70-71) leave `settings` null and start from Sigma = I(11)
//     (VIOFilter.h:47); where the reference would dereference the null pointer (src/VIOFilter.cpp:126) this class
//     throws std::logic_error;
//   * VIOFilter(aux, settings) does NOT take the initial biases from the settings (src/VIOFilter.cpp:50-58);
//   * void returns and silent no-ops: first IMU sample only initialises, dt <= 0 integrates nothing, a stale or
//     empty vision frame is dropped (src/VIOFilter.cpp:122-124,147-152,235-236,258-259).
// Errors: SO3FromVectors on opposing vectors throws std::domain_error like the reference (libs/core/src/SO3.cpp:160);
// what the reference checks by assert (id order :239-240, NaN :190,299) throws std::invalid_argument /
// std::runtime_error.  Device-side conditions are sticky flags reported by the next synchronising call
// (stateEstimate / stateCovariance / operator<<), see include/eqvio.h "Error model".
#pragma once

#include <cstring>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "eqf_vio/VIOFilter.h"
#include "eqf_vio/VIOFilterSettings.h"

#include "../eqvio.h"

class VIOFilterB200 {
  protected:
    AuxiliaryFilterData auxData;                       // VIOFilter.h:43
    mutable eqvio_handle_t handle = nullptr;           // the device-resident inputBias, xi0, X, Sigma, flags, time, velocities
    mutable eqvio_settings_t sent;                     // the settings the device side currently holds
    int device = 0;

  public:
    std::unique_ptr<VIOFilter::Settings> settings;     // VIOFilter.h:67

    // ---- Setup (VIOFilter.h:69-75) ----
    VIOFilterB200() { open(defaults(), true); }
    VIOFilterB200(const AuxiliaryFilterData& auxiliaryData) {
        open(defaults(), true);
        setAuxiliaryData(auxiliaryData);
    }
    VIOFilterB200(const AuxiliaryFilterData& auxiliaryData, const VIOFilter::Settings& settings_) {
        settings = std::make_unique<VIOFilter::Settings>(settings_);
        eqvio_settings_t s = toPOD(settings_);
        for (int i = 0; i < 3; ++i) s.initialOmegaBias[i] = s.initialAccelBias[i] = 0.0;   // src/VIOFilter.cpp:50-58 leaves inputBias zero
        open(s, false);
        sent = toPOD(settings_);
        setAuxiliaryData(auxiliaryData);
    }
    VIOFilterB200(const VIOFilter::Settings& settings_) {
        settings = std::make_unique<VIOFilter::Settings>(settings_);
        open(toPOD(settings_), false);
    }
    VIOFilterB200(const VIOFilter::Settings& settings_, int cudaDevice) : device(cudaDevice) {   // extension: which GPU
        settings = std::make_unique<VIOFilter::Settings>(settings_);
        open(toPOD(settings_), false);
    }
    VIOFilterB200(VIOFilterB200&& o) noexcept : auxData(o.auxData), handle(o.handle), sent(o.sent), device(o.device), settings(std::move(o.settings)) { o.handle = nullptr; }
    VIOFilterB200& operator=(VIOFilterB200&& o) noexcept {
        if (this != &o) {
            close();
            auxData = o.auxData; handle = o.handle; sent = o.sent; device = o.device; settings = std::move(o.settings);
            o.handle = nullptr;
        }
        return *this;
    }
    VIOFilterB200(const VIOFilterB200&) = delete;
    VIOFilterB200& operator=(const VIOFilterB200&) = delete;
    ~VIOFilterB200() { close(); }

    void initialiseFromIMUData(const IMUVelocity& imuVelocity) {                               // VIOFilter.h:74
        check(eqvio_initialise_from_imu(handle, imuVelocity.omega.data(), imuVelocity.accel.data()), "initialiseFromIMUData");
    }
    void reset() { check(eqvio_reset(handle), "reset"); }                                      // VIOFilter.h:75

    // ---- Input (VIOFilter.h:78-81) ----
    void setAuxiliaryData(const AuxiliaryFilterData& auxiliaryData) {
        auxData = auxiliaryData;
        const double q[4] = {auxiliaryData.initialAttitude.w(), auxiliaryData.initialAttitude.x(), auxiliaryData.initialAttitude.y(), auxiliaryData.initialAttitude.z()};
        double cam[7];
        putPose(cam, auxiliaryData.cameraOffset);
        check(eqvio_set_auxiliary_data(handle, q, auxiliaryData.initialPosition.data(), cam), "setAuxiliaryData");
    }
    void setInertialPoints(const std::vector<Point3d>& inertialPoints) {
        syncSettings("setInertialPoints");
        std::vector<int> ids(inertialPoints.size());
        std::vector<double> p(3 * inertialPoints.size());
        for (size_t i = 0; i < inertialPoints.size(); ++i) {
            ids[i] = inertialPoints[i].id;
            for (int k = 0; k < 3; ++k) p[3 * i + k] = inertialPoints[i].p(k);
        }
        check(eqvio_set_inertial_points(handle, (int)ids.size(), ids.data(), p.data()), "setInertialPoints");
    }
    void processIMUData(const IMUVelocity& imuVelocity) {
        syncSettings("processIMUData");
        check(eqvio_process_imu(handle, imuVelocity.stamp, imuVelocity.omega.data(), imuVelocity.accel.data()), "processIMUData");
    }
    void processVisionData(const VisionMeasurement& measurement) {
        syncSettings("processVisionData");
        const size_t n = measurement.bearings.size();   // the reference iterates the vector, not numberOfBearings
        idBuffer.resize(n);
        bearingBuffer.resize(3 * n);
        for (size_t i = 0; i < n; ++i) {
            idBuffer[i] = measurement.bearings[i].id;
            for (int k = 0; k < 3; ++k) bearingBuffer[3 * i + k] = measurement.bearings[i].p(k);
        }
        check(eqvio_process_vision(handle, measurement.stamp, (int)n, idBuffer.data(), bearingBuffer.data()), "processVisionData");
    }

    // ---- Output (VIOFilter.h:84-87) ----
    double getTime() const {
        double t = -1;
        check(eqvio_get_time(handle, &t), "getTime");
        return t;
    }
    VIOState stateEstimate() const {
        int n = 0;
        check(eqvio_get_num_landmarks(handle, &n), "stateEstimate");
        double pose[7], vel[3], cam[7];
        std::vector<int> ids(n > 0 ? n : 1);
        std::vector<double> lm(3 * (size_t)(n > 0 ? n : 1));
        check(eqvio_get_state(handle, pose, vel, cam, &n, n, ids.data(), lm.data()), "stateEstimate");
        VIOState xi;
        takePose(xi.pose, pose);
        takePose(xi.cameraOffset, cam);
        xi.velocity = Eigen::Vector3d(vel[0], vel[1], vel[2]);
        xi.bodyLandmarks.resize(n);
        for (int i = 0; i < n; ++i) {
            xi.bodyLandmarks[i].id = ids[i];
            xi.bodyLandmarks[i].p = Eigen::Vector3d(lm[3 * i], lm[3 * i + 1], lm[3 * i + 2]);
        }
        return xi;
    }
    Eigen::MatrixXd stateCovariance() const {
        int N = 0;
        check(eqvio_get_num_landmarks(handle, &N), "stateCovariance");
        const int n = EQVIO_SIGMA_BASE_SIZE + 3 * N;
        Eigen::MatrixXd Sigma(n, n);                    // column-major, like the library's buffer
        check(eqvio_get_covariance(handle, Sigma.data(), n), "stateCovariance");
        return Sigma;
    }
    // The internal-state row of src/VIOFilter.cpp:311-341: xi0 pose (x, then q wxyz), xi0 velocity, X.A (x, q), X.w, N,
    // per landmark id, q0, Q quaternion, Q scale, then Sigma row by row; every number through the stream's own
    // formatting.  (Eigen's IOFormat(-1, 0, ", ", ", ") pads Sigma's entries to a common width; no padding here.)
    friend std::ostream& operator<<(std::ostream& os, const VIOFilterB200& filter) {
        int N = 0;
        check(eqvio_get_num_landmarks(filter.handle, &N), "operator<<");
        std::vector<double> d(eqvio_snapshot_size(N));
        check(eqvio_get_snapshot(filter.handle, d.data(), d.size()), "operator<<");
        auto pose = [&](const double* q) { os << q[4] << ", " << q[5] << ", " << q[6] << ", " << q[0] << ", " << q[1] << ", " << q[2] << ", " << q[3] << ", "; };
        pose(&d[22]);
        os << d[29] << ", " << d[30] << ", " << d[31] << ", ";
        pose(&d[39]);
        os << d[46] << ", " << d[47] << ", " << d[48] << ", ";
        os << N;
        const double* rec = d.data() + EQVIO_SNAPSHOT_HEADER;
        for (int i = 0; i < N; ++i, rec += EQVIO_SNAPSHOT_PER_LANDMARK) {
            os << ", " << (int)rec[0];
            for (int k = 1; k < EQVIO_SNAPSHOT_PER_LANDMARK; ++k) os << ", " << rec[k];
        }
        const int n = EQVIO_SIGMA_BASE_SIZE + 3 * N;
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c) os << ", " << rec[r + (size_t)n * c];
        return os;
    }

    // ---- beyond the reference: what a caller may want from the device side ----
    eqvio_handle_t nativeHandle() const { return handle; }
    Eigen::Matrix<double, 6, 1> inputBiasEstimate() const {
        Eigen::Matrix<double, 6, 1> b;
        check(eqvio_get_bias(handle, b.data()), "inputBiasEstimate");
        return b;
    }

    static eqvio_settings_t toPOD(const VIOFilter::Settings& s) {   // all 22 fields of VIOFilterSettings.h:29-50
        eqvio_settings_t o;
        std::memset(&o, 0, sizeof o);
        o.biasOmegaProcessVariance = s.biasOmegaProcessVariance;
        o.biasAccelProcessVariance = s.biasAccelProcessVariance;
        o.gravityProcessVariance = s.gravityProcessVariance;
        o.velocityProcessVariance = s.velocityProcessVariance;
        o.pointProcessVariance = s.pointProcessVariance;
        o.velOmegaVariance = s.velOmegaVariance;
        o.velAccelVariance = s.velAccelVariance;
        o.measurementVariance = s.measurementVariance;
        o.initialGravityVariance = s.initialGravityVariance;
        o.initialVelocityVariance = s.initialVelocityVariance;
        o.initialPointVariance = s.initialPointVariance;
        o.initialBiasOmegaVariance = s.initialBiasOmegaVariance;
        o.initialBiasAccelVariance = s.initialBiasAccelVariance;
        o.initialSceneDepth = s.initialSceneDepth;
        o.outlierThreshold = s.outlierThreshold;
        o.useInnovationLift = s.useInnovationLift ? 1 : 0;
        o.useDiscreteInnovationLift = s.useDiscreteInnovationLift ? 1 : 0;
        o.useDiscreteVelocityLift = s.useDiscreteVelocityLift ? 1 : 0;
        o.fastRiccati = s.fastRiccati ? 1 : 0;
        for (int i = 0; i < 3; ++i) {
            o.initialAccelBias[i] = s.initialAccelBias(i);
            o.initialOmegaBias[i] = s.initialOmegaBias(i);
        }
        putPose(o.cameraOffset, s.cameraOffset);
        return o;
    }

  private:
    std::vector<int> idBuffer;
    std::vector<double> bearingBuffer;

    static eqvio_settings_t defaults() {
        eqvio_settings_t s;
        eqvio_settings_default(&s);
        return s;
    }
    static void putPose(double* p, const SE3& P) {      // x y z qw qx qy qz
        const Eigen::Quaterniond q = P.R().asQuaternion();
        for (int k = 0; k < 3; ++k) p[k] = P.x()(k);
        p[3] = q.w(); p[4] = q.x(); p[5] = q.y(); p[6] = q.z();
    }
    static void takePose(SE3& P, const double* p) {
        P.R().fromQuaternion(Eigen::Quaterniond(p[3], p[4], p[5], p[6]));
        P.x() = Eigen::Vector3d(p[0], p[1], p[2]);
    }
    static void check(int status, const char* where) {
        if (status >= 0) return;
        const std::string msg = std::string("VIOFilterB200::") + where + ": " + eqvio_status_string(status);
        if (status == EQVIO_ERR_SINGULAR_CHART) throw std::domain_error(msg);
        if (status == EQVIO_ERR_UNSORTED || status == EQVIO_ERR_ARG) throw std::invalid_argument(msg);
        throw std::runtime_error(msg);
    }
    // identityStart: the Settings-less constructors start from the member defaults of VIOFilter.h:46-55, Sigma = I(11)
    void open(const eqvio_settings_t& s, bool identityStart) {
        check(eqvio_create(&s, device, &handle), "VIOFilterB200");
        sent = s;
        if (identityStart) check(eqvio_reset(handle), "VIOFilterB200");
    }
    void close() {
        if (handle) eqvio_destroy(handle);
        handle = nullptr;
    }
    void syncSettings(const char* where) {
        if (!settings) throw std::logic_error(std::string("VIOFilterB200::") + where + ": settings is null (the reference dereferences it here, src/VIOFilter.cpp:126)");
        const eqvio_settings_t now = toPOD(*settings);
        if (std::memcmp(&now, &sent, sizeof now) != 0) {
            check(eqvio_set_settings(handle, &now), where);
            sent = now;
        }
    }
};
