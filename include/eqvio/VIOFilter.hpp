// VIOFilter.hpp — header-only C++ façade with the reference's `class VIOFilter` surface
// (reference: eqf_vio/include/eqf_vio/VIOFilter.h:41-88) over the C ABI in include/eqvio.h.
//
// The reference's signatures use Eigen and its own SE3/SOT3 pimpl classes; Eigen is a dependency of the
// reference, not of this library, so the façade works on POD mirror types.  When Eigen is available
// (`__has_include(<eigen3/Eigen/Dense>)`) the adapter section at the bottom adds overloads with the
// reference's exact argument types (Eigen::Vector3d / Quaterniond / MatrixXd).
//
// Same call pattern as the reference's drivers (eqf_vio/src/main.cpp:111-140):
//     eqvio::VIOFilter filter(settings);
//     filter.processIMUData(imu);  ...  filter.processVisionData(meas);  auto xi = filter.stateEstimate();
// Behavioural notes kept from the reference: the class is move-only (it owns a handle the way the
// reference owns a unique_ptr<Settings>), void returns, silent skips on dt <= 0 / first sample / empty
// measurement (VIOFilter.cpp:147-152, 235-236, 258-259).  Errors the reference signals by assert or by
// throwing std::domain_error (SO3.cpp:160) surface as eqvio::Error.
#pragma once
#include <array>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../eqvio.h"

namespace eqvio {

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string& where) : std::runtime_error(where + ": " + eqvio_status_string(st)), status(st) {}
};

// ---- POD mirrors of the reference's value types ----
struct Vec3 { double x = 0, y = 0, z = 0; };
struct Quat { double w = 1, x = 0, y = 0, z = 0; };
struct Pose { Vec3 x; Quat R; };                       // SE3 (libs/core/include/SE3.h)
struct Point3d { Vec3 p; int id = -1; };               // VIOState.h:38-41
struct IMUVelocity {                                   // IMUVelocity.h:24-38
    double stamp = 0;
    Vec3 omega, accel;
};
struct VisionMeasurement {                             // VisionMeasurement.h:24-28
    double stamp = 0;
    int numberOfBearings = 0;
    std::vector<Point3d> bearings;                     // unit vectors, ascending id
};
struct VIOState {                                      // VIOState.h:51-60
    Pose pose;
    Vec3 velocity;
    std::vector<Point3d> bodyLandmarks;
    Pose cameraOffset;
};

class VIOFilter {
  public:
    // VIOFilter::Settings (VIOFilterSettings.h:28-54): the POD eqvio_settings_t has the same fields.
    struct Settings : eqvio_settings_t {
        Settings() { eqvio_settings_default(this); }
    };
    std::unique_ptr<Settings> settings;                // VIOFilter.h:67 (public in the reference)

    VIOFilter() = default;                             // like the reference: unusable until given Settings
    explicit VIOFilter(const Settings& s, int device = 0) : settings(new Settings(s)) {
        check(eqvio_create(settings.get(), device, &h_), "eqvio_create");
    }
    VIOFilter(VIOFilter&& o) noexcept : settings(std::move(o.settings)), h_(o.h_) { o.h_ = nullptr; }
    VIOFilter& operator=(VIOFilter&& o) noexcept {     // the ROS node move-assigns (eqf_vio_ros_node.cpp:59)
        if (this != &o) { release(); settings = std::move(o.settings); h_ = o.h_; o.h_ = nullptr; }
        return *this;
    }
    VIOFilter(const VIOFilter&) = delete;
    VIOFilter& operator=(const VIOFilter&) = delete;
    ~VIOFilter() { release(); }

    void reset() { check(eqvio_reset(h_), "eqvio_reset"); }                                    // VIOFilter.h:75
    void setInertialPoints(const std::vector<Point3d>& pts) {                                  // VIOFilter.h:79
        std::vector<int> ids(pts.size());
        std::vector<double> p(3 * pts.size());
        for (size_t i = 0; i < pts.size(); ++i) { ids[i] = pts[i].id; p[3 * i] = pts[i].p.x; p[3 * i + 1] = pts[i].p.y; p[3 * i + 2] = pts[i].p.z; }
        check(eqvio_set_inertial_points(h_, (int)pts.size(), ids.data(), p.data()), "eqvio_set_inertial_points");
    }
    void processIMUData(const IMUVelocity& v) {                                                // VIOFilter.h:80
        const double om[3] = {v.omega.x, v.omega.y, v.omega.z}, ac[3] = {v.accel.x, v.accel.y, v.accel.z};
        check(eqvio_process_imu(h_, v.stamp, om, ac), "eqvio_process_imu");
    }
    void processVisionData(const VisionMeasurement& m) {                                       // VIOFilter.h:81
        ids_.resize(m.bearings.size());
        y_.resize(3 * m.bearings.size());
        for (size_t i = 0; i < m.bearings.size(); ++i) {
            ids_[i] = m.bearings[i].id;
            y_[3 * i] = m.bearings[i].p.x; y_[3 * i + 1] = m.bearings[i].p.y; y_[3 * i + 2] = m.bearings[i].p.z;
        }
        check(eqvio_process_vision(h_, m.stamp, (int)ids_.size(), ids_.data(), y_.data()), "eqvio_process_vision");
    }
    double getTime() const { double t; check(eqvio_get_time(h_, &t), "eqvio_get_time"); return t; }     // VIOFilter.h:84
    VIOState stateEstimate() const {                                                           // VIOFilter.h:85
        int n = 0;
        check(eqvio_get_num_landmarks(h_, &n), "eqvio_get_num_landmarks");
        double pose[7], vel[3], cam[7];
        std::vector<int> ids(n ? n : 1);
        std::vector<double> lm(3 * (n ? n : 1));
        check(eqvio_get_state(h_, pose, vel, cam, &n, n, ids.data(), lm.data()), "eqvio_get_state");
        VIOState s;
        s.pose = to_pose(pose); s.cameraOffset = to_pose(cam);
        s.velocity = {vel[0], vel[1], vel[2]};
        s.bodyLandmarks.resize(n);
        for (int i = 0; i < n; ++i) { s.bodyLandmarks[i].id = ids[i]; s.bodyLandmarks[i].p = {lm[3 * i], lm[3 * i + 1], lm[3 * i + 2]}; }
        return s;
    }
    // Column-major n x n, n = 11 + 3N (Eigen::MatrixXd layout)                                 // VIOFilter.h:86
    std::vector<double> stateCovariance(int* n_out = nullptr) const {
        int N = 0;
        check(eqvio_get_num_landmarks(h_, &N), "eqvio_get_num_landmarks");
        const int n = EQVIO_SIGMA_BASE_SIZE + 3 * N;
        std::vector<double> S((size_t)n * n);
        check(eqvio_get_covariance(h_, S.data(), n), "eqvio_get_covariance");
        if (n_out) *n_out = n;
        return S;
    }
    std::array<double, 6> inputBias() const { std::array<double, 6> b; check(eqvio_get_bias(h_, b.data()), "eqvio_get_bias"); return b; }
    eqvio_handle_t handle() const { return h_; }

    // operator<< of the reference (VIOFilter.cpp:311-341): xi0, X, N, landmarks, Q_i, Sigma row-major on one line
    friend std::ostream& operator<<(std::ostream& os, const VIOFilter& f) {
        int N = 0;
        check(eqvio_get_num_landmarks(f.h_, &N), "eqvio_get_num_landmarks");
        std::vector<double> d(eqvio_snapshot_size(N));
        check(eqvio_get_snapshot(f.h_, d.data(), d.size()), "eqvio_get_snapshot");
        os << d[26] << ", " << d[27] << ", " << d[28] << ", " << d[22] << ", " << d[23] << ", " << d[24] << ", " << d[25] << ", ";
        os << d[29] << ", " << d[30] << ", " << d[31] << ", ";
        os << d[43] << ", " << d[44] << ", " << d[45] << ", " << d[39] << ", " << d[40] << ", " << d[41] << ", " << d[42] << ", ";
        os << d[46] << ", " << d[47] << ", " << d[48] << ", " << N;
        const double* L = d.data() + EQVIO_SNAPSHOT_HEADER;
        for (int i = 0; i < N; ++i, L += EQVIO_SNAPSHOT_PER_LANDMARK) {
            os << ", " << (int)L[0] << ", " << L[1] << ", " << L[2] << ", " << L[3];
            os << ", " << L[4] << ", " << L[5] << ", " << L[6] << ", " << L[7] << ", " << L[8];
        }
        const int n = EQVIO_SIGMA_BASE_SIZE + 3 * N;
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c) os << ", " << L[r + (size_t)n * c];
        return os;
    }

  private:
    eqvio_handle_t h_ = nullptr;
    std::vector<int> ids_;
    std::vector<double> y_;
    static void check(int st, const char* where) { if (st < 0) throw Error(st, where); }
    static Pose to_pose(const double* p) { Pose P; P.x = {p[0], p[1], p[2]}; P.R = {p[3], p[4], p[5], p[6]}; return P; }
    void release() { if (h_) { eqvio_destroy(h_); h_ = nullptr; } }
};

}  // namespace eqvio

// ---- optional adapter with the reference's Eigen argument types ----
#if defined(__has_include)
#if __has_include(<eigen3/Eigen/Dense>)
#include <eigen3/Eigen/Dense>
namespace eqvio {
inline IMUVelocity makeIMUVelocity(double stamp, const Eigen::Vector3d& omega, const Eigen::Vector3d& accel) {
    IMUVelocity v; v.stamp = stamp; v.omega = {omega.x(), omega.y(), omega.z()}; v.accel = {accel.x(), accel.y(), accel.z()}; return v;
}
inline Eigen::MatrixXd stateCovarianceEigen(const VIOFilter& f) {
    int n = 0;
    std::vector<double> S = f.stateCovariance(&n);
    return Eigen::Map<Eigen::MatrixXd>(S.data(), n, n);
}
}  // namespace eqvio
#endif
#endif
