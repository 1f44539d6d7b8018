// Minimal C++ caller with the reference's call pattern (eqf_vio/src/main.cpp:111-140) on the façade.
//   g++ -std=c++17 -I include examples/replay_minimal.cpp -L eqf_vio_b200/csrc -leqvio_b200 -Wl,-rpath,eqf_vio_b200/csrc
#include <cmath>
#include <cstdio>

#include "eqvio/VIOFilter.hpp"

int main() {
    eqvio::VIOFilter::Settings s;
    s.initialPointVariance = 100.0;
    s.measurementVariance = 0.003;
    try {
        eqvio::VIOFilter filter(s);
        eqvio::IMUVelocity imu;
        imu.accel = {0.3, -0.2, 9.7};  // not exactly +z: the gravity chart pole (reference VIOState.cpp:243) is singular for a perfectly level start
        for (int k = 0; k < 22; ++k) {
            imu.stamp = 0.005 * k;
            filter.processIMUData(imu);
            if (k % 10 == 5) {
                eqvio::VisionMeasurement m;
                m.stamp = imu.stamp + 0.0025;
                for (int i = 0; i < 6; ++i) {
                    eqvio::Point3d y;
                    y.id = i;
                    const double a = 0.3 * i;
                    y.p = {0.5 * std::cos(a), 0.5 * std::sin(a), std::sqrt(0.75)};
                    m.bearings.push_back(y);
                }
                m.numberOfBearings = (int)m.bearings.size();
                filter.processVisionData(m);
                eqvio::VIOState xi = filter.stateEstimate();
                std::printf("t=%.4f  N=%zu  pos=(%.4f %.4f %.4f)\n", filter.getTime(), xi.bodyLandmarks.size(), xi.pose.x.x, xi.pose.x.y, xi.pose.x.z);
            }
        }
    } catch (const eqvio::Error& e) {
        std::printf("eqvio error %d: %s\n", e.status, e.what());
        return e.status == EQVIO_ERR_NO_DEVICE ? 0 : 1;
    }
    return 0;
}
