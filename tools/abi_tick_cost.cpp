// Host cost of the C ABI's IMU tick, without Python in the way (run under gpurun):
//   g++ -O2 -std=c++17 -I include tools/abi_tick_cost.cpp -L eqf_vio_b200/csrc -leqvio_b200 -Wl,-rpath,$PWD/eqf_vio_b200/csrc -o gpurun_out/abi_tick_cost
// Prints, per N: host time per eqvio_process_imu call (enqueue only) and the realised time per tick including the
// final synchronisation (whichever of host / device is slower).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "eqvio.h"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    std::vector<int> Ns;
    for (int i = 1; i < argc; ++i) Ns.push_back(atoi(argv[i]));
    if (Ns.empty()) Ns = {64, 256, 512};
    for (int N : Ns) {
        eqvio_settings_t s;
        eqvio_settings_default(&s);
        s.initialPointVariance = 100.0; s.initialSceneDepth = 8.0; s.measurementVariance = 0.003; s.outlierThreshold = 1e9;
        s.velOmegaVariance = s.velAccelVariance = 1e-4;
        eqvio_handle_t h;
        if (eqvio_create(&s, 0, &h) != EQVIO_OK) { printf("no device\n"); return 0; }
        std::mt19937 rng(7);
        std::uniform_real_distribution<double> U(-0.5, 0.5);
        std::vector<int> ids(N);
        std::vector<double> y(3 * N);
        for (int i = 0; i < N; ++i) {
            ids[i] = i;
            double a = U(rng), b = U(rng), c = 1.0, r = std::sqrt(a * a + b * b + c * c);
            y[3 * i] = a / r; y[3 * i + 1] = b / r; y[3 * i + 2] = c / r;
        }
        const double om[3] = {0.01, -0.02, 0.015}, ac[3] = {0.3, -0.2, 9.7};
        double t = 0.0;
        eqvio_process_imu(h, t, om, ac);
        t += 0.0025;
        eqvio_process_vision(h, t, N, ids.data(), y.data());
        t += 0.0025;
        for (int k = 0; k < 60; ++k, t += 0.005) eqvio_process_imu(h, t, om, ac);   // warm-up: graphs captured
        eqvio_synchronize(h);
        for (int rep = 0; rep < 3; ++rep) {
            const int K = 200;
            const double t0 = now();
            for (int k = 0; k < K; ++k, t += 0.005) eqvio_process_imu(h, t, om, ac);
            const double t1 = now();
            eqvio_synchronize(h);
            const double t2 = now();
            printf("N=%d: eqvio_process_imu host %.2f us/call, realised %.2f us/tick (%d ticks)\n", N, (t1 - t0) / K * 1e6, (t2 - t0) / K * 1e6, K);
        }
        eqvio_destroy(h);
    }
    return 0;
}
