// filter_kernels.cuh — device-side state of one filter session and the O(N) kernels around the
// dense Sigma contractions.  Reference counterparts are cited at each kernel in filter_kernels.cu.
#pragma once
#include "eqvio_math.cuh"

namespace eqvio {

// Error bits raised by kernels (read back at synchronisation points).
enum : int { FLAG_SINGULAR = 1, FLAG_NAN = 2, FLAG_NOT_SPD = 4 };

// Persistent state shared by all kernels of a session (VIOFilter members, VIOFilter.h:43-55).
struct BaseState {
    double bias[6];      // inputBias
    double curOmega[3], curAccel[3];  // currentVelocity (unbiased)
    double accOmega[3], accAccel[3];  // accumulatedVelocity
    Se3 pose0;           // xi0.pose
    V3 vel0;             // xi0.velocity
    Se3 cam;             // xi0.cameraOffset
    Se3 XA;              // X.A
    V3 Xw;               // X.w
    double pose_record[8];  // t, x y z, qw qx qy qz of stateEstimate().pose after the last update
    int flags;
    int pad;
};

// Per-landmark state, structure-of-arrays over a capacity `cap`:
//   q0x q0y q0z (xi0.bodyLandmarks[i].p), Qw Qx Qy Qz Qa (X.Q[i])
struct Landmarks {
    double* base;  // 8 * cap doubles
    int cap;
    __host__ __device__ double* q0(int c) const { return base + (size_t)c * cap; }
    __host__ __device__ double* Q(int c) const { return base + (size_t)(3 + c) * cap; }
};
static const int LM_FIELDS = 8;

// Quantities derived once per step from the pre-step state and broadcast to the per-feature warps.
struct StepScratch {
    // Riccati blocks
    M3 RICt_RAt;   // R_IC^T R_A^T                         (EqFMatrices.cpp:292-296)
    M3 RT_IC;      // R_IC^-1 as a matrix                  (EqFMatrices.cpp:371)
    M3 RT_IC_sx;   // R_IC^-1 [x_IC]x                      (EqFMatrices.cpp:376)
    V3 vC;         // camera-frame linear velocity, mean omega (EqFMatrices.cpp:302-304)
    double T;      // accumulated time of this Riccati step
    double Tpp[2]; // the same, by tick parity: read by the Riccati GEMM epilogue while the next tick's k_step_prepare runs
    double stamp;  // stamp of the step being processed (pose record of the vision update)
    double Rd[6];
    // state propagate
    Se3 camInv;    // SE3Exp(-dt U_C)                      (VIOGroup.cpp:229-230)
    V3 omC, vCcur; // U_C for the current velocity (continuous lift, VIOGroup.cpp:191-199)
    double dt;
    // update / lift
    V3 eta0n;          // normalised gravity direction of xi0
    double DUF[6];     // KPerp * DeltaU                   (EqFMatrices.cpp:212)
    double KPara[6][4];
    double AdP0[6][6];
    Quat RC;           // R_Phat * R_IC                    (EqFMatrices.cpp:209)
    M3 RCt;            // its inverse as a matrix
    Se3 PT;            // P_hat * T_IC
    Se3 DeltaA;        // lifted SE(3) increment
    V3 Deltaw;
};

struct ImuArgs {
    double omega[3], accel[3];  // raw sample (biased)
    double stamp;
    double dt;          // newTime - currentTime (valid when do_integrate)
    double T;           // accumulatedTime including dt (valid when do_riccati)
    int do_init, do_integrate, do_riccati, do_latch;
    int discrete_lift;
    int parity;         // which of the two F / W / T buffers this tick writes
};

struct RiccatiOut {
    double* F;   // ld x (>= n16+16), dense F = I + T*A_b in [0,n) x [0,n), columns [n16, n16+6) = B_b
    double* W;   // same shape; only columns [n16, n16+6) written here: T * B_b * diag(Rd)
    double* Bb;  // ld x 8 : B_b (n x 6)
    int ld, n, n16;
};

}  // namespace eqvio
