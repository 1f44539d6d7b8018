// filter_kernels.cu — O(N) kernels of the EqF-VIO hot path (everything except the dense GEMMs):
//   k_step_prepare      base rows of F/B_b, derived step quantities, X.A / X.w propagate, velocity latch (4 warps, one piece each)
//   k_feature_step      one warp per feature: F / B_b landmark rows + Q_i propagate
//   k_build_C_delta     one warp per feature: 2x3 C block + innovation delta_i
//   k_gemv              gamma = K delta
//   k_lift_prepare / k_lift_features / k_lift_solve / k_lift_apply   bundleLift + discrete lift + X <- Delta X
//   k_chain_block      diagonal-block link of the blocked Schur eliminations (S^-1, Sigma_sub^-1): look-ahead corner, LU, L^-1, U^-1
//   k_lift_rsolve      Ym^T Sigma_sub^-1 by a multi-CTA wavefront back-substitution
//   bookkeeping: outlier flags, Sigma / landmark compaction, median depth + landmark append
#include <cstdlib>
#include "filter_kernels.cuh"
#include "kernels_api.cuh"

namespace eqvio {

#define GRAV 9.81  // GRAVITY_CONSTANT, eqf_vio/include/eqf_vio/IMUVelocity.h:22

__device__ __forceinline__ V3 ld3(const double* p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(double* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

__device__ __forceinline__ Sot3 load_Q(const Landmarks& L, int i) {
    Sot3 Q;
    Q.R.w = L.Q(0)[i]; Q.R.x = L.Q(1)[i]; Q.R.y = L.Q(2)[i]; Q.R.z = L.Q(3)[i]; Q.a = L.Q(4)[i];
    return Q;
}
__device__ __forceinline__ void store_Q(const Landmarks& L, int i, Sot3 Q) {
    L.Q(0)[i] = Q.R.w; L.Q(1)[i] = Q.R.x; L.Q(2)[i] = Q.R.y; L.Q(3)[i] = Q.R.z; L.Q(4)[i] = Q.a;
}
__device__ __forceinline__ V3 load_q0(const Landmarks& L, int i) { return v3(L.q0(0)[i], L.q0(1)[i], L.q0(2)[i]); }

// ------------------------------------------------------------------------------------------------
// k_step_prepare — one CTA, four working threads.  processIMUData / integrateUpToTime bookkeeping that is O(1):
//   VIOFilter.cpp:120-131 (unbias, initialise, latch), :154-155 (accumulate), :162-185 base rows of
//   A/B (EqFMatrices.cpp:289, 364-368), :196-203 with liftVelocityDiscrete (VIOGroup.cpp:209-243)
//   / liftVelocity (VIOGroup.cpp:178-207) for the SE(3) x R^3 part, X <- X * lift (VIOGroup.cpp:92-97).
// ------------------------------------------------------------------------------------------------
// The work is O(1) but ~6000 dependent fp64 instructions; run by one thread it took 12.8 us, the largest item of an
// IMU tick at small N.  Four warps (lane 0 of each) take one independent piece each — all read the pre-step state
// first, a barrier separates those reads from the writes of the propagate:
//   warp 0: accumulators, scalars for the kernels behind (T, dt, stamp), A0 gravity block, velocity latch
//   warp 1: base rows of B_b / F / W (EqFMatrices.cpp:364-368)
//   warp 2: R_IC^T R_A^T, v_C, R_IC^-1, R_IC^-1 [x_IC]x for the per-feature blocks
//   warp 3: X.A / X.w propagate and the camera-frame increment of the feature propagate
__global__ void __launch_bounds__(128) k_step_prepare(BaseState* st, StepScratch* sc, ImuArgs a, RiccatiOut ro) {
    const int role = threadIdx.x >> 5;
    const bool lead = (threadIdx.x & 31) == 0;
    if (a.do_init) {  // initialiseFromIMUData, VIOFilter.cpp:133-144 (first sample only)
        if (threadIdx.x == 0) {
            int sing0 = 0;
            const V3 ua0 = v3(a.accel[0] - st->bias[3], a.accel[1] - st->bias[4], a.accel[2] - st->bias[5]);
            st->pose0 = se3_identity();
            st->vel0 = v3(0, 0, 0);
            st->pose0.R = so3_from_vectors(normalized(ua0), v3(0, 0, 1), &sing0);
            if (sing0) atomicOr(&st->flags, FLAG_SINGULAR);
        }
        __syncthreads();
    }
    int sing = 0;
    // ---- every warp reads the pre-step state it needs ----
    const V3 uo = v3(a.omega[0] - st->bias[0], a.omega[1] - st->bias[1], a.omega[2] - st->bias[2]);
    const V3 ua = v3(a.accel[0] - st->bias[3], a.accel[1] - st->bias[4], a.accel[2] - st->bias[5]);
    const V3 co = ld3(st->curOmega), ca = ld3(st->curAccel);
    V3 ao = ld3(st->accOmega) + co * a.dt, aa = ld3(st->accAccel) + ca * a.dt;
    const Se3 XA = st->XA, pose0 = st->pose0, cam = st->cam;
    const V3 Xw = st->Xw, vel0 = st->vel0;
    __syncthreads();
    if (!lead) return;
    if (role == 0) sc->stamp = a.stamp;
    if (a.do_integrate) {
        const V3 eta0 = rotate_inv(pose0.R, v3(0, 0, 1));          // projectToManifold, VIOState.cpp:90
        const V3 eta_hat = rotate_inv(XA.R, eta0);                 // VIOGroup.cpp:49
        const V3 v_hat = rotate_inv(XA.R, vel0 - Xw);              // VIOGroup.cpp:25,50
        if (a.do_riccati) {
            const int ld = ro.ld, n16 = ro.n16;
            double* F = ro.F; double* W = ro.W; double* Bb = ro.Bb;
            if (role == 0) {
                sc->T = a.T;
                sc->Tpp[a.parity & 1] = a.T;
                // A0[2:5,0:2] = -g * stereoSphereChartInvDiff(0, eta0)   (EqFMatrices.cpp:289)
                M32 D = stereo_sphere_chart_inv_diff(0.0, 0.0, eta0, &sing);
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 2; ++c) F[(8 + r) + (size_t)ld * (6 + c)] = (-D.m[r][c] * GRAV) * a.T;
            } else if (role == 1) {
                // B rows (EqFMatrices.cpp:364-368)
                const M3 RA = to_matrix(XA.R);
                M23 Dg = stereo_sphere_chart_diff(eta0, eta0, &sing);
                M3 RAse = RA * skew(eta_hat);
                double Bbase[5][6];
                for (int r = 0; r < 2; ++r)
                    for (int c = 0; c < 3; ++c) {
                        Bbase[r][c] = Dg.m[r][0] * RAse.m[0][c] + Dg.m[r][1] * RAse.m[1][c] + Dg.m[r][2] * RAse.m[2][c];
                        Bbase[r][3 + c] = 0.0;
                    }
                M3 RAsv = RA * skew(v_hat);
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) { Bbase[2 + r][c] = RAsv.m[r][c]; Bbase[2 + r][3 + c] = RA.m[r][c]; }
                for (int r = 0; r < 5; ++r)
                    for (int c = 0; c < 6; ++c) {
                        const double b = Bbase[r][c];
                        F[(6 + r) + (size_t)ld * c] = -b * a.T;                           // A_b = [[0,0],[-Bt,A0t]]; F = I + A_b T
                        Bb[(6 + r) + (size_t)ld * c] = b;
                        F[(6 + r) + (size_t)ld * (n16 + c)] = b;                          // [F | B_b]
                        W[(6 + r) + (size_t)ld * (n16 + c)] = a.T * (b * sc->Rd[c]);      // [F Sigma | T B_b R]
                    }
            } else if (role == 2) {
                const V3 om = ao * (1.0 / a.T);
                const M3 RA = to_matrix(XA.R);
                const M3 RIC = to_matrix(cam.R);
                sc->RICt_RAt = transpose(RIC) * transpose(RA);
                V3 omC, vC;
                se3_adjoint_apply(inverse(cam), om, v_hat, &omC, &vC);   // EqFMatrices.cpp:302-304
                sc->vC = vC;
                sc->RT_IC = to_matrix(q_inverse(cam.R));
                sc->RT_IC_sx = sc->RT_IC * skew(cam.x);
            }
            ao = v3(0, 0, 0); aa = v3(0, 0, 0);  // VIOFilter.cpp:192-193
        }
        if (role == 0) {
            st3(st->accOmega, ao); st3(st->accAccel, aa);
            sc->dt = a.dt;
        } else if (role == 3) {
            // state propagate of the SE(3) x R^3 part
            V3 omC, vC;
            se3_adjoint_apply(inverse(cam), co, v_hat, &omC, &vC);
            if (a.discrete_lift) {
                Se3 LA = se3_exp(co * a.dt, v_hat * a.dt);
                V3 inner = v_hat + a.dt * (-cross(co, v_hat) + ca - eta_hat * GRAV);
                V3 Lw = v_hat - rotate(LA.R, inner);
                sc->camInv = se3_exp(omC * (-a.dt), vC * (-a.dt));
                st->Xw = Xw + rotate(XA.R, Lw);
                st->XA = XA * LA;
            } else {
                sc->omC = omC; sc->vCcur = vC;
                V3 u = -ca + eta_hat * GRAV;
                Se3 EA = se3_exp(co * a.dt, v_hat * a.dt);
                st->Xw = Xw + rotate(XA.R, u * a.dt);
                st->XA = XA * EA;
            }
        }
    }
    if (role == 0 && a.do_latch) { st3(st->curOmega, uo); st3(st->curAccel, ua); }
    if (sing) atomicOr(&st->flags, FLAG_SINGULAR);
}

// ------------------------------------------------------------------------------------------------
// k_feature_step — one warp per feature.  Lanes 0..2 each build one 3x3 block, lane 3 propagates
// Q_i; the blocks go through a per-warp shared staging area so that the stores to F / B_b are issued
// by 27 + 9 lanes side by side (rows of one feature are contiguous in every column).
//   A_iv, A_ii : EqFStateMatrixA_euclid_impl, EqFMatrices.cpp:292-312
//   B_i        : EqFInputMatrixB_euclid_impl, EqFMatrices.cpp:371-377
//   F, B_b     : VIOFilter.cpp:177-185
//   Q_i <- Q_i * lift.Q_i : liftVelocityDiscrete VIOGroup.cpp:231-240 / liftVelocity :192-199, product :104-107
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_feature_step(BaseState* st, const StepScratch* sc, Landmarks L, int N,
                                                      int do_riccati, int discrete_lift, RiccatiOut ro) {
    __shared__ double stage[8][40];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + wib;
    if (i >= N) return;
    double* sm = stage[wib];
    const Sot3 Q = load_Q(L, i);
    const V3 q0 = load_q0(L, i);
    const V3 qh = inverse(Q) * q0;  // VIOGroup.cpp:39-44
    if (do_riccati) {
        if (lane < 3) {
            const M3 Qhat = as_matrix3(Q);
            M3 blk;
            if (lane == 0) {
                blk = Qhat * (skew(qh) * sc->RT_IC + sc->RT_IC_sx);                       // B_i
            } else if (lane == 1) {
                blk = (Qhat * sc->RICt_RAt) * -1.0;                                       // A_iv
            } else {
                const V3 vC = sc->vC;
                M3 inner = skew(qh) * skew(vC) + outer(vC, qh) * -2.0 + outer(qh, vC);
                blk = ((Qhat * inner) * inverse(Qhat)) * (-1.0 / dot(qh, qh));            // A_ii
            }
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) sm[lane * 9 + c * 3 + r] = blk.m[r][c];      // column-major 3x3
        }
        __syncwarp();
        const double T = sc->T;
        const int ld = ro.ld, row0 = 11 + 3 * i;
        if (lane < 27) {
            const int b = lane / 9, e = lane % 9, c = e / 3, r = e % 3;
            const double v = sm[lane];
            double* F = ro.F;
            if (b == 0) {
                F[(row0 + r) + (size_t)ld * c] = -v * T;
                ro.Bb[(row0 + r) + (size_t)ld * c] = v;
                F[(row0 + r) + (size_t)ld * (ro.n16 + c)] = v;
                ro.W[(row0 + r) + (size_t)ld * (ro.n16 + c)] = T * (v * sc->Rd[c]);
            } else if (b == 1) {
                F[(row0 + r) + (size_t)ld * (8 + c)] = v * T;
            } else {
                F[(row0 + r) + (size_t)ld * (row0 + c)] = v * T + (r == c ? 1.0 : 0.0);
            }
        }
    }
    if (lane == 3) {
        int sing = 0;
        Sot3 dQ;
        if (discrete_lift) {
            const V3 p1 = sc->camInv * qh;
            dQ.R = so3_from_vectors(normalized(p1), normalized(qh), &sing);
            dQ.a = norm(qh) / norm(p1);
        } else {
            const double n2 = dot(qh, qh), dt = sc->dt;
            const V3 w = sc->omC + (skew(qh) * sc->vCcur) / n2;
            dQ = sot3_exp(w * dt, dt * (dot(qh, sc->vCcur) / n2));
        }
        store_Q(L, i, Q * dQ);
        if (sing) atomicOr(&st->flags, FLAG_SINGULAR);
    }
}

// ------------------------------------------------------------------------------------------------
// k_build_C_delta — one warp per feature (lane 0 computes, lanes 0..5 / 0..1 store).
//   y0 = measureSystemState(xi0)            VIOState.cpp:58-70
//   ye = outputGroupAction(X^-1, y)         VIOGroup.cpp:71-90, 124-134
//   delta_i = stereoSphereChart(ye, y0)     VisionMeasurement.cpp:24-34
//   C0_i = 1/|q0| chartDiff(y0,y0)(I - y0 y0^T)   EqFMatrices.cpp:332-338; C = [0, C0] VIOFilter.cpp:272-273
// C is m x n column-major with leading dimension ldc; only the 2x3 block of feature i is written.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_build_C_delta(BaseState* st, Landmarks L, int N, const double* bearings,
                                                       double* C, int ldc, double* delta) {
    __shared__ double stage[8][8];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + wib;
    if (i >= N) return;
    double* sm = stage[wib];
    if (lane == 0) {
        int sing = 0;
        const V3 q0 = load_q0(L, i);
        const V3 y0 = normalized(q0);
        const Sot3 Q = load_Q(L, i);
        if (delta && bearings) {
            const V3 y = v3(bearings[3 * i], bearings[3 * i + 1], bearings[3 * i + 2]);
            const V3 ye = rotate(q_inverse(q_inverse(Q.R)), y);
            stereo_sphere_chart(ye, y0, sm + 6, &sing);
        }
        M23 D = stereo_sphere_chart_diff(y0, y0, &sing);
        M3 proj = m3_identity() + outer(y0, y0) * -1.0;
        const double sc = 1 / norm(q0);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                sm[c * 2 + r] = (sc * D.m[r][0]) * proj.m[0][c] + (sc * D.m[r][1]) * proj.m[1][c] + (sc * D.m[r][2]) * proj.m[2][c];
        if (sing) atomicOr(&st->flags, FLAG_SINGULAR);
    }
    __syncwarp();
    if (C && lane < 6) C[(2 * i + (lane & 1)) + (size_t)ldc * (11 + 3 * i + (lane >> 1))] = sm[lane];
    if (delta && bearings && lane < 2) delta[2 * i + lane] = sm[6 + lane];
}

// gamma = K delta  (VIOFilter.cpp:279).  K is n x m column-major.  It sits on the update's critical path between K and
// the Sigma update, so the column range is split over GEMV_SPLIT CTAs per 32-row block (one CTA per row block took 48 us
// at n = 1547: 128 dependent loads + adds per thread); each CTA leaves its partial sums in `part`, the CTA that arrives
// last at the row block's counter adds them in a fixed order (deterministic, unlike fp64 atomics) and re-arms the counter.
constexpr int GEMV_SPLIT = 8;
__global__ void __launch_bounds__(256) k_gemv(const double* K, int ldk, int n, int m, const double* x, double* y, double* part, int* cnt) {
    __shared__ double red[8][33];
    __shared__ int s_last;
    const int rx = threadIdx.x & 31, cy = threadIdx.x >> 5;
    const int r = blockIdx.x * 32 + rx;
    const int per = (m + GEMV_SPLIT - 1) / GEMV_SPLIT, c_lo = blockIdx.y * per, c_hi = min(m, c_lo + per);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (r < n) {
        int c = c_lo + cy;
        for (; c + 24 < c_hi; c += 32) {
            s0 += K[r + (size_t)ldk * c] * x[c];
            s1 += K[r + (size_t)ldk * (c + 8)] * x[c + 8];
            s2 += K[r + (size_t)ldk * (c + 16)] * x[c + 16];
            s3 += K[r + (size_t)ldk * (c + 24)] * x[c + 24];
        }
        for (; c < c_hi; c += 8) s0 += K[r + (size_t)ldk * c] * x[c];
    }
    red[cy][rx] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (cy == 0 && r < n) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][rx];
        part[(size_t)blockIdx.y * n + r] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&cnt[blockIdx.x], 1) == GEMV_SPLIT - 1;
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (cy == 0 && r < n) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < GEMV_SPLIT; ++k) t += __ldcg(&part[(size_t)k * n + r]);
            y[r] = t;
        }
        if (threadIdx.x == 0) cnt[blockIdx.x] = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// Innovation lift.  gamma is the biased innovation (n entries): gamma[0:6] bias, gamma[6:] EqF part.
// k_lift_prepare: O(1) parts of bundleLift (EqFMatrices.cpp:173-213): DeltaU default, KPara, KPerp,
//   R_C, Ad(P0), P_hat T_IC.
// ------------------------------------------------------------------------------------------------
__global__ void k_lift_prepare(BaseState* st, StepScratch* sc, const double* gamma) {
    // gamma == nullptr: the parts that do not depend on the innovation (they ride on the Sigma_sub
    // elimination, which runs concurrently with the S / K / gamma chain); gamma != nullptr: DUF only.
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int sing = 0;
    if (gamma == nullptr) {
        const V3 eta0 = normalized(rotate_inv(st->pose0.R, v3(0, 0, 1)));
        sc->eta0n = eta0;
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 4; ++c) sc->KPara[r][c] = 0.0;
        sc->KPara[0][0] = eta0.x; sc->KPara[1][0] = eta0.y; sc->KPara[2][0] = eta0.z;
        sc->KPara[3][1] = sc->KPara[4][2] = sc->KPara[5][3] = 1.0;
        const Se3 Phat = st->pose0 * st->XA;
        sc->RC = Phat.R * st->cam.R;
        sc->RCt = to_matrix(q_inverse(sc->RC));
        sc->PT = Phat * st->cam;
        // Ad(P0), SE3.cpp:95-103
        M3 R = to_matrix(st->pose0.R), sxR = skew(st->pose0.x) * R;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                sc->AdP0[r][c] = R.m[r][c]; sc->AdP0[r][3 + c] = 0.0;
                sc->AdP0[3 + r][c] = sxR.m[r][c]; sc->AdP0[3 + r][3 + c] = R.m[r][c];
            }
        return;
    }
    const V3 eta0 = sc->eta0n;
    const double* g = gamma + 6;
    M32 D = stereo_sphere_chart_inv_diff(0.0, 0.0, eta0, &sing);
    V3 t = v3(D.m[0][0] * g[0] + D.m[0][1] * g[1], D.m[1][0] * g[0] + D.m[1][1] * g[1], D.m[2][0] * g[0] + D.m[2][1] * g[1]);
    V3 Om = -(skew(eta0) * t);  // :187
    // DUF = KPerp * DeltaU with KPerp = diag(I - eta eta^T, 0)  (:201-206, :212)
    M3 P = m3_identity() + outer(eta0, eta0) * -1.0;
    V3 duf = P * Om;
    sc->DUF[0] = duf.x; sc->DUF[1] = duf.y; sc->DUF[2] = duf.z; sc->DUF[3] = sc->DUF[4] = sc->DUF[5] = 0.0;
    if (sing) atomicOr(&st->flags, FLAG_SINGULAR);
}

// k_lift_features: per feature (EqFMatrices.cpp:218-236)
//   M_i = [-[p_i]x, I] Ad(P0) KPara,  obs_i = -R_C Q_i^-1 gamma_qi - [-[p_i]x, I] Ad(P0) DUF,  D_i = Qhat_i R_C^T.
// gamma == nullptr (innovation-independent part): Ym = D M (p x 4) as the border of the Schur problem
//   [[Sigma_sub, Ym], [Ym^T, 0]]:  Aug[pb + c, 5 + 3i + r] = Aug[5 + 3i + r, pb + c] = (D_i M_i)[r][c];
//   the first five rows of Ym are zero.
// gamma != nullptr: yo = D obs (p entries, yo[0:5] = 0) into `yo`.
__global__ void __launch_bounds__(128) k_lift_features(const StepScratch* sc, Landmarks L, int N, const double* gamma,
                                                       double* Aug, int lda, int pb, double* yo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (gamma == nullptr) {
        if (i < 5)
            for (int c = 0; c < 4; ++c) {
                Aug[(pb + c) + (size_t)lda * i] = 0.0;
                Aug[i + (size_t)lda * (pb + c)] = 0.0;
                if (i < 4) Aug[(pb + i) + (size_t)lda * (pb + c)] = 0.0;
            }
    } else {
        if (i < 5) yo[i] = 0.0;
        if (i < 16 && 5 + 3 * N + i < pb) yo[5 + 3 * N + i] = 0.0;  // identity-padded tail of the Schur problem
    }
    if (i >= N) return;
    const Sot3 Q = load_Q(L, i);
    const V3 q0 = load_q0(L, i);
    const V3 qh = inverse(Q) * q0;
    const V3 pH = sc->PT * qh;
    // pHatMat * AdP0 (3 x 6)
    M3 nsp = skew(pH) * -1.0;
    double pmAd[3][6];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c)
            pmAd[r][c] = nsp.m[r][0] * sc->AdP0[0][c] + nsp.m[r][1] * sc->AdP0[1][c] + nsp.m[r][2] * sc->AdP0[2][c] + sc->AdP0[3 + r][c];
    const M3 Dm = as_matrix3(Q) * sc->RCt;
    if (gamma == nullptr) {
        double Mi[3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                double m = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) m += pmAd[r][k] * sc->KPara[k][c];
                Mi[r][c] = m;
            }
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double yv = Dm.m[r][0] * Mi[0][c] + Dm.m[r][1] * Mi[1][c] + Dm.m[r][2] * Mi[2][c];
                Aug[(pb + c) + (size_t)lda * (5 + 3 * i + r)] = yv;
                Aug[(5 + 3 * i + r) + (size_t)lda * (pb + c)] = yv;
            }
    } else {
        const double* gq = gamma + 11 + 3 * i;
        const V3 alpha = -rotate(sc->RC, inverse(Q) * v3(gq[0], gq[1], gq[2]));
        double ob[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            double sacc = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) sacc += pmAd[r][k] * sc->DUF[k];
            ob[r] = v3_get(alpha, r) - sacc;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) yo[5 + 3 * i + r] = Dm.m[r][0] * ob[0] + Dm.m[r][1] * ob[1] + Dm.m[r][2] * ob[2];
    }
}

// k_lift_rsolve — R^T = Ym^T Sigma_sub^-1 (4 x pb), the gamma-INDEPENDENT factor of the normal equations'
// right-hand side M^T W obs = (Ym^T Sigma_sub^-1) (D obs)  (EqFMatrices.cpp:239-242).  The Schur elimination
// leaves G = Ym^T U^-1 in the border rows pb..pb+3 and the unit-lower factor L in the sub-diagonal blocks of
// Aug (with every L_jj^-1 kept), so R^T = G L^-1 is a block back-substitution from the right:
//     R_j = (G_j - sum_{i > j} R_i L_ij) L_jj^-1,      j = last block ... 0.
// One CTA per 64-wide block, all launched together: CTA c owns block j = nblk-1-c, streams its L_ij tiles through
// shared memory and consumes R_i in the order they are published, so the only serial part is one 4 x 64 x 64
// product per link instead of a whole single-CTA sweep.  CTAs only wait on lower-numbered CTAs, which the hardware
// dispatches first: no co-residency requirement.
// Hand-over: the data validates itself.  R^T is pre-filled with a NaN payload no computation produces
// (LIFT_RT_PENDING, by k_fill_u64 at the head of the lift chain); a producer stores its 8-byte entries with plain
// L2 stores, consumer thread (a, c) polls its own entry of R_i until it is no longer the sentinel.  One L2 write + one
// L2 read per link instead of store, fence, release-flag, acquire-poll, load (four round trips: 6.6 -> 3.4 us per link
// under the load of the covariance update's products).  A wait that gives up (~1 s) raises FLAG_NAN.
// Runs on the lift stream behind the elimination, before the innovation exists; once gamma is known
// M^T W obs = R^T yo is four dot products (k_lift_solve).
static constexpr unsigned long long LIFT_RT_PENDING = 0xFFF8DEADBEEF5A5Aull;
__global__ void k_fill_u64(unsigned long long* p, int n, unsigned long long v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void __launch_bounds__(256) k_lift_rsolve(const double* Aug, int lda, int pb, const double* LinvBlocks, double* Rt,
                                                     int* err) {
    extern __shared__ double sm_rs[];
    double* tiles = sm_rs;                    // three L_ij buffers, 64 x 64, ld 65: two tiles in flight (cp.async) under the product of a third
    double* linv = sm_rs + 3 * 64 * 65;       // L_jj^-1, loaded once up front
    double(*Ri)[4][64] = reinterpret_cast<double(*)[4][64]>(sm_rs + 4 * 64 * 65);   // R_i, double-buffered: one barrier per tile
    const int nblk = (pb + 63) >> 6, j = nblk - 1 - (int)blockIdx.x, j0 = j << 6, nbj = min(64, pb - j0);
    const int tid = threadIdx.x, a = tid >> 6, c = tid & 63;   // thread (a, c): entry (a, j0 + c) of R^T
    double acc = (c < nbj) ? Aug[(pb + a) + (size_t)lda * (j0 + c)] : 0.0;
    const unsigned tiles_s = (unsigned)__cvta_generic_to_shared(tiles);
    const int nt = nblk - 1 - j;   // tiles L_ij to consume, i = nblk-1 ... j+1 (tile t <-> i = nblk-1-t): known up front, so the
    // loads run two tiles ahead of the products whatever the state of the chain.  (One tile in flight, loaded through registers,
    // made a CTA's per-tile time — not the hand-over — the period of the whole wavefront: 6.2 us per link.)
    auto issue_tile = [&](int t) {
        if (t < nt) {
            const int i0 = (nblk - 1 - t) << 6, nbi = min(64, pb - i0);
            const unsigned dst = tiles_s + (unsigned)(t % 3) * (64 * 65 * 8);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int idx = tid + 256 * u, r = idx & 63, cc = idx >> 6;
                const bool in = r < nbi && cc < nbj;
                const double* src = in ? Aug + (i0 + r) + (size_t)lda * (j0 + cc) : Aug;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + 8u * (r + 65 * cc)), "l"(src), "r"(in ? 8 : 0) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");   // (an empty group past the last tile keeps the count below uniform)
    };
    issue_tile(0);
    issue_tile(1);
    {
        const double* Li = LinvBlocks + (size_t)j * 4096;
        for (int idx = tid; idx < 64 * 64; idx += 256) linv[(idx & 63) + 65 * (idx >> 6)] = Li[idx];
    }
    for (int t = 0; t < nt; ++t) {
        const int i0 = (nblk - 1 - t) << 6, nbi = min(64, pb - i0);
        const double* tile = tiles + (t % 3) * 64 * 65;
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // this thread's copies of tile t have landed
        {
            unsigned long long v = 0ull;
            if (c < nbi) {
                const unsigned long long* src = reinterpret_cast<const unsigned long long*>(Rt + (size_t)a * pb + i0 + c);
                long long spins = 0;
                do {
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
                    if (v == LIFT_RT_PENDING && ++spins > (1ll << 26)) { atomicOr(err, FLAG_NAN); v = 0ull; }
                } while (v == LIFT_RT_PENDING);
            }
            Ri[t & 1][a][c] = __longlong_as_double((long long)v);
        }
        __syncthreads();   // tile t and R_i are in shared memory; every thread is past the product of tile t-1
        issue_tile(t + 2);   // into the buffer of tile t-1
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll 4
        for (int q = 0; q < 64; q += 4) {
            p0 = fma(Ri[t & 1][a][q], tile[q + 65 * c], p0);
            p1 = fma(Ri[t & 1][a][q + 1], tile[q + 1 + 65 * c], p1);
            p2 = fma(Ri[t & 1][a][q + 2], tile[q + 2 + 65 * c], p2);
            p3 = fma(Ri[t & 1][a][q + 3], tile[q + 3 + 65 * c], p3);
        }
        acc -= (p0 + p1) + (p2 + p3);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // R_j = acc * L_jj^-1 (64 x 64, identity-padded)
    Ri[0][a][c] = acc;
    __syncthreads();
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll 4
    for (int q = 0; q < 64; q += 4) {
        p0 = fma(Ri[0][a][q], linv[q + 65 * c], p0);
        p1 = fma(Ri[0][a][q + 1], linv[q + 1 + 65 * c], p1);
        p2 = fma(Ri[0][a][q + 2], linv[q + 2 + 65 * c], p2);
        p3 = fma(Ri[0][a][q + 3], linv[q + 3 + 65 * c], p3);
    }
    if (c < nbj) {
        unsigned long long v = (unsigned long long)__double_as_longlong((p0 + p1) + (p2 + p3));
        if (v == LIFT_RT_PENDING) v = 0x7FF8000000000000ull;   // (cannot come out of arithmetic; kept for the proof that no consumer waits forever)
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(Rt + (size_t)a * pb + j0 + c), "l"(v) : "memory");
    }
}

// 4x4 Householder QR solve (EqFMatrices.cpp:240-242)
__device__ void qr_solve4(double A[4][4], double b[4], double x[4]) {
    for (int k = 0; k < 4; ++k) {
        double tail2 = 0;
        for (int r = k + 1; r < 4; ++r) tail2 += A[r][k] * A[r][k];
        const double c0 = A[k][k];
        if (tail2 == 0.0) continue;
        double beta = sqrt(c0 * c0 + tail2);
        if (c0 >= 0) beta = -beta;
        double v[4] = {0, 0, 0, 0};
        for (int r = k + 1; r < 4; ++r) v[r] = A[r][k] / (c0 - beta);
        v[k] = 1.0;
        const double tau = (beta - c0) / beta;
        for (int c = k; c < 4; ++c) {
            double s = 0;
            for (int r = k; r < 4; ++r) s += v[r] * A[r][c];
            s *= tau;
            for (int r = k; r < 4; ++r) A[r][c] -= s * v[r];
        }
        double s = 0;
        for (int r = k; r < 4; ++r) s += v[r] * b[r];
        s *= tau;
        for (int r = k; r < 4; ++r) b[r] -= s * v[r];
    }
    for (int k = 3; k >= 0; --k) {
        double s = b[k];
        for (int c = k + 1; c < 4; ++c) s -= A[k][c] * x[c];
        x[k] = s / A[k][k];
    }
}

// k_lift_solve — single thread.  mode use_lift: M^T W M (W = D^T Sigma_sub^-1 D, EqFMatrices.cpp:239-242)
// is the negated bottom-right 4 x 4 block left by the Schur elimination, M^T W obs comes from
// R^T yo with R^T from k_lift_rsolve; 4x4 QR solve, DeltaU = DUF + KPara x (:243), then
// the SE(3) x R^3 part of the lift and X <- Delta X, bias update:
//   discrete   liftTotalSpaceInnovationDiscrete EqFMatrices.cpp:254-259
//   continuous VIOExp(liftTotalSpaceInnovation)  EqFMatrices.cpp:69-78, VIOGroup.cpp:245-248
// use_lift = 0 (useInnovationLift = false): VIOExp(liftInnovation(gamma, xi0)) EqFMatrices.cpp:35-48.
// Then VIOFilter.cpp:295-296 and the pose record.
__global__ void __launch_bounds__(128) k_lift_solve(BaseState* st, StepScratch* sc, const double* gamma, const double* Aug,
                                                    int lda, int p, const double* Rt, long rt_rs, long rt_cs, double rt_sign,
                                                    const double* yo, int use_lift, int discrete, double* Gamma_out, int apply) {
    __shared__ double b4[4];
    const int tid = threadIdx.x;
    if (use_lift) {   // M^T W obs = R^T yo: warp a forms entry a
        const int a = tid >> 5;
        double sacc = 0.0;
        // R^T entry (a, col) = rt_sign * Rt[a * rt_rs + col * rt_cs]: row-major from k_lift_rsolve, or (negated) the border block
        // the elimination leaves when Sigma_sub was bordered by the identity (small N)
        for (int col = tid & 31; col < p; col += 32) sacc = fma(rt_sign * Rt[a * rt_rs + col * rt_cs], yo[col], sacc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
        if ((tid & 31) == 0) b4[a] = sacc;
    }
    __syncthreads();
    double G[4][5];
    if (use_lift && tid == 0)
        for (int a = 0; a < 4; ++a) {
            for (int b = 0; b < 4; ++b) G[a][b] = -Aug[(p + a) + (size_t)lda * (p + b)];
            G[a][4] = b4[a];
        }
    if (tid != 0) return;
    double DU[6];
    V3 Dw;
    const double* g = gamma + 6;
    if (use_lift) {
        double A[4][4], b[4], x[4];
        for (int a = 0; a < 4; ++a) { for (int c = 0; c < 4; ++c) A[a][c] = G[a][c]; b[a] = G[a][4]; }
        qr_solve4(A, b, x);
        for (int r = 0; r < 6; ++r) {
            double s = 0.0;
            for (int c = 0; c < 4; ++c) s += sc->KPara[r][c] * x[c];
            DU[r] = sc->DUF[r] + s;
        }
        if (Gamma_out)
            for (int r = 0; r < 6; ++r) Gamma_out[r] = DU[r];
        const V3 gv = v3(g[2], g[3], g[4]);
        if (discrete) {
            sc->DeltaA = se3_exp(v3(DU[0], DU[1], DU[2]), v3(DU[3], DU[4], DU[5]));
            Dw = st->vel0 - rotate(sc->DeltaA.R, st->vel0 + gv);
        } else {
            sc->DeltaA = se3_exp(v3(DU[0], DU[1], DU[2]), v3(DU[3], DU[4], DU[5]));
            Dw = -gv - skew(v3(DU[0], DU[1], DU[2])) * st->vel0;
        }
    } else {
        int sing = 0;
        const V3 eta0 = rotate_inv(st->pose0.R, v3(0, 0, 1));
        M32 D = stereo_sphere_chart_inv_diff(0.0, 0.0, eta0, &sing);
        V3 t = v3(D.m[0][0] * g[0] + D.m[0][1] * g[1], D.m[1][0] * g[0] + D.m[1][1] * g[1], D.m[2][0] * g[0] + D.m[2][1] * g[1]);
        V3 Om = -(skew(eta0) * t);
        sc->DeltaA = se3_exp(Om, v3(0, 0, 0));
        Dw = -v3(g[2], g[3], g[4]) - skew(Om) * st->vel0;
        if (sing) atomicOr(&st->flags, FLAG_SINGULAR);
    }
    if (!apply) return;  // kernel-level bundle-lift entry point only wants Gamma[0:6]
    sc->Deltaw = Dw;
    for (int k = 0; k < 6; ++k) st->bias[k] += gamma[k];                       // VIOFilter.cpp:295
    st->Xw = Dw + rotate(sc->DeltaA.R, st->Xw);                               // VIOGroup.cpp:96
    st->XA = sc->DeltaA * st->XA;                                             // VIOGroup.cpp:95
    const Se3 P = st->pose0 * st->XA;
    st->pose_record[0] = sc->stamp;  // written by k_step_prepare of this frame's integrateUpToTime
    st->pose_record[1] = P.x.x; st->pose_record[2] = P.x.y; st->pose_record[3] = P.x.z;
    st->pose_record[4] = P.R.w; st->pose_record[5] = P.R.x; st->pose_record[6] = P.R.y; st->pose_record[7] = P.R.z;
}

// k_lift_apply — thread per feature: Q_i <- DeltaQ_i * Q_i (VIOGroup.cpp:104-107) with
//   discrete: DeltaQ_i from q0 + Gamma_qi (EqFMatrices.cpp:264-270)
//   continuous: SOT3Exp(W_i), W_i = (-q0 x g / |q0|^2, -q0.g / |q0|^2) (EqFMatrices.cpp:50-63 / 80-93)
__global__ void __launch_bounds__(128) k_lift_apply(BaseState* st, Landmarks L, int N, const double* gamma, int discrete) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int sing = 0;
    const V3 q0 = load_q0(L, i);
    const double* gq = gamma + 11 + 3 * i;
    const V3 gv = v3(gq[0], gq[1], gq[2]);
    Sot3 dQ;
    if (discrete) {
        const V3 q1 = q0 + gv;
        dQ.R = so3_from_vectors(normalized(q1), normalized(q0), &sing);
        dQ.a = norm(q0) / norm(q1);
    } else {
        const double n2 = dot(q0, q0);
        dQ = sot3_exp(-cross(q0, gv) / n2, -dot(q0, gv) / n2);
    }
    store_Q(L, i, dQ * load_Q(L, i));
    if (sing) atomicOr(&st->flags, FLAG_SINGULAR);
}

// ------------------------------------------------------------------------------------------------
// Blocked Schur elimination by UNPIVOTED LU (stand-in for the reference's explicit inverses,
// VIOFilter.cpp:277 `S.inverse()` and EqFMatrices.cpp:239 `Sigma.inverse()`, which Eigen evaluates with
// PartialPivLU on the matrices AS THEY ARE, i.e. including their round-off-level asymmetry).
// The augmented buffer is [[A, Cc], [R, Z]] with A k x k; eliminating A leaves Z - R A^-1 Cc in the
// bottom-right block.  A is symmetric positive definite up to rounding, so no pivoting is needed;
// nothing here assumes symmetry (a Cholesky sweep would, and measurably departs from the reference:
// the asymmetry of S is ~1e-11 relative because C annihilates the large radial variance — see
// DESIGN.md "Numerical conditioning").
// ------------------------------------------------------------------------------------------------
#ifdef EQVIO_DEBUG_CLOCKS
__device__ long long g_dbg_clk[16];
#define DBG_CLK(i) do { if (threadIdx.x == 0) g_dbg_clk[i] = clock64(); } while (0)
#else
#define DBG_CLK(i) do { } while (0)
#endif
// ------------------------------------------------------------------------------------------------
// k_chain_block — the sequential link of the blocked Schur eliminations, one CTA of 16 warps per 64-wide
// diagonal block.  Everything that is a product runs on DMMA.8x8x4 from shared memory; only the 64 pivots
// themselves are scalar.
//   prologue (block j > 0): D_j = A[j,j] - L[j,j-1] U[j-1,j]  (the look-ahead "corner": the trailing update of
//             step j-1 skips this block, so the diagonal block never waits for that GEMM)
//   LU:       eight 8-wide sub-steps: LU of the 8 x 8 diagonal block (8 lanes of warp 0, rows in registers, pivot row
//             by shuffle), the 8-wide row / column panels by substitution (one thread per row / column), rank-8
//             trailing update on DMMA (each warp a 2 x 2 group of 8 x 8 blocks)
//   inverses: the eight 8 x 8 diagonal inverses of L and U at once (128 threads, substitution), then L^-1, U^-1
//             (64 x 64) by recursive doubling:
//             X21 = -L22^-1 (L21 L11^-1),  X12 = -U11^-1 (U12 U22^-1)   for block sizes 8, 16, 32
// Shared-memory matrices are column-major with leading dimension 68 (= 4 mod 16 doubles): every DMMA fragment
// load (A: row g, k t;  B: k t, column g) hits 16 distinct 8-byte slots per half-warp, and the global <-> shared
// copies are coalesced and conflict-free.
// ------------------------------------------------------------------------------------------------
constexpr int LDW = 68;
constexpr int CHAIN_WARPS = 16;

// Reciprocal to ~1 ulp without leaving the fp64 pipe: MUFU.RCP64H seed, one cubic and one quadratic Newton
// step (the sequence nvcc's own division uses, minus its range fix-ups: pivots are variances, far from the
// ends of the exponent range; a zero / non-finite pivot is caught by the caller's check).
__device__ __forceinline__ double pivot_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    e = fma(e, e, e);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

__device__ __forceinline__ void dmma_f64(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// One 8 x 8 output block by one warp:  C = (ACC ? C : 0) + sign * A (8 x K) B (K x 8); operands column-major in
// shared memory with leading dimension LDW; A points at (row0, k = 0), B at (k = 0, col0), C at (row0, col0).
// C may alias A or B (in-place panel products): every fragment is read before anything is written.
template <int K, bool ACC>
__device__ __forceinline__ void blk8(double* C, const double* A, const double* B, double sign, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double c0 = 0.0, c1 = 0.0;
    if (ACC) { c0 = C[g + (2 * t) * LDW]; c1 = C[g + (2 * t + 1) * LDW]; }
    double a[K / 4], b[K / 4];
#pragma unroll
    for (int q = 0; q < K / 4; ++q) { a[q] = sign * A[g + (4 * q + t) * LDW]; b[q] = B[(4 * q + t) + g * LDW]; }
#pragma unroll
    for (int q = 0; q < K / 4; ++q) dmma_f64(c0, c1, a[q], b[q]);
    __syncwarp();
    C[g + (2 * t) * LDW] = c0;
    C[g + (2 * t + 1) * LDW] = c1;
}

// C (8 mb x 8 nb blocks, mb, nb <= 8) = (ACC ? C : 0) + sign * A B.  The 16 warps form a 4 x 4 grid, warp (wi, wj)
// owns the 2 x 2 group of 8 x 8 output blocks at block row 2 wi, block column 2 wj: four independent DMMA chains per
// warp and each operand fragment is loaded once for two blocks.  No divisions; the kernel's code runs once per
// launch, so instruction fetch is part of the critical path and loops stay rolled wherever a register array does
// not force unrolling.
template <int K, bool ACC>
__device__ __forceinline__ void mm_smem(double* C, const double* A, const double* B, int mb, int nb, double sign, int warp, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const int bi0 = 2 * (warp & 3), bj0 = 2 * (warp >> 2);
    const int ni = min(2, mb - bi0), nj = min(2, nb - bj0);   // warp-uniform
    if (ni <= 0 || nj <= 0) return;
    const bool i1 = ni > 1, j1 = nj > 1;
    const double* A0 = A + 8 * bi0 + g;              // + (k) * LDW; second row block at + 8
    const double* B0 = B + (8 * bj0 + g) * LDW;      // + k;         second column block at + 8 * LDW
    double* C0 = C + 8 * bi0 + g + (8 * bj0 + 2 * t) * LDW;
    double c[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
    if (ACC) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if ((i == 0 || i1) && (j == 0 || j1)) {
                    c[i][j][0] = C0[8 * i + (8 * j) * LDW];
                    c[i][j][1] = C0[8 * i + (8 * j + 1) * LDW];
                }
    }
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        const double a0 = sign * A0[(k + t) * LDW], a1 = i1 ? sign * A0[8 + (k + t) * LDW] : 0.0;
        const double b0 = B0[k + t], b1 = j1 ? B0[k + t + 8 * LDW] : 0.0;
        dmma_f64(c[0][0][0], c[0][0][1], a0, b0);
        if (j1) dmma_f64(c[0][1][0], c[0][1][1], a0, b1);
        if (i1) dmma_f64(c[1][0][0], c[1][0][1], a1, b0);
        if (i1 && j1) dmma_f64(c[1][1][0], c[1][1][1], a1, b1);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if ((i == 0 || i1) && (j == 0 || j1)) {
                C0[8 * i + (8 * j) * LDW] = c[i][j][0];
                C0[8 * i + (8 * j + 1) * LDW] = c[i][j][1];
            }
}

// Sub-step pb of the 64 x 64 LU: LU of the 8 x 8 diagonal block (warp 0), the 8-wide panels by substitution (one
// thread per row / column), rank-8 update of the rest on DMMA.  rp[0..64) collects the pivot reciprocals.
__device__ __forceinline__ void lu_substep(int pb, double* S, double* rp, int tid, int warp, int lane, int& bad) {
    const int c0 = 8 * pb, rem = 7 - pb;   // rem: 8-blocks right of / below the diagonal block
    double* Sd = S + c0 + c0 * LDW;
    if (pb == 0) DBG_CLK(5);
    if (warp == 0) {
        // lane r (< 8; the other lanes mirror) keeps row r in registers, the pivot row travels by shuffle, every lane
        // forms the pivot reciprocal itself
        const int r = lane & 7;
        double a[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) a[c] = Sd[r + c * LDW];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double piv = __shfl_sync(0xffffffffu, a[k], k);
            bad |= !(fabs(piv) > 0.0);
            const double rk = pivot_rcp(piv);
            if (lane == k) rp[c0 + k] = rk;
            const double l = (r > k) ? a[k] * rk : 0.0;
#pragma unroll
            for (int c = k + 1; c < 8; ++c) {
                const double u = __shfl_sync(0xffffffffu, a[c], k);
                a[c] = fma(-l, u, a[c]);
            }
            if (r > k) a[k] = l;
        }
        __syncwarp();   // the mirror lanes have read the block before lanes 0..7 overwrite it
        if (lane < 8) {
#pragma unroll
            for (int c = 0; c < 8; ++c) Sd[r + c * LDW] = a[c];
        }
    }
    if (pb == 0) DBG_CLK(6);
    __syncthreads();
    if (pb == 0) DBG_CLK(7);
    if (rem > 0) {
        // panels, in place, by substitution: threads [0, 8 rem) one row each of L[c0+8.., c0..c0+8) = A U8^-1,
        // threads [64, 64 + 8 rem) one column each of U[c0..c0+8, c0+8..) = L8^-1 A
        if (tid < 8 * rem) {
            double* row = S + (c0 + 8 + tid) + c0 * LDW;
            double x[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] = row[c * LDW];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                x[c] *= rp[c0 + c];
#pragma unroll
                for (int q = c + 1; q < 8; ++q) x[q] = fma(-x[c], Sd[c + q * LDW], x[q]);
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) row[c * LDW] = x[c];
        } else if (tid >= 64 && tid < 64 + 8 * rem) {
            double* col = S + c0 + (c0 + 8 + (tid - 64)) * LDW;
            double x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = col[i];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int q = i + 1; q < 8; ++q) x[q] = fma(-Sd[q + i * LDW], x[i], x[q]);
#pragma unroll
            for (int i = 0; i < 8; ++i) col[i] = x[i];
        }
        if (pb == 0) DBG_CLK(8);
        __syncthreads();
        if (pb == 0) DBG_CLK(9);
        mm_smem<8, true>(S + (c0 + 8) + (c0 + 8) * LDW, S + (c0 + 8) + c0 * LDW, S + c0 + (c0 + 8) * LDW, rem, rem, -1.0, warp, lane);
        if (pb == 0) DBG_CLK(10);
        __syncthreads();
        if (pb == 0) DBG_CLK(11);
    }
}

// The eight 8 x 8 diagonal blocks of L^-1 and U^-1, all at once after the LU: 16 lanes per block, lanes 0..7 column
// `cc` of L8^-1 (unit lower, forward substitution), lanes 8..15 a column of U8^-1 on the index-reversed block
// (upper -> lower, diagonal 1 / pivot) — one code path for both.
__device__ __forceinline__ void diag_inverses(const double* S, double* LI, double* UI, const double* rp, int tid) {
    if (tid >= 128) return;
    const int blk = tid >> 4, c0 = 8 * blk, l16 = tid & 15;
    const bool up = l16 >= 8;
    const int cc = l16 & 7;
    const double* Sd = S + c0 + c0 * LDW;
    double x[8], sacc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sacc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double d = up ? rp[c0 + 7 - i] : 1.0;
        x[i] = (i < cc) ? 0.0 : (i == cc ? d : -(d * sacc[i]));
#pragma unroll
        for (int q = i + 1; q < 8; ++q) {
            const double mqi = up ? Sd[(7 - q) + (7 - i) * LDW] : Sd[q + i * LDW];
            sacc[q] = fma(mqi, x[i], sacc[q]);
        }
    }
    double* XI = up ? UI : LI;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int ri = up ? 7 - i : i, ci = up ? 7 - cc : cc;
        XI[(c0 + ri) + (c0 + ci) * LDW] = x[i];
    }
}

// One level of the recursive doubling for L^-1 and U^-1: diagonal blocks of size SZ are inverted, the blocks that
// couple consecutive pairs follow as  X21 = -L22^-1 (L21 L11^-1),  X12 = -U11^-1 (U12 U22^-1).
template <int SZ>
__device__ __forceinline__ void inv_level(const double* S, double* LI, double* UI, double* TMP, int warp, int lane) {
    constexpr int SB = SZ / 8, PER = SB * SB, NBLK = (64 / (2 * SZ)) * 2 * PER;   // 8, 16, 32 output blocks
#pragma unroll
    for (int phase = 0; phase < 2; ++phase) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int b = warp + CHAIN_WARPS * i;
            if (i * CHAIN_WARPS < NBLK && b < NBLK) {
                const int pair = b / (2 * PER), rm = b % (2 * PER), isU = rm / PER, bb = rm % PER, bi = bb % SB, bj = bb / SB;
                const int o = 2 * SZ * pair;
                // L: rows [o+SZ, o+2SZ), cols [o, o+SZ);  U: rows [o, o+SZ), cols [o+SZ, o+2SZ)
                const int R0 = (isU ? o : o + SZ) + 8 * bi, C0 = (isU ? o + SZ : o) + 8 * bj;
                if (phase == 0) {      // TMP = L21 L11^-1   |   TMP = U12 U22^-1
                    blk8<SZ, false>(TMP + R0 + C0 * LDW, S + R0 + (isU ? o + SZ : o) * LDW,
                                    (isU ? UI + (o + SZ) : LI + o) + C0 * LDW, 1.0, lane);
                } else {               // X21 = -L22^-1 TMP  |   X12 = -U11^-1 TMP
                    blk8<SZ, false>((isU ? UI : LI) + R0 + C0 * LDW, isU ? UI + R0 + o * LDW : LI + R0 + (o + SZ) * LDW,
                                    TMP + (isU ? o : o + SZ) + C0 * LDW, -1.0, lane);
                }
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(512) k_chain_block(double* A, int lda, int j, int nb, int prev_nb, const double* Din, int ldin,
                                                     double* LUout, int ldout, double* Linv, double* Uinv, int* flags) {
    extern __shared__ double sm_ch[];
    double* S = sm_ch;                  // work matrix -> L (strictly lower, multipliers) and U
    double* LI = sm_ch + 64 * LDW;      // L^-1 (prologue: the L panel block)
    double* UI = sm_ch + 2 * 64 * LDW;  // U^-1 (prologue: the U panel block)
    double* TMP = sm_ch + 3 * 64 * LDW;
    double* rp = sm_ch + 4 * 64 * LDW;  // 64 pivot reciprocals
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    DBG_CLK(0);
    // ---- load: the diagonal block (identity-padded to 64), and for j > 0 the two panel blocks next to it ----
    const double* D = Din ? Din : A + j + (size_t)lda * j;
    const int ldd = Din ? ldin : lda;
    for (int idx = tid; idx < 64 * 64; idx += 512) {
        const int r = idx & 63, c = idx >> 6;
        S[r + c * LDW] = (r < nb && c < nb) ? D[r + (size_t)ldd * c] : (r == c ? 1.0 : 0.0);
        if (prev_nb > 0) {
            // L[j, j-1] (nb x prev_nb) sits left of the block, U[j-1, j] (prev_nb x nb) above it
            LI[r + c * LDW] = (r < nb && c < prev_nb) ? A[(j + r) + (size_t)lda * (j - prev_nb + c)] : 0.0;
            UI[r + c * LDW] = (r < prev_nb && c < nb) ? A[(j - prev_nb + r) + (size_t)lda * (j + c)] : 0.0;
        }
    }
    __syncthreads();
    if (prev_nb > 0) {
        mm_smem<64, true>(S, LI, UI, 8, 8, -1.0, warp, lane);
        __syncthreads();
    }
    for (int idx = tid; idx < 64 * 64; idx += 512) {
        const int r = idx & 63, c = idx >> 6;
        LI[r + c * LDW] = 0.0;
        UI[r + c * LDW] = 0.0;
    }
    __syncthreads();
    DBG_CLK(1);
    int bad = 0;
#pragma unroll 1
    for (int pb = 0; pb < 8; ++pb) lu_substep(pb, S, rp, tid, warp, lane, bad);
    diag_inverses(S, LI, UI, rp, tid);
    __syncthreads();
    if (bad && tid == 0) atomicOr(flags, FLAG_NOT_SPD);
    DBG_CLK(2);
    inv_level<8>(S, LI, UI, TMP, warp, lane);
    inv_level<16>(S, LI, UI, TMP, warp, lane);
    inv_level<32>(S, LI, UI, TMP, warp, lane);
    DBG_CLK(3);
    for (int idx = tid; idx < 64 * 64; idx += 512) {
        const int r = idx & 63, c = idx >> 6;
        Linv[r + 64 * c] = LI[r + c * LDW];
        Uinv[r + 64 * c] = UI[r + c * LDW];
        if (LUout != nullptr && r < nb && c < nb) LUout[r + (size_t)ldout * c] = S[r + c * LDW];
    }
    DBG_CLK(4);
}

#ifdef EQVIO_DEBUG_CLOCKS
extern "C" int eqvio_debug_clocks(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_dbg_clk, sizeof(long long) * 16); }
#endif

// Everything of the Schur problem except the leading k x k block A and (when !identity_border) the
// border entries written elsewhere.  Layout: A in [0,k)^2, identity padding on [k,kpad), border rows /
// columns at offset kpad:  [[A, 0, Cc], [0, I, 0], [R, 0, 0]].  identity_border: R = Cc = I (k x k).
__global__ void k_schur_setup(double* A, int lda, int k, int kpad, int r, int c, int identity_border) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x, col = blockIdx.y;
    if (row >= kpad + r || col >= kpad + c) return;
    if (row < k && col < k) return;
    double v;
    if (row < kpad && col < kpad) v = (row == col) ? 1.0 : 0.0;          // padding block
    else if (row >= kpad && col >= kpad) v = 0.0;                         // bottom-right
    else if (row < kpad) {                                                // top-right border
        if (row >= k) v = 0.0;
        else if (identity_border) v = (row == col - kpad) ? 1.0 : 0.0;
        else return;
    } else {                                                              // bottom-left border
        if (col >= k) v = 0.0;
        else if (identity_border) v = (row - kpad == col) ? 1.0 : 0.0;
        else return;
    }
    A[row + (size_t)lda * col] = v;
}

// Identity border columns [col0, col0 + ncols) over `rows` rows: entry (r, col0 + c) = (r == c && r < k).
__global__ void k_schur_identity_cols(double* A, int lda, int rows, int k, int col0, int ncols) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (row >= rows || c >= ncols) return;
    A[row + (size_t)lda * (col0 + c)] = (row == c && row < k) ? 1.0 : 0.0;
}

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
// %globaltimer stamp (ns): the only way to time points INSIDE a replayed CUDA graph (tools/graph_stamps.py)
__global__ void k_stamp(unsigned long long* slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *slot = t;
}
void launch_stamp(cudaStream_t s, unsigned long long* slot) { k_stamp<<<1, 1, 0, s>>>(slot); }
__global__ void k_nop() {}
void launch_nop(cudaStream_t s) { k_nop<<<1, 32, 0, s>>>(); }

__global__ void k_copy_block(const double* src, int lds, double* dst, int ldd, int rows, int cols) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r < rows && c < cols) dst[r + (size_t)ldd * c] = src[r + (size_t)lds * c];
}
__global__ void k_set_identity_rows(double* A, int lda, int row0, int n) {
    // A[row0 + i, c] = (i == c), i, c in [0, n)
    const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r < n && c < n) A[(row0 + r) + (size_t)lda * c] = (r == c) ? 1.0 : 0.0;
}
__global__ void k_add_diag_const(double* A, int lda, int n, double v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[i + (size_t)lda * i] += v;
}
__global__ void k_set_diag_one(double* A, int lda, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[i + (size_t)lda * i] = 1.0;
}

// removeOutliers test, VIOFilter.cpp:429-443: flag_i = |y_i - normalise(q_hat_i)| > threshold
__global__ void k_outlier_flags(Landmarks L, int N, const double* bearings, double thr, int* flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const V3 qh = inverse(load_Q(L, i)) * load_q0(L, i);
    const V3 yh = normalized(qh);
    const V3 d = v3(bearings[3 * i], bearings[3 * i + 1], bearings[3 * i + 2]) - yh;
    flags[i] = norm(d) > thr ? 1 : 0;
}

// removeRows / removeCols (VIOFilter.cpp:29-47) as one gather: dst[r,c] = src[map[r], map[c]]
__global__ void k_gather_sigma(const double* src, double* dst, int ld, const int* map, int n_new) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r < n_new && c < n_new) dst[r + (size_t)ld * c] = src[map[r] + (size_t)ld * map[c]];
}
// landmark arrays: dst field f, slot i = src field f, slot keep[i]
__global__ void k_gather_landmarks(const double* src, double* dst, int cap, const int* keep, int n_new) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (i < n_new) dst[(size_t)f * cap + i] = src[(size_t)f * cap + keep[i]];
}
__global__ void k_gather_bearings(const double* src, double* dst, const int* idx, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int s = idx[i];
        dst[3 * i] = src[3 * s]; dst[3 * i + 1] = src[3 * s + 1]; dst[3 * i + 2] = src[3 * s + 2];
    }
}

// addNewLandmarks, VIOFilter.cpp:345-391.  One block.  Median: element of rank oldN/2 of the squared
// depths (what nth_element at size/2 leaves there), then q0 = y * depth, Q = identity.
__global__ void __launch_bounds__(1024) k_add_landmarks(Landmarks L, int oldN, int newN, const double* bearings,
                                                        double initialSceneDepth, double* scratch) {
    __shared__ double s_depth;
    const int tid = threadIdx.x;
    if (tid == 0) s_depth = initialSceneDepth;
    for (int i = tid; i < oldN; i += blockDim.x) {
        const V3 qh = inverse(load_Q(L, i)) * load_q0(L, i);
        scratch[i] = dot(qh, qh);
    }
    __syncthreads();
    const int target = oldN / 2;
    for (int i = tid; i < oldN; i += blockDim.x) {
        const double v = scratch[i];
        int rank = 0;
        for (int j = 0; j < oldN; ++j) {
            const double u = scratch[j];
            rank += (u < v) || (u == v && j < i);
        }
        if (rank == target) s_depth = pow(v, 0.5);
    }
    __syncthreads();
    const double depth = s_depth;
    for (int k = tid; k < newN; k += blockDim.x) {
        const int i = oldN + k;
        L.q0(0)[i] = bearings[3 * i] * depth;
        L.q0(1)[i] = bearings[3 * i + 1] * depth;
        L.q0(2)[i] = bearings[3 * i + 2] * depth;
        L.Q(0)[i] = 1.0; L.Q(1)[i] = 0.0; L.Q(2)[i] = 0.0; L.Q(3)[i] = 0.0; L.Q(4)[i] = 1.0;
    }
}
// Sigma growth: rows/cols [n0, n1) zero, diagonal = initialPointVariance (VIOFilter.cpp:384-390)
__global__ void k_grow_sigma(double* S, int ld, int n0, int n1, double var) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r >= n1 || c >= n1) return;
    if (r < n0 && c < n0) return;
    S[r + (size_t)ld * c] = (r == c) ? var : 0.0;
}

// setInertialPoints, VIOFilter.cpp:93-118 (points in the world frame -> camera frame of xi0)
__global__ void k_set_inertial_points(const BaseState* st, Landmarks L, int N, const double* points) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const Se3 inv = inverse(st->pose0 * st->cam);
    const V3 q = inv * v3(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
    L.q0(0)[i] = q.x; L.q0(1)[i] = q.y; L.q0(2)[i] = q.z;
    L.Q(0)[i] = 1.0; L.Q(1)[i] = 0.0; L.Q(2)[i] = 0.0; L.Q(3)[i] = 0.0; L.Q(4)[i] = 1.0;
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

static const int LIFT_RSOLVE_SMEM = (4 * 64 * 65 + 2 * 4 * 64) * (int)sizeof(double);
static const int CHAIN_BLOCK_SMEM = (4 * 64 * LDW + 64) * (int)sizeof(double);
// The > 48 KB dynamic shared-memory opt-in is a per-device function attribute: once per device, with it current.
cudaError_t kernels_init_device() {
    cudaError_t e = cudaFuncSetAttribute(k_lift_rsolve, cudaFuncAttributeMaxDynamicSharedMemorySize, LIFT_RSOLVE_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_chain_block, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_BLOCK_SMEM);
}

void launch_step_prepare(cudaStream_t s, BaseState* st, StepScratch* sc, const ImuArgs& a, const RiccatiOut& ro) {
    k_step_prepare<<<1, 128, 0, s>>>(st, sc, a, ro);
}
void launch_feature_step(cudaStream_t s, BaseState* st, const StepScratch* sc, Landmarks L, int N, int do_riccati,
                         int discrete, const RiccatiOut& ro) {
    if (N > 0) k_feature_step<<<cdiv(N, 8), 256, 0, s>>>(st, sc, L, N, do_riccati, discrete, ro);
}
void launch_build_C_delta(cudaStream_t s, BaseState* st, Landmarks L, int N, const double* bearings, double* C, int ldc,
                          double* delta) {
    if (N > 0) k_build_C_delta<<<cdiv(N, 8), 256, 0, s>>>(st, L, N, bearings, C, ldc, delta);
}
void launch_gemv(cudaStream_t s, const double* K, int ldk, int n, int m, const double* x, double* y, double* part, int* cnt) {
    k_gemv<<<dim3(cdiv(n, 32), GEMV_SPLIT), 256, 0, s>>>(K, ldk, n, m, x, y, part, cnt);
}
void launch_lift_prepare(cudaStream_t s, BaseState* st, StepScratch* sc, const double* gamma) {
    k_lift_prepare<<<1, 32, 0, s>>>(st, sc, gamma);
}
void launch_lift_features(cudaStream_t s, const StepScratch* sc, Landmarks L, int N, const double* gamma, double* Aug,
                          int lda, int pb, double* yo) {
    k_lift_features<<<cdiv(N > 16 ? N : 16, 128), 128, 0, s>>>(sc, L, N, gamma, Aug, lda, pb, yo);
}
// (before the lift chain, off the critical path: every entry of R^T "pending")
void launch_lift_rsolve_reset(cudaStream_t s, double* Rt, int pb) {
    k_fill_u64<<<cdiv(4 * pb, 256), 256, 0, s>>>(reinterpret_cast<unsigned long long*>(Rt), 4 * pb, LIFT_RT_PENDING);
}
cudaError_t launch_lift_rsolve(cudaStream_t s, const double* Aug, int lda, int pb, const double* LinvBlocks, double* Rt, int* err) {
    const int nblk = (pb + 63) / 64;
    k_lift_rsolve<<<nblk, 256, LIFT_RSOLVE_SMEM, s>>>(Aug, lda, pb, LinvBlocks, Rt, err);
    return cudaGetLastError();
}
void launch_lift_solve(cudaStream_t s, BaseState* st, StepScratch* sc, const double* gamma, const double* Aug, int lda, int p, long rt_rs, long rt_cs, double rt_sign,
                       const double* Rt, const double* yo, int use_lift, int discrete, double* Gamma_out, int apply) {
    k_lift_solve<<<1, 128, 0, s>>>(st, sc, gamma, Aug, lda, p, Rt, rt_rs, rt_cs, rt_sign, yo, use_lift, discrete, Gamma_out, apply);
}
void launch_lift_apply(cudaStream_t s, BaseState* st, Landmarks L, int N, const double* gamma, int discrete) {
    if (N > 0) k_lift_apply<<<cdiv(N, 128), 128, 0, s>>>(st, L, N, gamma, discrete);
}
cudaError_t launch_chain_block(cudaStream_t s, double* A, int lda, int j, int nb, int prev_nb, const double* Din, int ldin,
                               double* LUout, int ldout, double* Linv, double* Uinv, int* flags) {
    k_chain_block<<<1, 512, CHAIN_BLOCK_SMEM, s>>>(A, lda, j, nb, prev_nb, Din, ldin, LUout, ldout, Linv, Uinv, flags);
    return cudaGetLastError();
}
void launch_schur_setup(cudaStream_t s, double* A, int lda, int k, int kpad, int r, int c, int identity_border) {
    k_schur_setup<<<dim3(cdiv(kpad + r, 256), kpad + c), 256, 0, s>>>(A, lda, k, kpad, r, c, identity_border);
}
void launch_schur_identity_cols(cudaStream_t s, double* A, int lda, int rows, int k, int col0, int ncols) {
    if (rows > 0 && ncols > 0) k_schur_identity_cols<<<dim3(cdiv(rows, 256), ncols), 256, 0, s>>>(A, lda, rows, k, col0, ncols);
}
void launch_copy_block(cudaStream_t s, const double* src, int lds, double* dst, int ldd, int rows, int cols) {
    if (rows > 0 && cols > 0) k_copy_block<<<dim3(cdiv(rows, 256), cols), 256, 0, s>>>(src, lds, dst, ldd, rows, cols);
}
void launch_set_identity_rows(cudaStream_t s, double* A, int lda, int row0, int n) {
    if (n > 0) k_set_identity_rows<<<dim3(cdiv(n, 256), n), 256, 0, s>>>(A, lda, row0, n);
}
void launch_add_diag_const(cudaStream_t s, double* A, int lda, int n, double v) {
    if (n > 0) k_add_diag_const<<<cdiv(n, 256), 256, 0, s>>>(A, lda, n, v);
}
void launch_set_diag_one(cudaStream_t s, double* A, int lda, int n) {
    if (n > 0) k_set_diag_one<<<cdiv(n, 256), 256, 0, s>>>(A, lda, n);
}
void launch_outlier_flags(cudaStream_t s, Landmarks L, int N, const double* bearings, double thr, int* flags) {
    if (N > 0) k_outlier_flags<<<cdiv(N, 128), 128, 0, s>>>(L, N, bearings, thr, flags);
}
void launch_gather_sigma(cudaStream_t s, const double* src, double* dst, int ld, const int* map, int n_new) {
    if (n_new > 0) k_gather_sigma<<<dim3(cdiv(n_new, 256), n_new), 256, 0, s>>>(src, dst, ld, map, n_new);
}
void launch_gather_landmarks(cudaStream_t s, const double* src, double* dst, int cap, const int* keep, int n_new) {
    if (n_new > 0) k_gather_landmarks<<<dim3(cdiv(n_new, 128), LM_FIELDS), 128, 0, s>>>(src, dst, cap, keep, n_new);
}
void launch_gather_bearings(cudaStream_t s, const double* src, double* dst, const int* idx, int n) {
    if (n > 0) k_gather_bearings<<<cdiv(n, 128), 128, 0, s>>>(src, dst, idx, n);
}
void launch_add_landmarks(cudaStream_t s, Landmarks L, int oldN, int newN, const double* bearings, double depth0,
                          double* scratch) {
    k_add_landmarks<<<1, 1024, 0, s>>>(L, oldN, newN, bearings, depth0, scratch);
}
void launch_grow_sigma(cudaStream_t s, double* S, int ld, int n0, int n1, double var) {
    k_grow_sigma<<<dim3(cdiv(n1, 256), n1), 256, 0, s>>>(S, ld, n0, n1, var);
}
void launch_set_inertial_points(cudaStream_t s, const BaseState* st, Landmarks L, int N, const double* points) {
    if (N > 0) k_set_inertial_points<<<cdiv(N, 128), 128, 0, s>>>(st, L, N, points);
}

}  // namespace eqvio
