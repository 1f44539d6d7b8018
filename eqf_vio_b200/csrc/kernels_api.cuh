// kernels_api.cuh — host-callable launchers of the kernels in filter_kernels.cu.
#pragma once
#include "filter_kernels.cuh"

namespace eqvio {
// once per device (with it current) before the first launch there: per-device function attributes
cudaError_t kernels_init_device();
void launch_step_prepare(cudaStream_t s, BaseState* st, StepScratch* sc, const ImuArgs& a, const RiccatiOut& ro);
void launch_feature_step(cudaStream_t s, BaseState* st, const StepScratch* sc, Landmarks L, int N, int do_riccati,
                         int discrete, const RiccatiOut& ro);
void launch_build_C_delta(cudaStream_t s, BaseState* st, Landmarks L, int N, const double* bearings, double* C, int ldc,
                          double* delta);
// y = K x (K n x m column-major).  part: 8 n doubles of scratch; cnt: ceil(n/32) ints, zero before the first use (left zero).
void launch_gemv(cudaStream_t s, const double* K, int ldk, int n, int m, const double* x, double* y, double* part, int* cnt);
void launch_lift_prepare(cudaStream_t s, BaseState* st, StepScratch* sc, const double* gamma);
void launch_lift_features(cudaStream_t s, const StepScratch* sc, Landmarks L, int N, const double* gamma, double* Aug,
                          int lda, int pb, double* yo);
void launch_lift_rsolve_reset(cudaStream_t s, double* Rt, int pb);
cudaError_t launch_lift_rsolve(cudaStream_t s, const double* Aug, int lda, int pb, const double* LinvBlocks, double* Rt, int* err);
// R^T = Ym^T Sigma_sub^-1, entry (a, col) = rt_sign * Rt[a * rt_rs + col * rt_cs]
void launch_lift_solve(cudaStream_t s, BaseState* st, StepScratch* sc, const double* gamma, const double* Aug, int lda, int p, long rt_rs, long rt_cs, double rt_sign,
                       const double* Rt, const double* yo, int use_lift, int discrete, double* Gamma_out, int apply);
void launch_schur_identity_cols(cudaStream_t s, double* A, int lda, int rows, int k, int col0, int ncols);
void launch_lift_apply(cudaStream_t s, BaseState* st, Landmarks L, int N, const double* gamma, int discrete);
// Diagonal block j (nb x nb, nb <= 64) of the blocked Schur elimination of A: D = A[j,j] (or Din) minus, when
// prev_nb > 0, the product of the panel blocks L[j, j-prev_nb] U[j-prev_nb, j] that the trailing update skipped;
// writes L^-1, U^-1 (64 x 64, identity-padded) and optionally the factors.
cudaError_t launch_chain_block(cudaStream_t s, double* A, int lda, int j, int nb, int prev_nb, const double* Din, int ldin,
                               double* LUout, int ldout, double* Linv, double* Uinv, int* flags);
void launch_stamp(cudaStream_t s, unsigned long long* slot);
void launch_nop(cudaStream_t s);   // empty kernel: a few microseconds of stream-ordered delay
void launch_schur_setup(cudaStream_t s, double* A, int lda, int k, int kpad, int r, int c, int identity_border);
void launch_copy_block(cudaStream_t s, const double* src, int lds, double* dst, int ldd, int rows, int cols);
void launch_set_identity_rows(cudaStream_t s, double* A, int lda, int row0, int n);
void launch_add_diag_const(cudaStream_t s, double* A, int lda, int n, double v);
void launch_set_diag_one(cudaStream_t s, double* A, int lda, int n);
void launch_outlier_flags(cudaStream_t s, Landmarks L, int N, const double* bearings, double thr, int* flags);
void launch_gather_sigma(cudaStream_t s, const double* src, double* dst, int ld, const int* map, int n_new);
void launch_gather_landmarks(cudaStream_t s, const double* src, double* dst, int cap, const int* keep, int n_new);
void launch_gather_bearings(cudaStream_t s, const double* src, double* dst, const int* idx, int n);
void launch_add_landmarks(cudaStream_t s, Landmarks L, int oldN, int newN, const double* bearings, double depth0,
                          double* scratch);
void launch_grow_sigma(cudaStream_t s, double* S, int ld, int n0, int n1, double var);
void launch_set_inertial_points(cudaStream_t s, const BaseState* st, Landmarks L, int N, const double* points);
}  // namespace eqvio
