// dgemm_sm100.cuh — fp64 GEMM for sm_100a: TMA-staged (cp.async.bulk.tensor, SWIZZLE_128B) operand
// tiles in shared memory, an mbarrier full/empty ring fed by one producer warp, and DMMA.8x8x4
// (mma.sync.aligned.m8n8k4.f64 — the only native fp64 tensor shape on sm_100a; larger PTX shapes
// decompose into it) issued by the consumer warps with accumulators in registers.  tcgen05 has no
// f64 kind, so there is no TMEM path for this arithmetic.
//
// Computes, column-major:   D = alpha * A * op(B) + beta * Cin   (+ optional Riccati epilogue)
//   A   : M x K, element (m,k) at A[m + k*lda]                      ("M-major")
//   B   : transB == 0 : K x N, element (k,n) at B[k + n*ldb]        ("K-major")
//         transB == 1 : N x K, element (n,k) at B[n + k*ldb]        ("N-major"), op(B) = B^T
//   Cin, D : M x N column-major; D may alias Cin (each element is read then written by one thread).
//
// These are the reference's dense Sigma contractions (eqf_vio/src/VIOFilter.cpp:188-189, 276-277,
// 297; eqf_vio/src/EqFMatrices.cpp:239) which Eigen evaluates as general dynamic GEMMs.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace eqvio {

enum GemmEpilogue : int {
    EPI_AXPBY = 0,   // D = alpha*acc + beta*Cin
    EPI_RICCATI = 1  // D = acc + T*(Pdiag[m]*(m==n) + sum_j Rd[j]*Bb[m,j]*Bb[n,j])   (VIOFilter.cpp:188-189)
};

struct GemmEpilogueArgs {
    double alpha, beta;
    const double* Cin;
    int ldcin;
    // Riccati epilogue
    double T;
    const double* T_dev;  // when non-null the step length is read from device memory (lets a CUDA graph replay the launch)
    const double* Bb;  // n x 6 column-major, leading dimension ldbb
    int ldbb;
    double Rd[6];      // diag(velOmegaVariance x3, velAccelVariance x3)
    double Pd[5];      // process variances: biasOmega, biasAccel, gravity, velocity, point
};

struct GemmProblem {
    int M, N, K;
    const double* A; int lda;
    const double* B; int ldb; int transB;
    double* D; int ldd;
    int epilogue;
    GemmEpilogueArgs epi;
    // Output elements with m < skip_m and n < skip_n are left untouched (tiles entirely inside the box exit at
    // once): the trailing update of a Schur step leaves the next diagonal block to the chain kernel.
    int skip_m = 0, skip_n = 0;
};

// Once per device before the first launch on it (the > 48 KB dynamic shared-memory opt-in is a per-device function
// attribute): call with that device current.  eqvio_create and the handle-less entry points do.
cudaError_t dgemm_init_device();

// Host API.  All pointers are device pointers; A, B must be 16-byte aligned with even lda/ldb and
// be readable up to the next multiple of 16 rows (the library's buffers are padded accordingly).
// Returns cudaSuccess or the launch / encode error.  `flops_out` (optional) receives 2*M*N*K.
cudaError_t dgemm_launch(const GemmProblem& p, cudaStream_t stream, int force_config = -1);
// Two dependent products W = A1 B1 (NN) and D = W op(B2) (+ epilogue) as ONE launch — the reference's (A B) C chains
// (F Sigma) F^T, (C Sigma) C^T, (K C) Sigma.  The second product's tiles start as soon as their row block of W is
// complete (per-row-block counters in `sync`: DGEMM_PAIR_SYNC_INTS ints, zero before the first use and left zero) and
// fill the SM slots the first product's tail leaves idle.  Requires second.A == first.D, equal M, first not transposed.
// second.D must not alias first.A / first.B (rejected): second-product tiles are stored while first-product tiles of
// later row blocks still read their operands.
// Launches on one stream must not overlap launches on another with the same `sync` buffer.
static const int DGEMM_PAIR_MAX_ROW_BLOCKS = 8192;
// split-K form (dgemm_pair_splitk_launch): one arrival counter per tile of both products behind the row-block counters, and a
// workspace of DGEMM_SPLITK_MAX_SLOTS partial tiles (32 x 32 doubles each)
static const int DGEMM_SPLITK_MAX_SLOTS = 5120;
static const int DGEMM_PAIR_SYNC_INTS = 8 + DGEMM_PAIR_MAX_ROW_BLOCKS + DGEMM_SPLITK_MAX_SLOTS;
inline size_t dgemm_splitk_ws_doubles() { return (size_t)DGEMM_SPLITK_MAX_SLOTS * 32 * 32; }
// Split-K factor the pair launch should use for these shapes (1 = none), and the launch itself: each tile of both products is
// computed by `ks` CTAs over k-ranges, partials summed in a fixed order by the last arriver.
int dgemm_pair_splitk(const GemmProblem& first, const GemmProblem& second);
cudaError_t dgemm_pair_splitk_launch(const GemmProblem& first, const GemmProblem& second, int ks, int* sync, double* ws, cudaStream_t stream);
cudaError_t dgemm_pair_launch(const GemmProblem& first, const GemmProblem& second, int* sync, cudaStream_t stream);
// Stream-K form of the same two-product step (persistent CTAs, equal DMMA work per SM, one grid barrier between the
// products) for shapes whose products are a single partial wave of tiles.  `sync`: DGEMM_STREAMK_SYNC_INTS ints, zero
// before the first use and left zero; `ws`: dgemm_streamk_ws_doubles() doubles of scratch.  Same operand rules as
// dgemm_pair_launch.  Requires all its CTAs co-resident (6 per SM): not for use concurrently with other big launches.
static const int DGEMM_STREAMK_SYNC_INTS = 8 + 148 * 6;
size_t dgemm_streamk_ws_doubles();
bool dgemm_streamk_pays(const GemmProblem& first, const GemmProblem& second);
cudaError_t dgemm_streamk_pair_launch(const GemmProblem& first, const GemmProblem& second, int* sync, double* ws, cudaStream_t stream);
// One product, every 32 x 32 tile cut into ks k-ranges (partials in `ws`: tiles * ks slots of 32 x 32 doubles; `cnt`: one int per tile,
// zero before the first use and left zero): thin products with a long k loop.
cudaError_t dgemm_splitk_launch(const GemmProblem& p, int ks, double* ws, int* cnt, cudaStream_t stream);
// Whether the single launch is the faster choice for these shapes (else: two dgemm_launch calls).
bool dgemm_pair_pays(const GemmProblem& first, const GemmProblem& second);
// Which tile configuration dgemm_launch would pick (for DESIGN.md / tests).
int dgemm_pick_config(int M, int N, int K);
const char* dgemm_config_name(int cfg);
int dgemm_num_configs();

}  // namespace eqvio
