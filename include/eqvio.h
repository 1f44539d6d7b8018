/*
 * eqvio.h — C ABI of the B200-native EqF-VIO filter hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI layer; its boundary is the
 * public surface of `class VIOFilter` (reference: eqf_vio/include/eqf_vio/VIOFilter.h:64-88).  Every
 * entry point below names the reference member it replaces.  All types are POD, every function
 * returns an `int` status (0 = OK), no exception crosses this boundary, output buffers are
 * caller-owned.  One handle = one CUDA device + one CUDA stream; a handle is not thread-safe but
 * distinct handles are independent (one filter session per GPU for the multi-session configuration).
 *
 * Conventions (identical to the reference):
 *   - Sigma is n x n, n = 11 + 3N, column-major (Eigen default), index map
 *     [0,3) gyro bias, [3,6) accel bias, [6,8) gravity chart, [8,11) body velocity,
 *     [11+3i, 14+3i) landmark i          (eqf_vio/src/VIOFilter.cpp:54-57,163-167)
 *   - rotations are unit quaternions in (w, x, y, z) order; poses are (x, y, z, qw, qx, qy, qz),
 *     the "xw" order of the YAML `cameraOffset` key (eqf_vio/include/eqf_vio/VIOFilterSettings.h:95-108)
 *   - bearings are unit 3-vectors in the camera frame, sorted by ascending id
 *     (eqf_vio/src/VIOFilter.cpp:239-240)
 */
#ifndef EQVIO_H
#define EQVIO_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EQVIO_SIGMA_BASE_SIZE 11 /* eqf_vio/include/eqf_vio/VIOFilter.h:28 */
#define EQVIO_GRAVITY_CONSTANT 9.81 /* eqf_vio/include/eqf_vio/IMUVelocity.h:22 */

/* Field-for-field mirror of VIOFilter::Settings (VIOFilterSettings.h:28-50), same defaults via
 * eqvio_settings_default().  bools are ints. */
typedef struct eqvio_settings {
    double biasOmegaProcessVariance;
    double biasAccelProcessVariance;
    double gravityProcessVariance;
    double velocityProcessVariance;
    double pointProcessVariance;
    double velOmegaVariance;
    double velAccelVariance;
    double measurementVariance;
    double initialGravityVariance;
    double initialVelocityVariance;
    double initialPointVariance;
    double initialBiasOmegaVariance;
    double initialBiasAccelVariance;
    double initialSceneDepth;
    double outlierThreshold;
    int useInnovationLift;
    int useDiscreteInnovationLift;
    int useDiscreteVelocityLift;
    int fastRiccati;
    double initialAccelBias[3];
    double initialOmegaBias[3];
    double cameraOffset[7]; /* x y z qw qx qy qz */
} eqvio_settings_t;

typedef struct eqvio_filter* eqvio_handle_t;

/* Status codes.  The reference returns void and signals these conditions by silently returning
 * (VIOFilter.cpp:147-152, 235-236, 258-259), by assert (:190,205,299-300) or by throwing
 * std::domain_error (libs/core/src/SO3.cpp:160-161). */
enum {
    EQVIO_OK = 0,
    EQVIO_SKIPPED_DT = 1,       /* dt <= 0 or no previous stamp: nothing integrated (VIOFilter.cpp:147-152) */
    EQVIO_NOT_INITIALISED = 2,  /* vision before the first IMU sample (VIOFilter.cpp:235) */
    EQVIO_EMPTY_MEASUREMENT = 3,/* no bearings left after bookkeeping (VIOFilter.cpp:258-259) */
    EQVIO_ERR_ARG = -1,
    EQVIO_ERR_CUDA = -2,
    EQVIO_ERR_NAN = -3,            /* NaN in Sigma or X (the reference's asserts) */
    EQVIO_ERR_SINGULAR_CHART = -4, /* SO3FromVectors on opposing vectors (SO3.cpp:160) */
    EQVIO_ERR_NOT_SPD = -5,        /* the unpivoted LU of S or Sigma_sub met a zero / non-finite pivot */
    EQVIO_ERR_NO_DEVICE = -6,
    EQVIO_ERR_UNSORTED = -7        /* bearings not sorted by ascending id (VIOFilter.cpp:239-240) */
};

/* ---- lifetime ------------------------------------------------------------------------------ */

/* Struct defaults of VIOFilter::Settings (VIOFilterSettings.h:29-50). */
int eqvio_settings_default(eqvio_settings_t* s);

/* VIOFilter::VIOFilter(const Settings&) (VIOFilter.cpp:60-73).  `device` is the CUDA ordinal. */
int eqvio_create(const eqvio_settings_t* settings, int device, eqvio_handle_t* out);
int eqvio_destroy(eqvio_handle_t h);
/* The reference's `settings` member is public and read at use time (VIOFilter.cpp:126,163-175,269,293): this
 * replaces the settings the handle works with from the next call on.  Like assigning through the reference's
 * pointer it does not re-apply constructor-time effects (initial variances, initial biases, camera offset). */
int eqvio_set_settings(eqvio_handle_t h, const eqvio_settings_t* settings);
/* VIOFilter::reset() (VIOFilter.cpp:84-91), exactly its member list: xi0 = VIOState(), X = Identity (no landmarks),
 * Sigma = I(11), currentTime = -1, currentVelocity = 0.  Like the reference it does NOT touch inputBias,
 * initialisedFlag, the accumulated velocity / time or the settings.  One stated deviation: `VIOState()` leaves its
 * Eigen members uninitialised in the reference; here xi0 becomes identity pose, zero velocity, identity camera
 * offset. */
int eqvio_reset(eqvio_handle_t h);
/* VIOFilter::setAuxiliaryData (VIOFilter.cpp:75-82): xi0.pose = (attitude (w,x,y,z), position), xi0.velocity = 0,
 * xi0.cameraOffset = cam_offset (x y z qw qx qy qz), initialisedFlag = true. */
int eqvio_set_auxiliary_data(eqvio_handle_t h, const double attitude_wxyz[4], const double position[3],
                             const double cam_offset[7]);
/* VIOFilter::initialiseFromIMUData (VIOFilter.cpp:133-144) as a public call: attitude from the accelerometer direction
 * of the sample AS GIVEN (no un-biasing), xi0.velocity = 0, initialisedFlag = true.  Opposing vectors (the reference
 * throws, SO3.cpp:160) return EQVIO_ERR_SINGULAR_CHART. */
int eqvio_initialise_from_imu(eqvio_handle_t h, const double omega[3], const double accel[3]);

/* ---- inputs --------------------------------------------------------------------------------
 * Error model.  process_imu / process_vision are ASYNCHRONOUS: they validate their arguments, enqueue the step
 * and return.  Argument errors (EQVIO_ERR_ARG, EQVIO_ERR_UNSORTED — checked before anything is integrated, so a
 * rejected frame leaves the filter untouched) and CUDA launch errors are returned by the call itself.  Conditions
 * only the device can see — a singular chart (SO3FromVectors on opposing vectors), a zero / non-finite pivot,
 * a NaN — set sticky bits in device memory; they surface as the return value of the next synchronising call
 * (eqvio_get_state, eqvio_synchronize, the kernel-level entry points) or through eqvio_get_flags, which can also
 * clear them.  set_snapshot and reset clear them too. */

/* VIOFilter::processIMUData (VIOFilter.cpp:120-131).  Asynchronous: enqueues on the handle's stream
 * and returns; the 7 doubles travel as kernel arguments.  Returns EQVIO_SKIPPED_DT when nothing was
 * integrated (first sample, dt <= 0) — the velocity/time latch still happens, as in the reference. */
int eqvio_process_imu(eqvio_handle_t h, double stamp, const double omega[3], const double accel[3]);

/* VIOFilter::processVisionData (VIOFilter.cpp:232-302).  `ids` (n ints) ascending, `bearings`
 * 3n doubles (x,y,z per bearing, unit norm), both HOST pointers. */
int eqvio_process_vision(eqvio_handle_t h, double stamp, int n, const int* ids, const double* bearings);

/* Same call with the bearings already resident in device memory (ids stay on the host: landmark
 * bookkeeping is host-side index work).  Used for the HBM-resident throughput figure. */
int eqvio_process_vision_dev(eqvio_handle_t h, double stamp, int n, const int* ids,
                             const double* bearings_dev);

/* VIOFilter::setInertialPoints (VIOFilter.cpp:93-118): n points (id, world xyz). */
int eqvio_set_inertial_points(eqvio_handle_t h, int n, const int* ids, const double* points);

/* ---- outputs (synchronise the handle's stream) ---------------------------------------------- */

/* VIOFilter::getTime (VIOFilter.cpp:343) */
int eqvio_get_time(eqvio_handle_t h, double* t);
/* Current number of landmarks N (n = 11 + 3N). */
int eqvio_get_num_landmarks(eqvio_handle_t h, int* n);
/* VIOFilter::stateEstimate (VIOFilter.cpp:304) = stateGroupAction(X, xi0) (VIOGroup.cpp:23).
 * pose/cam_offset: x y z qw qx qy qz; ids/landmarks may be NULL; *n receives N.  cap = capacity of
 * ids/landmarks in landmarks. */
int eqvio_get_state(eqvio_handle_t h, double pose[7], double velocity[3], double cam_offset[7],
                    int* n, int cap, int* ids, double* landmarks);
/* Pose only (x y z qw qx qy qz) + stamp: the 8-double record gathered across sessions. */
int eqvio_get_pose_record(eqvio_handle_t h, double rec[8]);
/* Device pointer to the 8-double pose record refreshed by every vision update (t x y z qw qx qy qz);
 * valid for the lifetime of the handle; stream-ordered after the update. */
int eqvio_pose_record_dev(eqvio_handle_t h, double** dev_ptr);
/* Multi-session pose gather without the main stream: the first call switches publishing on and returns a side
 * stream (cudaStream_t as void*) and a device buffer of 8 doubles; from then on every vision update is followed, ON
 * THAT STREAM, by a copy of the pose record into the buffer.  A caller's collective (NCCL all-gather of the buffer)
 * enqueued on the returned stream therefore waits for the update that produced the record but never sits in front
 * of the next IMU tick; the next frame's copy is stream-ordered behind the collective. */
int eqvio_pose_publish(eqvio_handle_t h, void** gather_stream, double** published_dev);
/* Sticky device-side error bits as a status code (EQVIO_OK / ERR_SINGULAR_CHART / ERR_NOT_SPD / ERR_NAN); synchronises
 * the handle's stream; `clear` != 0 resets them. */
int eqvio_get_flags(eqvio_handle_t h, int* status, int clear);
/* VIOFilter::stateCovariance (VIOFilter.cpp:306-309): n x n column-major into dst with leading
 * dimension ld >= n. */
int eqvio_get_covariance(eqvio_handle_t h, double* dst, int ld);
/* inputBias (VIOFilter.h:46): omega bias [0,3), accel bias [3,6). */
int eqvio_get_bias(eqvio_handle_t h, double bias[6]);

/* ---- snapshot / restore (lossless; the reference only has the write-only CSV dump
 *      operator<< at VIOFilter.cpp:311-341) --------------------------------------------------
 * Layout (doubles):  [0] N  [1] currentTime  [2] initialisedFlag  [3] accumulatedTime
 *   [4,10) inputBias  [10,16) currentVelocity (omega, accel)  [16,22) accumulatedVelocity
 *   [22,29) xi0.pose (qw qx qy qz x y z)  [29,32) xi0.velocity  [32,39) xi0.cameraOffset (q, x)
 *   [39,46) X.A (q, x)  [46,49) X.w
 *   then N records of 9: id, q0 xyz, Q quaternion wxyz, Q scale a
 *   then Sigma, n x n column-major.
 */
#define EQVIO_SNAPSHOT_HEADER 49
#define EQVIO_SNAPSHOT_PER_LANDMARK 9
size_t eqvio_snapshot_size(int n_landmarks); /* in doubles */
int eqvio_get_snapshot(eqvio_handle_t h, double* dst, size_t cap);
int eqvio_set_snapshot(eqvio_handle_t h, const double* src, size_t len);

/* ---- kernel-level entry points (unit parity + ncu).  They act on the handle's current
 *      (xi0, X, Sigma); outputs are HOST buffers. ------------------------------------------- */

/* EqFStateMatrixA_euclid_impl + EqFInputMatrixB_euclid_impl + biased assembly
 * (EqFMatrices.cpp:277-317,346-382; VIOFilter.cpp:177-185).  Writes F = I + T*[[0,0],[-Bt,A0t]]
 * (n x n, col-major, ld = n) and B_b = [0;Bt] (n x 6) for mean angular velocity `omega`. */
int eqvio_build_FB(eqvio_handle_t h, double T, const double omega[3], double* F, double* Bb);
/* Riccati step Sigma <- T (P + B_b R B_b^T) + F Sigma F^T (VIOFilter.cpp:162-189) on the handle's
 * Sigma, with A/B evaluated at the current state.  Does not propagate X. */
int eqvio_riccati_propagate(eqvio_handle_t h, double T, const double omega[3]);
/* delta = outputCoordinateChart(outputGroupAction(X^-1, y), measureSystemState(xi0)) and
 * C = [0, C0] (VIOFilter.cpp:264-273; EqFMatrices.cpp:319-344).  bearings: 3N host doubles aligned
 * with the state.  C is m x n col-major (ld = m), delta has m = 2N entries; either may be NULL. */
int eqvio_build_C_delta(eqvio_handle_t h, const double* bearings, double* C, double* delta);
/* S, K, gamma = K delta, Sigma <- Sigma - K C Sigma (VIOFilter.cpp:276-279,297).  Outputs optional:
 * K (n x m col-major), gamma (n).  Does not lift / touch X. */
int eqvio_gain_update(eqvio_handle_t h, const double* bearings, double* K, double* gamma);
/* bundleLift (EqFMatrices.cpp:173-252) of a base innovation gamma_eqf (5+3N) with the handle's
 * current Sigma[6:,6:]; Gamma out has 9+3N entries. */
int eqvio_bundle_lift(eqvio_handle_t h, const double* gamma_eqf, double* Gamma);

/* Plain fp64 GEMM on the library's DMMA kernel, host buffers, column-major:
 * C <- alpha * A * op(B) + beta * C,  A is M x K, op(B) is K x N (transB: B stored N x K).
 * `reps` > 1 repeats the launch and *ms receives the mean CUDA-event time per launch (may be NULL). */
int eqvio_dgemm(int device, int transB, int M, int N, int K, double alpha, const double* A, int lda,
                const double* B, int ldb, double beta, double* C, int ldc, int reps, float* ms);

/* The reference's left-to-right product chains (A B) C — (F Sigma) F^T VIOFilter.cpp:189, (C Sigma) C^T :276,
 * (K C) Sigma :297 — as ONE launch of the library's pair kernel: W = A1 * B1 (M x N1, K1 deep), then
 * D = alpha2 * W * op(B2) (M x N2, N1 deep; transB2: B2 stored N2 x N1), the second product's tiles starting as
 * their row block of W completes.  Host buffers, column-major; W and D are outputs (either may be NULL).
 * `reps` > 1 repeats the launch and *ms receives the mean CUDA-event time per launch (may be NULL). */
int eqvio_dgemm_pair(int device, int M, int N1, int K1, const double* A1, int lda1, const double* B1, int ldb1,
                     int transB2, int N2, double alpha2, const double* B2, int ldb2, double* W, int ldw, double* D,
                     int ldd, int reps, float* ms);

/* OPT-IN, kernel level: the same product C = A * op(B) (the reference's dense Sigma contractions, VIOFilter.cpp:188-189,
 * 276-277, 297) assembled EXACTLY from 8-bit integer products on the 5th-generation tensor cores (tcgen05.mma kind::i8,
 * accumulators in TMEM) by Ozaki splitting: `slices` 7-bit signed digits per operand entry (8: as accurate as an fp64 GEMM
 * relative to row-max x column-max; 9: entry by entry on operands with Sigma's dynamic range), all digit pairs of one
 * weight summed in one int32 accumulator, only the per-weight sums converted to fp64.  The DMMA pipe the native path runs
 * on tops out at 37 TFLOP/s; this path is not bound by it.  The 128-aligned block at the END of C runs on tcgen05, the
 * strips in front of it on the DMMA kernel.  M, N >= 128.  Host buffers, column-major; `reps` > 1 times the whole call
 * (split + products + strips, *ms_total) and the tcgen05 kernel alone (*ms_gemm), mean per repetition.  Not used by the
 * filter path (eqvio_process_*) in this version. */
int eqvio_dgemm_ozaki(int device, int transB, int M, int N, int K, const double* A, int lda, const double* B, int ldb,
                      double* C, int ldc, int slices, int reps, float* ms_total, float* ms_gemm);

/* `S.inverse()` (VIOFilter.cpp:277) on its own: the blocked Schur elimination of [[S, I], [I, 0]] by unpivoted LU that the
 * update runs (chain kernels, in-place panel solves, look-ahead trailing updates), for an m x m matrix S (col-major, host;
 * symmetric positive definite up to round-off — no pivoting).  Uses the handle's work buffers (its capacity grows to m/2
 * landmarks if needed); the filter state is not touched. */
int eqvio_schur_inverse(eqvio_handle_t h, int m, const double* S, int lds, double* Sinv, int ldsi);

/* One diagonal block of the blocked Schur eliminations that stand in for `S.inverse()` (VIOFilter.cpp:277)
 * and `Sigma.inverse()` (EqFMatrices.cpp:239): unpivoted LU of the nb x nb block A (nb <= 64, col-major,
 * host) and the triangular inverses L^-1, U^-1 as 64 x 64 identity-padded col-major matrices.  LU (nb x nb,
 * ld = nb: unit-lower multipliers below the diagonal, U on and above), Linv, Uinv may be NULL.
 * `reps` > 1 repeats the launch and *us receives the mean CUDA-event time per launch in microseconds. */
int eqvio_getrf_block(int device, int nb, const double* A, int lda, double* LU, double* Linv, double* Uinv,
                      int reps, float* us);

/* ---- instrumentation ----------------------------------------------------------------------- */

int eqvio_synchronize(eqvio_handle_t h);
/* Number of kernel launches issued by this handle since creation / last reset of the counter. */
int eqvio_launch_count(eqvio_handle_t h, long long* count, int reset);
/* CUDA graphs: the launch sequence of a vision update (and of the Riccati step behind k_step_prepare) depends
 * only on the landmark count, on which of the two Sigma buffers is current and on mode flags; the second time
 * such a key is seen the sequence is stream-captured and from then on replayed with one cudaGraphLaunch.
 * eqvio_set_graphs(h, 0) issues every launch directly (also: environment EQVIO_GRAPHS=0 at eqvio_create);
 * eqvio_graph_stats reports graph replays so far and instantiated graphs held. */
int eqvio_set_graphs(eqvio_handle_t h, int on);
int eqvio_graph_stats(eqvio_handle_t h, long long* graph_launches, int* cached_graphs);
/* When enabled, every Sigma-contraction (GEMM) launch is bracketed by CUDA events on the handle's
 * stream; eqvio_profile_read synchronises and returns launches, total ms and executed flops. */
int eqvio_profile_enable(eqvio_handle_t h, int on);
int eqvio_profile_read(eqvio_handle_t h, long long* gemm_launches, double* gemm_ms, double* gemm_flops,
                       int reset);
/* Same, per class of launch: 0 = Riccati GEMMs (F Sigma, W F^T), 1 = update GEMMs (C Sigma, S, Sigma C^T,
 * K, K C, Sigma - K C Sigma), 2 = GEMMs inside the blocked Schur eliminations (S^-1, Sigma_sub^-1),
 * 3 = the diagonal-block LU kernels of those eliminations (flops reported as 0), 4 = the O(N) / O(n m) kernels,
 * 5 = the Riccati products issued on the int8 tensor cores (fp64-equivalent flops 2 M N K; see eqvio_riccati_arith). */
int eqvio_profile_read_class(eqvio_handle_t h, int cls, long long* launches, double* ms, double* flops, int reset);
/* Timeline of every bracketed launch since profiling was enabled: 5 doubles per entry — class (0-3 as
 * above, 4 = the O(N) / O(n m) kernels), stream lane (0 main, 1 side, 2 lift chain, 3 / 4 the chains' helper
 * streams), start ms, end ms (relative to eqvio_profile_enable), flops.  *count receives the number of
 * entries available; at most cap_entries are copied to out (may be NULL). */
int eqvio_profile_timeline(eqvio_handle_t h, double* out, size_t cap_entries, size_t* count);
/* Which arithmetic the Riccati step's two Sigma contractions (VIOFilter.cpp:188-189) run on at the current landmark count:
 * *int8_slices = 0 — fp64 DMMA tiles; S > 0 — the 128-aligned landmark block as an exact sum of S (S + 1) / 2 int8 x int8 -> int32
 * tensor-core products per fp64 product (tcgen05.mma kind::i8, Ozaki splitting with S slices of 7 bits; S = 8 carries 55 mantissa
 * bits), the rows / columns in front of that block on fp64 DMMA.  Selected per size (environment EQVIO_OZAKI, 0 = never). */
int eqvio_riccati_arith(eqvio_handle_t h, int* int8_slices);
/* Diagnostics of the int8 Riccati kernel (environment EQVIO_OZ_STAMPS=1 at eqvio_create, else *count = 0): per-CTA clock stamps of
 * the last two steps (four launches, by tick parity), 4 x 1024 x 16 words (tools/oz_stamps.py names them). */
int eqvio_oz_stamps(eqvio_handle_t h, long long* out, size_t cap_words, size_t* count);
/* The handle's CUDA stream (cudaStream_t as void*), for callers that order their own work after it. */
int eqvio_stream(eqvio_handle_t h, void** stream);
const char* eqvio_status_string(int status);
const char* eqvio_version(void);

#ifdef __cplusplus
}
#endif
#endif /* EQVIO_H */
