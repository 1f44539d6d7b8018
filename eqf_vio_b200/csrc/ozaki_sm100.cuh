// ozaki_sm100.cuh — fp64 GEMM on the INT8 tensor cores of sm_100a (tcgen05.mma kind::i8, accumulators in TMEM).
//
// The reference's dense Sigma contractions (eqf_vio/src/VIOFilter.cpp:188-189, 276-277, 297) are fp64 GEMMs and fp64 has no
// tcgen05 kind: the native path (dgemm_sm100.cu) runs on DMMA.8x8x4, whose pipe tops out at 37.1 TFLOP/s on B200 — the Riccati
// kernel already sits at 90 % of it.  B200's 5th-generation tensor cores do 8-bit integer MMA at ~100x that rate, and an fp64
// product can be assembled EXACTLY from integer products (Ozaki splitting):
//
//   a_ik = 2^ea_i * sum_s A_s[i,k] 2^(-6-7s),   b_kj = 2^eb_j * sum_t B_t[k,j] 2^(-6-7t),   A_s, B_t in [-64, 64] (int8),
//   ea_i / eb_j the binary exponent of the largest entry of row i of A / column j of B,
//   c_ij = 2^(ea_i + eb_j - 12) * sum_d 2^(-7d) * ( sum_{s+t=d} A_s B_t )_ij          d = 0 .. S-1
//
// every inner product A_s B_t is an int8 GEMM accumulated exactly in int32 (K * 64 * 64 * S < 2^31), all pairs of one diagonal d
// share one TMEM accumulator, and only the S per-diagonal sums are converted to fp64.  With S = 8 slices (55 mantissa bits) and the
// 36 pairs s + t <= 7 the result is as accurate as an fp64 GEMM relative to (row max of A) x (column max of B); S = 9 (45 pairs)
// matches fp64 entry by entry on operands with Sigma's dynamic range.
//
// Kernel shape: one CTA per 128 x 128 output tile; warps 0-7 epilogue (TMEM -> registers -> fp64), warp 8 TMA producer, warp 9 MMA
// issuer.  The diagonals are processed in batches of four (4 x 128 TMEM columns = all 512); within a batch every 32-deep k-block of
// ALL slices needed is staged once (cp.async.bulk of one contiguous, pre-swizzled run of 4 KB slice tiles per operand — the first
// version fetched 128 separate 32-byte rows per slice tile through a tensor map and reached a third of L2's rate) and every (s, t)
// pair of the batch is issued against it, so a slice tile is fetched from L2 once per batch instead of once per pair (36 pairs -> 2
// fetches: the int8 MMA rate would need 125 B/clk/SM otherwise, three times what L2 delivers).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace eqvio {

static const int OZ_MAX_SLICES = 9;
static const int OZ_TILE = 128;      // output tile (M and N), also the row padding of the slice arrays
static const int OZ_KBLOCK = 32;     // int8 MMA depth = one SWIZZLE_32B row

// One operand split into int8 slices, tiled for the kernel's staging: slices[row tile][k-block][s][128 rows x 32 bytes], each 4 KB
// slice tile K-major and pre-swizzled (the SWIZZLE_32B image the tensor core's shared-memory descriptor expects), so that the S' <= S
// slice tiles a batch of diagonals needs for one k-block are one contiguous run fetched by a single bulk copy.  Rows padded to 128,
// k to 32, zero-filled.  ex[row] is the row's binary exponent.  "row" is a row of A or a COLUMN of B (both operands are K-major).
struct OzOperand {
    int8_t* slices;   // S * rows_pad * k_pad bytes
    int* ex;          // rows_pad ints
    int rows, k, rows_pad, k_pad, S;
    int ex_margin;    // added to every exponent by the split and by the product's epilogue (exponents taken from a transposed, nearly symmetric matrix)
};

// Epilogue of the Riccati step's second product (VIOFilter.cpp:188-189) for the core block: D += sum_c Wx[gm, c] Fx[gn, c] (the rank-6
// term T B_b R B_b^T: Wx = T B_b R and Fx = B_b are six columns with leading dimension ldx) and + T P on the diagonal; gm / gn =
// tile-local row / column + row_off / col_off = index in Sigma.
struct OzRiccatiEpilogue {
    int on, row_off, col_off, ldx;
    const double* T_dev;
    const double* Wx;
    const double* Fx;
    double Pd[5];
};

// Inner-dimension equilibration.  The digit grid of a row is relative to the row's largest entry, so an output entry's error is
// ~2^-53 x (row max of A) x (column max of B) — poor for the small entries of a covariance whose variances span eight decades.
// Scaling the inner index by powers of two, A' = A 2^(+h), B' = 2^(-h) B with 2^h[k] ~ sqrt(Sigma_kk), is exact, cancels in the
// product, and makes both maxima ~ sqrt(variance): the error becomes relative to sqrt(Sigma_ii Sigma_jj), the natural scale of entry
// (i, j).  `sign` = +1 / -1 selects which side this operand is.
struct OzKScale { const int* h; int sign; };
// Exponent maxima of the output (for splitting it as the next product's operand without a separate pass): rows_out[row] over the
// columns, each entry scaled by 2^(-h[col + col_off]); cols_out[col] over the rows, scaled by 2^(-h[row + row_off]).  Either may be
// null; the arrays must have been reset (oz_reset_exponents) and are atomically maxed.
struct OzExponentsOut { int* rows_out; int* cols_out; const int* h; int row_off, col_off; };

size_t oz_slices_bytes(int rows, int k, int S);
// Splits X (element (r, k) at X[r * stride_r + k * stride_k]) into S slices; `slices` sized with oz_slices_bytes.  ex_ready: `ex`
// already holds the rows' exponents (of the scaled entries); else they are computed first (one more pass over X).
// k_rot > 0: the inner index is stored rotated, k' = k - k_rot for k >= k_rot and k' = (k - k_rot) + k for the first k_rot (they go
// behind the others); k - k_rot must be a multiple of 128.
cudaError_t oz_split(const double* X, long stride_r, long stride_k, int rows, int k, int S, OzOperand* op, int8_t* slices, int* ex,
                     cudaStream_t stream, const OzKScale* ks = nullptr, bool ex_ready = false, int ex_margin = 0, int k_rot = 0);
cudaError_t oz_reset_exponents(int* ex, int count, cudaStream_t stream);
// ex[r] = max(ex[r], largest exponent of row r of X over its k columns)
cudaError_t oz_rowmax(const double* X, long stride_r, long stride_k, int rows, int k, int* ex, cudaStream_t stream, const OzKScale* ks = nullptr);
// h[k] = floor(exponent(Sigma_kk) / 2)
cudaError_t oz_diag_scale(const double* Sigma, int ld, int n, int* h, cudaStream_t stream);
// D[0:M, 0:N] (column-major, ldd) = alpha * A * B + beta * Cin for the split operands (A: M rows, B: N "rows" = columns of B), M and N
// multiples of 128 up to the padding (rows beyond M / N are computed on zero slices and not stored).
// ksplit = 2: every tile's k-blocks are shared by two CTAs that ADD their fp64 results into D, which the caller has zeroed (two terms:
// the order of the additions cannot matter) — for products with too few tiles to fill the GPU; no Cin / ric / exo then.
cudaError_t oz_gemm(const OzOperand& A, const OzOperand& B, int M, int N, double alpha, double beta, const double* Cin, int ldcin, double* D,
                    int ldd, cudaStream_t stream, const OzRiccatiEpilogue* ric = nullptr, const OzExponentsOut* exo = nullptr, int ksplit = 1);
cudaError_t oz_init_device();
// C (m x n = 2N x (11 + 3N), one 2 x 3 block per landmark) split from its structural entries: its rows (inner index = column, rotated
// by m0, scaled by `ks`) or its columns [m0, n) (inner index = row).  `slices` must be zero outside the structural positions (they
// depend on m, n only).
cudaError_t oz_split_C_rows(const double* C, int ldm, int m, int n, int m0, int S, const OzKScale* ks, OzOperand* op, int8_t* slices, int* ex, cudaStream_t stream);
cudaError_t oz_split_C_cols(const double* C, int ldm, int m, int n, int m0, int S, OzOperand* op, int8_t* slices, int* ex, cudaStream_t stream);

// ---- the Riccati step as two launches of one kernel ------------------------------------------------------------------------------
// Sigma (n x n, n = m0 + Mc) = [border: the first m0 = 11 + (3N mod 128) rows / columns, i.e. at least the base states | block: Mc = 128 Mt of landmark states].  Both products of
// VIOFilter.cpp:188-189 are computed for the block on the int8 tensor cores with F's rows as the A operand:
//   phase 1:  D[i, c] = W[i, c]       = sum_k F[i, k] Sigma[k, c]         B operand = Sigma's columns  (X = Sigma, Out = W)
//   phase 2:  D[j, i] = Sigma'[i, j]  = sum_k F[j, k] W[i, k] + ...       B operand = W's rows         (X = W, Out = Sigma', stored transposed)
// so that in both phases the rows of D are the operand rows of the NEXT product (W's rows for phase 2, Sigma''s columns for the next
// step's phase 1) and the epilogue writes them out as int8 slices directly: every tile publishes its rows' exponent maxima, waits for
// the other tiles of its tile row (tiles are taken by ticket in row-major order, so the wait cannot deadlock at any grid size) and then
// emits its 128 x 128 block of digits from registers — no separate split pass, exact exponents.  The inner index of all slice arrays
// is rotated, k' = k - m0 for the block and k' = Mc + k for the border, so that a tile's columns are whole 32-deep k-blocks.
// The border rows / columns (2 m0 n outputs per phase, 0.7 % of the work at N = 512) are fp64 dot products done by the epilogue
// warps while the tensor core runs ("border jobs"); those that are inner-border entries of operand rows are emitted as well.
// Synchronisation words and exponent arrays exist twice (by tick parity): each launch clears the set the other parity uses next.
struct OzFusedParams {
    int Mc, m0, n, n16, KB, Mt, phase, ld;
    const int8_t* slA; const int* exA;      // F rows [m0, n), scaled 2^(+h)
    const int8_t* slB; const int* exB;      // phase 1: Sigma columns [m0, n); phase 2: W rows [m0, n); scaled 2^(-h)
    const double* F;                        // [F | B_b] (B_b in columns n16 .. n16+5)
    const double* X;                        // phase 1: Sigma; phase 2: [W | T B_b R]
    double* Out;                            // phase 1: W; phase 2: Sigma'
    int8_t* slOut; int* exOut;              // the output as the next product's B operand
    const int* h;
    int* exReset;                           // Mc exponents to clear
    int* sync;                              // [0] ticket counter, [1 + tile row] arrivals; zero at launch
    int* syncReset;                         // 1 + Mt words to clear
    const double* T_dev;                    // phase 2: the step length (device memory, written by k_step_prepare)
    double Pd[5];                           // process variances: bias omega, bias accel, gravity, velocity, point
    int* err;                               // device flags word: bit 1 (FLAG_NAN) is raised if an in-kernel wait gives up after about a second
    long long* stamps;                      // diagnostics (may be null): OZ_STAMPS clock64 stamps per CTA, see tools/oz_stamps.py
};
static const int OZ_STAMPS = 16;
static const int OZ_FUSED_SYNC_INTS = 64;   // >= 1 + Mt
bool oz_fused_supported(int S, int Mt);
// pdl: programmatic dependent launch behind the previous kernel of the stream (the prologue overlaps that kernel's tail)
cudaError_t oz_riccati_fused(const OzFusedParams& p, int S, cudaStream_t stream, bool pdl = false);
// F's rows [m0, n) as slices from their nine structural entries per row (see k_oz_split_F_rows); `slices` must be zero elsewhere.
cudaError_t oz_split_F_rows(const double* F, int ld, int n, int m0, int S, const int* h, int8_t* slices, int* ex, cudaStream_t stream);

}  // namespace eqvio
