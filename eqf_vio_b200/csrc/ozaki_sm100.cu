// ozaki_sm100.cu — see ozaki_sm100.cuh.  Hand-written sm_100a: tcgen05.mma kind::i8 (SASS UTCIMMA / UTCQMMA family) with int32
// accumulators in TMEM, tcgen05.ld epilogue (LDTM), TMA SWIZZLE_32B operand staging (UTMALDG), mbarrier pipelines.
#include "ozaki_sm100.cuh"

#include <cudaTypedefs.h>
#include <limits.h>
#include <stdio.h>
#include <string.h>

namespace eqvio {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
namespace oz {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) {}
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// contiguous global -> shared bulk copy (no tensor map), completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
// ---- tcgen05 ----
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// all MMAs issued so far by this thread arrive on the mbarrier when they have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one lane of a converged warp (the warp stays converged around it: address arithmetic stays on the uniform datapath)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32, M = 128, N = 128, K = 32
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread `lane` of the warp receives its lane's 32 consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major operand tile staged by TMA with SWIZZLE_32B: rows of 32 bytes (= the 32-deep int8
// k-block), 8-row groups 256 bytes apart (stride byte offset), leading byte offset unused for swizzled K-major layouts (1),
// descriptor version 1 (sm_100), layout type 6 = SWIZZLE_32B.
__device__ __forceinline__ uint64_t smem_desc_sw32(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}
}  // namespace oz

// Instruction descriptor: dense, no saturate, D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), both K-major (bits 15, 16 = 0),
// N = 128 (N >> 3 at bit 17), M = 128 (M >> 4 at bit 24).
static constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_TILE >> 3) << 17) | ((uint32_t)(OZ_TILE >> 4) << 24);

static constexpr int OZ_SLICE_TILE_BYTES = OZ_TILE * OZ_KBLOCK;                     // 4096
static constexpr int OZ_STAGES = 3;
static constexpr int OZ_DIAGS_PER_BATCH = 4;   // 4 x 128 TMEM columns
static constexpr int OZ_SPLIT_SMEM = 32 * 129 * 8;

struct OzParams {
    int M, N, KB, S;
    const int* exA;
    const int* exB;
    int marginA, marginB;
    double alpha, beta;
    const double* Cin;
    int ldcin;
    double* D;
    int ldd;
    OzRiccatiEpilogue ric;
    OzExponentsOut exo;
    int ksplit;   // 1, or 2: every tile's k-blocks are shared by two CTAs whose fp64 results are added into a zeroed D (a + b = b + a: the order cannot matter)
};

__device__ __forceinline__ double oz_pow2(int e) {   // 2^e for e in the normal range
    return __hiloint2double((e + 1023) << 20, 0);
}
__device__ __forceinline__ int oz_exponent(double x);

// Warp roles: sixteen epilogue warps (lane quarter = warp % 4, column quarter = warp / 4: 32 accumulator columns per thread — the epilogue
// is dependent fp64 / integer chains, more warps hide them; the first version had eight warps of 64 columns, 168 registers and spills,
// and took 30 us of a 100 us launch), one producer, one MMA issuer.
static constexpr int OZG_EPI_WARPS = 16, OZG_PRODUCER_WARP = 16, OZG_MMA_WARP = 17, OZG_THREADS = 576;
static constexpr int OZG_AUX_BYTES = 128 * 8 + 128 * 4 + 6 * 128 * 8;   // per tile column: 2^(exB + margin), h (exponent outputs), B_b (Riccati epilogue)
template <int S> struct OzGemmCfg {
    static constexpr int STAGE_BYTES = 2 * S * OZ_SLICE_TILE_BYTES;   // A slices then B slices
    static constexpr int PIPE_BYTES = OZ_STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = PIPE_BYTES + OZG_AUX_BYTES + 1024 + 256;
};

template <int S>
__global__ void __launch_bounds__(OZG_THREADS, 1)
k_oz_gemm(const int8_t* __restrict__ slA, const int8_t* __restrict__ slB, const OzParams p) {
    using namespace oz;
    using Cfg = OzGemmCfg<S>;
    extern __shared__ uint8_t oz_smem_raw[];
    const uint32_t raw = smem_u32(oz_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = oz_smem_raw + (base - raw);
    double* s_cb = reinterpret_cast<double*>(smem + Cfg::PIPE_BYTES);   // 2^(exB[col] + marginB)
    double* s_fx = s_cb + 128;                                          // [6][128] B_b rows of the tile's columns (Riccati epilogue)
    int* s_h = reinterpret_cast<int*>(s_fx + 6 * 128);                  // h[col + col_off] (exponent outputs)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::PIPE_BYTES + OZG_AUX_BYTES);
    uint64_t* empty = full + OZ_STAGES;
    uint64_t* acc_full = empty + OZ_STAGES;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Mt = (p.M + OZ_TILE - 1) / OZ_TILE;
    const int tile_id = (int)blockIdx.x / p.ksplit, khalf = (int)blockIdx.x - tile_id * p.ksplit;
    const int tile_m = tile_id % Mt, tile_n = tile_id / Mt;
    const int kb_lo = khalf * ((p.KB + p.ksplit - 1) / p.ksplit), kb_hi = min(p.KB, kb_lo + (p.KB + p.ksplit - 1) / p.ksplit);
    const int KB = kb_hi - kb_lo;   // this CTA's k-blocks
    constexpr int nbatch = (S + OZ_DIAGS_PER_BATCH - 1) / OZ_DIAGS_PER_BATCH;

    if (threadIdx.x == 0) {
        for (int s = 0; s < OZ_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, OZG_EPI_WARPS);
        mbar_fence_init();
    }
    if (warp == OZG_MMA_WARP) {   // one warp allocates all 512 TMEM columns (one CTA per SM) and frees them at the end
        tmem_alloc(smem_u32(tmem_slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == OZG_PRODUCER_WARP) {
        if (lane == 0) {
            // the slice arrays are tiled [row tile][k-block][slice][128 rows x 32 B, pre-swizzled]: the nS slice tiles a batch needs for one
            // k-block are ONE contiguous run of nS x 4 KB per operand — two bulk copies per stage
            const int8_t* gA = slA + ((size_t)tile_m * p.KB + kb_lo) * S * OZ_SLICE_TILE_BYTES;
            const int8_t* gB = slB + ((size_t)tile_n * p.KB + kb_lo) * S * OZ_SLICE_TILE_BYTES;
            uint32_t it = 0;
            for (int b = 0; b < nbatch; ++b) {
                const int nS = min(OZ_DIAGS_PER_BATCH * (b + 1), S);   // slices 0 .. nS-1 of both operands take part in this batch's diagonals
                const uint32_t bytes = (uint32_t)(nS * OZ_SLICE_TILE_BYTES);
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % OZ_STAGES;
                    mbar_wait(&empty[s], ((it / OZ_STAGES) & 1) ^ 1);
                    mbar_expect_tx(&full[s], 2 * bytes);
                    const uint32_t sa = base + s * Cfg::STAGE_BYTES, sb = sa + S * OZ_SLICE_TILE_BYTES;
                    bulk_load(sa, gA + (size_t)kb * S * OZ_SLICE_TILE_BYTES, bytes, &full[s]);
                    bulk_load(sb, gB + (size_t)kb * S * OZ_SLICE_TILE_BYTES, bytes, &full[s]);
                }
            }
        }
    } else if (warp == OZG_MMA_WARP) {
        // The whole warp walks the loop (waits included) and ONE elected lane issues: with a single lane inside a divergent branch
        // every descriptor had to be moved from vector to uniform registers per instruction and the issue loop ran at ~160 clk per MMA
        // against the 64 clk the tensor core needs (first version: 43 % tensor-pipe active, producer never the one waited for).
        uint32_t it = 0;
        const uint64_t adesc0 = smem_desc_sw32(base), bdesc0 = smem_desc_sw32(base + S * OZ_SLICE_TILE_BYTES);
#pragma unroll
        for (int b = 0; b < nbatch; ++b) {
            constexpr int DPB = OZ_DIAGS_PER_BATCH;
            const int dmin = DPB * b, dmax = (dmin + DPB - 1 < S - 1) ? dmin + DPB - 1 : S - 1;
            if (b > 0) {   // the epilogue has read the previous batch's accumulators out of TMEM
                mbar_wait(acc_empty, (uint32_t)((b - 1) & 1));
                tc_fence_after();
            }
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const int s = it % OZ_STAGES;
                mbar_wait(&full[s], (it / OZ_STAGES) & 1);
                tc_fence_after();
                const uint64_t soff = (uint64_t)((s * Cfg::STAGE_BYTES) >> 4);
                const uint32_t first = kb > 0 ? 1u : 0u;
                if (elect_one()) {
#pragma unroll
                    for (int d = dmin; d <= dmax; ++d) {
                        const uint32_t acc = tmem_base + (uint32_t)((d - dmin) * OZ_TILE);
#pragma unroll
                        for (int a = 0; a <= d; ++a)   // all pairs (a, d - a) of this diagonal into one int32 accumulator
                            mma_i8(acc, adesc0 + soff + (uint64_t)(a * (OZ_SLICE_TILE_BYTES >> 4)), bdesc0 + soff + (uint64_t)((d - a) * (OZ_SLICE_TILE_BYTES >> 4)),
                                   OZ_IDESC, a > 0 ? 1u : first);
                    }
                    tc_commit(&empty[s]);       // the stage is free once these MMAs have read it
                }
                __syncwarp();
            }
            if (elect_one()) tc_commit(acc_full);   // the batch's accumulators are complete
            __syncwarp();
        }
    } else {
        // ===== epilogue warps: lane quarter q (TMEM lanes 32q .. 32q+31 = tile rows), column quarter cq =====
        const int tid = threadIdx.x;
        const int q = warp & 3, cq = warp >> 2;
        const bool ric = p.ric.on != 0;
        // per-column factors of this tile, staged once while the tensor core works
        if (tid < OZ_TILE) {
            const int col = tile_n * OZ_TILE + tid;
            s_cb[tid] = col < p.N ? oz_pow2(max(p.exB[col], -900) + p.marginB) : 0.0;
            s_h[tid] = (p.exo.h && col < p.N) ? p.exo.h[col + p.exo.col_off] : 0;
        } else if (ric && tid < OZ_TILE + 6 * 32) {
            const int c = (tid - OZ_TILE) >> 5, l = tid & 31;
            for (int j = l; j < OZ_TILE; j += 32) {
                const int col = tile_n * OZ_TILE + j;
                s_fx[c * 128 + j] = col < p.N ? p.ric.Fx[(size_t)(col + p.ric.col_off) + (size_t)p.ric.ldx * c] : 0.0;
            }
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        double acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.0;
        for (int b = 0; b < nbatch; ++b) {
            const int dmin = OZ_DIAGS_PER_BATCH * b, dmax = min(dmin + OZ_DIAGS_PER_BATCH - 1, S - 1);
            mbar_wait(acc_full, (uint32_t)(b & 1));
            tc_fence_after();
            for (int d = dmin; d <= dmax; ++d) {
                const double scale = oz_pow2(-7 * d);
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((d - dmin) * OZ_TILE + cq * 32), v);
                tmem_ld_wait();
                // int32 -> fp64 without the conversion unit: 2^52 + 2^31 + x has x + 2^31 in its low mantissa word
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = fma(__hiloint2double(0x43300000, (int)(v[j] ^ 0x80000000u)) - 4503601774854144.0, scale, acc[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
        const int row = tile_m * OZ_TILE + q * 32 + lane;
        const int lc0 = cq * 32, col0 = tile_n * OZ_TILE + lc0;
        const bool row_ok = row < p.M;
        const double ra = row_ok ? oz_pow2(max(p.exA[row], -900) + p.marginA - 12) : 0.0;
        // Riccati epilogue (VIOFilter.cpp:188-189): + T (B_b R B_b^T) as six products of the border columns, + T P on the diagonal
        const int gm = row + p.ric.row_off;
        double wx[6] = {0, 0, 0, 0, 0, 0}, Tstep = 0.0;
        if (ric && row_ok) {
            Tstep = *p.ric.T_dev;
#pragma unroll
            for (int c = 0; c < 6; ++c) wx[c] = p.ric.Wx[(size_t)gm + (size_t)p.ric.ldx * c];
        }
        int emax = -2000;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            double v = 0.0;
            if (row_ok && col < p.N) {
                v = p.alpha * ((acc[j] * ra) * s_cb[lc0 + j]);
                if (p.beta != 0.0) v += p.beta * p.Cin[(size_t)row + (size_t)p.ldcin * col];
                if (ric) {
                    const int gn = col + p.ric.col_off;
                    double r6 = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) r6 = fma(wx[c], s_fx[c * 128 + lc0 + j], r6);
                    v += r6;
                    if (gm == gn) v += Tstep * (gm < 3 ? p.ric.Pd[0] : gm < 6 ? p.ric.Pd[1] : gm < 8 ? p.ric.Pd[2] : gm < 11 ? p.ric.Pd[3] : p.ric.Pd[4]);
                }
                if (p.ksplit > 1) atomicAdd(p.D + (size_t)row + (size_t)p.ldd * col, v);
                else p.D[(size_t)row + (size_t)p.ldd * col] = v;
                emax = max(emax, oz_exponent(v) - s_h[lc0 + j]);
            }
            acc[j] = v;
        }
        // Exponents of what was just written, for the split of this output as the next product's operand: row maxima over the columns
        // (each entry scaled by 2^(-h[column]) first) and / or column maxima over the rows (scaled by 2^(-h[row])).
        if (p.exo.rows_out != nullptr && row_ok) atomicMax(p.exo.rows_out + row, emax);
        if (p.exo.cols_out != nullptr) {
            const int hr = (p.exo.h && row_ok) ? p.exo.h[row + p.exo.row_off] : 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                int e = row_ok ? oz_exponent(acc[j]) - hr : -2000;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) e = max(e, __shfl_xor_sync(0xffffffffu, e, o));
                if (lane == 0 && col0 + j < p.N) atomicMax(p.exo.cols_out + col0 + j, e);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == OZG_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// splitting
// ------------------------------------------------------------------------------------------------
// Every entry may be scaled by a power of two that depends on its k index, x'(r, k) = x(r, k) 2^(hsign h[k]): the two operands of a
// product use opposite signs, so the scales cancel exactly in the product while the digit grids (relative to the row maximum of
// the SCALED entries) stop being dominated by a few large-variance states (see OzKScale in the header).
__device__ __forceinline__ double oz_kscaled(double x, const int* h, int hsign, int kidx) {
    return h ? x * oz_pow2(hsign * h[kidx]) : x;
}
// stored (rotated) inner index -> source index: the first krot source indices are stored behind the other k - krot
__device__ __forceinline__ int oz_unrot(int kk, int k, int krot) { return kk < k - krot ? kk + krot : kk - (k - krot); }
template <int KW>   // tile of 32 rows x KW k (KW = 32 or 128), t[32][KW + 1]; k0 is a STORED index
__device__ __forceinline__ void oz_load_tile(const double* X, long sr, long sk, int r0, int k0, int rows, int k, const int* h, int hsign,
                                             double (*t)[KW + 1], int krot = 0) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    if (sk == 1) {   // k contiguous: a warp reads 32 consecutive k of one row
#pragma unroll
        for (int i = 0; i < KW / 8; ++i) {
            const int r = ty + 8 * (i & 3), kk = tx + 32 * (i >> 2);
            const int ks = oz_unrot(k0 + kk, k, krot);
            t[r][kk] = (r0 + r < rows && k0 + kk < k) ? oz_kscaled(X[(size_t)(r0 + r) * sr + ks], h, hsign, ks) : 0.0;
        }
    } else {         // rows contiguous (or general): a warp reads 32 consecutive rows at one k
#pragma unroll
        for (int i = 0; i < KW / 8; ++i) {
            const int kk = ty + 8 * i;
            const int ks = oz_unrot(k0 + kk, k, krot);
            t[tx][kk] = (r0 + tx < rows && k0 + kk < k) ? oz_kscaled(X[(size_t)(r0 + tx) * sr + (size_t)ks * sk], h, hsign, ks) : 0.0;
        }
    }
}
__device__ __forceinline__ int oz_exponent(double x) {   // e with |x| = f 2^e, f in [0.5, 1); zero / denormal: very small
    const int hi = __double2hiint(x);
    const int be = (hi >> 20) & 0x7ff;
    return be == 0 ? -2000 : be - 1022;
}

// ex[r] = max(ex[r], exponent of the largest scaled entry of row r over k in [0, k))
__global__ void __launch_bounds__(256) k_oz_rowmax(const double* X, long sr, long sk, int rows, int k, const int* h, int hsign, int* ex) {
    __shared__ double t[32][33];
    const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    oz_load_tile<32>(X, sr, sk, r0, k0, rows, k, h, hsign, t);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = threadIdx.y + 8 * i;
        int e = oz_exponent(t[r][threadIdx.x]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e = max(e, __shfl_xor_sync(0xffffffffu, e, o));
        if (threadIdx.x == 0 && r0 + r < rows) atomicMax(ex + r0 + r, e);
    }
}

// digits: x 2^-e 2^6 = d0 + r0, |r| <= 1/2; then r 2^7 = d + r' ... every step exact in fp64, every digit in [-64, 64].
// A block takes 32 rows x 128 k; a thread 16 consecutive k of one row, one 16-byte store per slice (a warp: four full 128-byte lines).
__global__ void __launch_bounds__(256) k_oz_split(const double* X, long sr, long sk, int rows, int k, const int* h, int hsign, int* ex,
                                                  int ex_margin, int8_t* slices, int rows_pad, int k_pad, int S, int krot) {
    extern __shared__ double oz_split_smem[];
    double(*t)[129] = reinterpret_cast<double(*)[129]>(oz_split_smem);
    const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 128;
    oz_load_tile<128>(X, sr, sk, r0, k0, rows, k, h, hsign, t, krot);
    __syncthreads();
    const int id = threadIdx.y * 32 + threadIdx.x, r = id >> 3, kg = id & 7;
    if (r0 + r >= rows || k0 + 16 * kg >= k_pad) return;
    // ex_margin: the exponents came from the transposed entries of a matrix that is symmetric only up to round-off; widened by the
    // margin here and, identically, in the product's epilogue (OzOperand::ex_margin)
    const int e = max(ex[r0 + r], -900) + ex_margin;
    const double up = oz_pow2(6 - e);
    double v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = t[r][16 * kg + c] * up;
    // tiled layout [row tile][k-block][slice][128 rows x 32 B]; inside a 4 KB slice tile the image is what TMA's SWIZZLE_32B would
    // have produced (16-byte halves of a row exchanged in rows 4-7 of every group of eight), so a plain bulk copy stages it
    const int row = r0 + r, kk = k0 + 16 * kg;
    const int KB = k_pad / OZ_KBLOCK;
    const size_t tile = ((size_t)(row / OZ_TILE) * KB + kk / OZ_KBLOCK) * S;
    const int in_tile = (row % OZ_TILE) * OZ_KBLOCK + (((kk % OZ_KBLOCK) >> 4) ^ ((row >> 2) & 1)) * 16;
    uint4* dst = reinterpret_cast<uint4*>(slices + tile * OZ_SLICE_TILE_BYTES + in_tile);
    constexpr size_t plane = OZ_SLICE_TILE_BYTES;
    (void)rows_pad;
    // round to nearest by adding 1.5 * 2^52: the digit appears as a two's-complement integer in the low word of the sum and as a double
    // after subtracting the constant again — two full-rate additions instead of FRND + F2I (the first version of this kernel took 36 us
    // for a 19 MB operand, four times what its memory traffic costs)
    const double magic = 6755399441055744.0;
    for (int s = 0; s < S; ++s) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const double t = v[c] + magic;
            const double d = t - magic;
            v[c] = (v[c] - d) * 128.0;
            w[c >> 2] |= ((uint32_t)__double2loint(t) & 0xffu) << (8 * (c & 3));
        }
        dst[(size_t)s * plane / 16] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// h[k] = half the binary exponent of the diagonal entry (k, k): 2^h[k] ~ sqrt(Sigma_kk)
__global__ void k_oz_diag_scale(const double* Sigma, int ld, int n, int* h) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double d = Sigma[(size_t)k * (ld + 1)];
    const int e = (d > 0.0) ? oz_exponent(d) : 0;
    h[k] = (e >= -1000 && e <= 1000) ? (e >= 0 ? e / 2 : -((-e + 1) / 2)) : 0;
}

// ------------------------------------------------------------------------------------------------
// The Riccati step as two launches of ONE kernel (fused form, see OzFusedParams in the header)
// ------------------------------------------------------------------------------------------------
static constexpr int OZF_YC = 12;                                  // short side of a border job (covers the 11 base states in one piece)
static constexpr int OZ_EX_RESET = (int)0xC0C0C0C0;               // what oz_reset_exponents' memset leaves: a very negative exponent

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Bounded wait for a counter other CTAs of the launch advance (release side: __threadfence + atomicAdd).  Tiles are taken by ticket, so
// the wait ends as soon as the hardware has run the CTAs in front; if it does not within ~2^31 clocks (about a second: a fault in
// another CTA, the context losing its SMs) the wait gives up and raises the filter's NaN flag instead of hanging the GPU — the
// launch's results are then invalid and the host sees EQVIO_ERR_NAN at its next status read.
__device__ __forceinline__ void oz_wait_ge(const int* ctr, int target, int* err) {
    if (ld_acquire(ctr) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire(ctr) < target) {
        __nanosleep(100);
        if (clock64() - t0 > (1LL << 31)) { if (err) atomicOr(err, 2); return; }   // (2 = FLAG_NAN of the filter's device flags)
    }
}

// Border outputs D[fr, o] = sum_k F[fr, k] B(o, k) (B(o, k) = Sigma[k, o] in phase 1, W[o, k] in phase 2).  Row fr of F = I + T A_b is
// zero outside the 11 base columns and — for a landmark row — its own 3 x 3 block (EqFMatrices.cpp:289-312: the only couplings are
// bias / gravity / velocity -> everything and landmark -> itself), so the sum over all k has at most 14 terms that are not exact zeros;
// they are added in ascending k like the dense loop would.
__device__ __forceinline__ double oz_border_dot(const OzFusedParams& p, int fr, int o) {
    const double* Frow = p.F + fr;
    const bool ph1 = p.phase == 1;
    const double* B0 = ph1 ? p.X + (size_t)p.ld * o : p.X + o;     // B(o, k) at B0[k * bs]
    const size_t bs = ph1 ? 1 : (size_t)p.ld;
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 11; ++k) v = fma(Frow[(size_t)p.ld * k], B0[bs * k], v);
    if (fr >= 11) {
        const int b0 = 11 + (fr - 11) / 3 * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) v = fma(Frow[(size_t)p.ld * (b0 + k)], B0[bs * (b0 + k)], v);
    }
    return v;
}
__device__ __forceinline__ double oz_process_noise(const OzFusedParams& p, int g) {
    return g < 3 ? p.Pd[0] : g < 6 ? p.Pd[1] : g < 8 ? p.Pd[2] : g < 11 ? p.Pd[3] : p.Pd[4];
}

// Digits of 16 values x (|x| <= 64) as S slices of one 16-byte chunk each.  Slices are taken four at a time from one value:
// adding M_s = 1.5 * 2^(52 - 7s) rounds x to a multiple of 2^(-7s) and leaves R_s = round(x 2^(7s)) as a two's-complement integer in
// the low word of the sum, and digit_s = R_s - 128 R_(s-1) (|digit| <= 64) is the balanced base-128 digit the sequential algorithm
// (k_oz_split) produces — one fp64 addition and one integer multiply-add per digit instead of four dependent fp64 operations; after
// four digits the exact residual x - R_3 2^(-21) is rescaled by 2^28 and the next four follow.
__device__ __forceinline__ uint32_t oz_pack4(int a, int b, int c, int d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}
template <int S>
__device__ __forceinline__ void oz_emit16(const double* x, uint4* dst) {
    int dig[S][16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        double v = x[c];
#pragma unroll
        for (int g = 0; g < S; g += 4) {
            int prev = 0;
            double t = 0.0;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (g + s < S) {
                    const double M = __hiloint2double((1023 + 52 - 7 * s) << 20 | 0x80000, 0);   // 1.5 * 2^(52 - 7 s)
                    t = v + M;
                    const int R = __double2loint(t);
                    dig[g + s][c] = s == 0 ? R : R - (prev << 7);
                    prev = R;
                    if (s == 3 && g + 4 < S) v = (v - (t - M)) * 268435456.0;   // 2^28: the residual behind four digits
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < S; ++s)
        dst[(size_t)s * (OZ_SLICE_TILE_BYTES / 16)] = make_uint4(oz_pack4(dig[s][0], dig[s][1], dig[s][2], dig[s][3]), oz_pack4(dig[s][4], dig[s][5], dig[s][6], dig[s][7]),
                                                               oz_pack4(dig[s][8], dig[s][9], dig[s][10], dig[s][11]), oz_pack4(dig[s][12], dig[s][13], dig[s][14], dig[s][15]));
}

// Warp roles of the fused kernel: sixteen epilogue warps (lane quarter = warp % 4, column quarter = warp / 4: 32 accumulator columns per
// thread — the epilogue is dependent fp64 / integer chains, more warps hide them), one producer, one MMA issuer.
static constexpr int OZF_EPI_WARPS = 16, OZF_PRODUCER_WARP = 16, OZF_MMA_WARP = 17, OZF_THREADS = 576, OZF_EPI_THREADS = OZF_EPI_WARPS * 32;
static constexpr int OZF_TR_LD = 33;                               // row pitch (doubles) of the transposed-store staging
static constexpr int OZF_AUX_BYTES = 128 * 8 * 2 + 128 * 4 + 6 * 128 * 8;   // per tile column: 2^exB, 2^-h, h, and (phase 2) T B_b R
template <int S> struct OzFusedCfg {
    static constexpr int STAGE_BYTES = 2 * S * OZ_SLICE_TILE_BYTES;   // A slices then B slices
    static constexpr int STAGES = 3;
    static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = PIPE_BYTES + OZF_AUX_BYTES + 1024 + 256;
    static_assert(OZF_EPI_WARPS * 32 * OZF_TR_LD * 8 <= PIPE_BYTES, "transposed-store staging reuses the pipeline stages");
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }   // the sixteen epilogue warps

// The inner border of a tile row as int8 slices (inner indices [0, m0), stored behind the block: k' = Mc + index, zero-padded to a whole
// k-block): done by the producer warp, which is idle once the last stage is in flight, while the epilogue warps emit the tile itself.
// The tile row's 128 rows are shared out among its Mt CTAs.
template <int S>
__device__ __forceinline__ void oz_emit_inner_border(const OzFusedParams& p, int tile_m, int tile_n, int lane) {
    const int Mt = p.Mt, m0 = p.m0, ld = p.ld, KB = p.KB;
    const int rb_lo = tile_n * OZ_TILE / Mt, rb_n = (tile_n + 1) * OZ_TILE / Mt - rb_lo;
    const int nchunk = (KB - p.Mc / OZ_KBLOCK) * 2;           // 16-byte chunks per row
    for (int item = lane; item < rb_n * nchunk; item += 32) {
        const int rb = rb_lo + item % rb_n, ck = item / rb_n;
        const int rowb = tile_m * OZ_TILE + rb, growb = m0 + rowb;
        const int e = max(__ldcg(p.exOut + rowb), -900);
        const double up = oz_pow2(6 - e);
        const int sw = (rb >> 2) & 1;
        double x[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const int idx = ck * 16 + c;                      // border inner index
            double xv = 0.0;
            if (idx < m0) {
                xv = __ldcg(p.phase == 1 ? p.Out + (size_t)growb + (size_t)ld * idx : p.Out + (size_t)idx + (size_t)ld * growb);
                xv = (xv * oz_pow2(-p.h[idx])) * up;
            }
            x[c] = xv;
        }
        const int kk = p.Mc + ck * 16;
        const size_t tile = ((size_t)tile_m * KB + (kk >> 5)) * S;
        const int in_tile = rb * OZ_KBLOCK + ((((kk & 31) >> 4) ^ sw) << 4);
        oz_emit16<S>(x, reinterpret_cast<uint4*>(p.slOut + tile * OZ_SLICE_TILE_BYTES + in_tile));
    }
}

#define OZ_STAMP(i) do { if (p.stamps && lane == 0 && (warp == 0 || warp == OZF_MMA_WARP)) p.stamps[(size_t)ticket * OZ_STAMPS + (i)] = clock64(); } while (0)
template <int S>
__global__ void __launch_bounds__(OZF_THREADS, 1) k_oz_riccati(const OzFusedParams p) {
    using namespace oz;
    using Cfg = OzFusedCfg<S>;
    extern __shared__ uint8_t oz_smem_raw[];
    const uint32_t raw = smem_u32(oz_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = oz_smem_raw + (base - raw);
    double* s_cb = reinterpret_cast<double*>(smem + Cfg::PIPE_BYTES);   // 2^exB[col]
    double* s_ch = s_cb + 128;                                          // 2^-h[col]
    double* s_wx = s_ch + 128;                                          // [6][128] T B_b R rows of the tile's columns (phase 2)
    int* s_h = reinterpret_cast<int*>(s_wx + 6 * 128);                  // h[col]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::PIPE_BYTES + OZF_AUX_BYTES);
    uint64_t* empty = full + Cfg::STAGES;
    uint64_t* acc_full = empty + Cfg::STAGES;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
    int* ticket_slot = reinterpret_cast<int*>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Mt = p.Mt, T = Mt * Mt, KB = p.KB;
    constexpr int nbatch = (S + OZ_DIAGS_PER_BATCH - 1) / OZ_DIAGS_PER_BATCH;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, OZF_EPI_WARPS);
        mbar_fence_init();
    }
    if (warp == OZF_MMA_WARP) {
        tmem_alloc(smem_u32(tmem_slot), 512);
        tmem_relinquish();
    }
    // launched with programmatic stream serialisation (oz_riccati_fused, pdl): everything above overlaps the previous launch's tail;
    // its results (and the words it clears) are only touched behind this wait.  Without the attribute both instructions are no-ops.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        // tiles by ticket: a CTA that waits for its tile row only ever waits for CTAs that started before it or that the hardware can
        // still start (the lowest unfinished tile row is always completely dispatched), whatever the grid size
        *ticket_slot = atomicAdd(p.sync, 1);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int ticket = *ticket_slot;
    const int tile_m = ticket / Mt, tile_n = ticket - tile_m * Mt;
    // the other tick parity's synchronisation words and exponent maxima (nothing running touches them) are cleared for its next launch
    if (ticket == 0)
        for (int i = threadIdx.x; i <= Mt; i += OZF_THREADS) p.syncReset[i] = 0;
    if (tile_n == 0 && threadIdx.x < OZ_TILE) p.exReset[tile_m * OZ_TILE + threadIdx.x] = OZ_EX_RESET;
    OZ_STAMP(0);
    if (p.stamps && threadIdx.x == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); p.stamps[(size_t)ticket * OZ_STAMPS + 11] = (long long)g; }

    if (warp == OZF_PRODUCER_WARP) {
        if (lane == 0) {
            const int8_t* gA = p.slA + (size_t)tile_m * KB * S * OZ_SLICE_TILE_BYTES;
            const int8_t* gB = p.slB + (size_t)tile_n * KB * S * OZ_SLICE_TILE_BYTES;
            uint32_t it = 0;
            for (int b = 0; b < nbatch; ++b) {
                const int nS = min(OZ_DIAGS_PER_BATCH * (b + 1), S);
                const uint32_t bytes = (uint32_t)(nS * OZ_SLICE_TILE_BYTES);
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % Cfg::STAGES;
                    mbar_wait(&empty[s], ((it / Cfg::STAGES) & 1) ^ 1);
                    mbar_expect_tx(&full[s], 2 * bytes);
                    const uint32_t sa = base + s * Cfg::STAGE_BYTES, sb = sa + S * OZ_SLICE_TILE_BYTES;
                    bulk_load(sa, gA + (size_t)kb * S * OZ_SLICE_TILE_BYTES, bytes, &full[s]);
                    bulk_load(sb, gB + (size_t)kb * S * OZ_SLICE_TILE_BYTES, bytes, &full[s]);
                }
            }
        }
        __syncwarp();
        // the tile-row barrier, observed from this warp: the rows' exponents and the inner-border values are final
        {
            const int nch = (p.m0 + OZF_YC - 1) / OZF_YC;
            oz_wait_ge(p.sync + 1 + tile_m, Mt + 4 * nch, p.err);
        }
        oz_emit_inner_border<S>(p, tile_m, tile_n, lane);
    } else if (warp == OZF_MMA_WARP) {
        uint32_t it = 0;
        const uint64_t adesc0 = smem_desc_sw32(base), bdesc0 = smem_desc_sw32(base + S * OZ_SLICE_TILE_BYTES);
#pragma unroll
        for (int b = 0; b < nbatch; ++b) {
            constexpr int DPB = OZ_DIAGS_PER_BATCH;
            const int dmin = DPB * b, dmax = (dmin + DPB - 1 < S - 1) ? dmin + DPB - 1 : S - 1;
            if (b > 0) {
                OZ_STAMP(12);
                mbar_wait(acc_empty, (uint32_t)((b - 1) & 1));
                tc_fence_after();
                OZ_STAMP(13);
            }
            long long waited = 0;
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const int s = it % Cfg::STAGES;
                const long long w0 = p.stamps ? clock64() : 0;
                mbar_wait(&full[s], (it / Cfg::STAGES) & 1);
                if (p.stamps) waited += clock64() - w0;
                tc_fence_after();
                const uint64_t soff = (uint64_t)((s * Cfg::STAGE_BYTES) >> 4);
                const uint32_t first = kb > 0 ? 1u : 0u;
                if (elect_one()) {
#pragma unroll
                    for (int d = dmin; d <= dmax; ++d) {
                        const uint32_t acc = tmem_base + (uint32_t)((d - dmin) * OZ_TILE);
#pragma unroll
                        for (int a = 0; a <= d; ++a)
                            mma_i8(acc, adesc0 + soff + (uint64_t)(a * (OZ_SLICE_TILE_BYTES >> 4)), bdesc0 + soff + (uint64_t)((d - a) * (OZ_SLICE_TILE_BYTES >> 4)),
                                   OZ_IDESC, a > 0 ? 1u : first);
                    }
                    tc_commit(&empty[s]);
                }
                __syncwarp();
            }
            if (elect_one()) tc_commit(acc_full);
            __syncwarp();
            if (p.stamps && lane == 0) p.stamps[(size_t)ticket * OZ_STAMPS + 9 + b] = waited;   // 9, 10: cycles the issuer waited for operands
            if (b == nbatch - 1) OZ_STAMP(14);
        }
    } else {
        // ===== the sixteen epilogue warps =====
        const int tid = threadIdx.x;            // 0 .. 511
        const int q = warp & 3, cq = warp >> 2;
        const int m0 = p.m0, n = p.n, ld = p.ld;
        const double* Wx = (p.phase == 2 ? p.X : p.Out) + (size_t)p.n16 * ld;   // T B_b R (six columns behind W)
        const double* Fx = p.F + (size_t)p.n16 * ld;                            // B_b
        const double Tstep = p.phase == 2 ? *p.T_dev : 0.0;
        int* cnt = p.sync + 1;

        // ---- per-column factors of this tile, staged once: 2^exB, 2^-h, h and (phase 2) the six entries of T B_b R
        if (tid < OZ_TILE) {
            const int col = tile_n * OZ_TILE + tid;
            const int hc = p.h[m0 + col];
            s_cb[tid] = oz_pow2(max(p.exB[col], -900));
            s_ch[tid] = oz_pow2(-hc);
            s_h[tid] = hc;
        } else if (p.phase == 2 && tid < OZ_TILE + 6 * 32) {
            const int c = (tid - OZ_TILE) >> 5, l = tid & 31;
            for (int j = l; j < OZ_TILE; j += 32) s_wx[c * 128 + j] = Wx[(size_t)(m0 + tile_n * OZ_TILE + j) + (size_t)ld * c];
        }

        // ---- border jobs: the rows / columns in front of the 128-aligned block, in fp64, while the tensor core works.
        // inner jobs (operand rows of this tile row x border inner indices) feed the exponents and count towards the tile-row barrier
        const int nch = (m0 + OZF_YC - 1) / OZF_YC;
        const int nIB = 4 * nch, nOB = ((n + 31) / 32) * nch;
        const int oq = (Mt - 1 - tile_n) * Mt + tile_m;   // outer jobs start with the CTAs that have no inner job
        int my_inner = 0, my_outer = 0;
        for (int j = tile_n; j < nIB; j += Mt) ++my_inner;
        for (int j = oq; j < nOB; j += T) ++my_outer;
        const int my_jobs = my_inner + my_outer;
        // all of this CTA's jobs as one flat loop over their outputs (no barrier between jobs: a thread has several independent dots in flight)
        for (int t = tid; t < my_jobs * 32 * OZF_YC; t += OZF_EPI_THREADS) {
            const int jn = t / (32 * OZF_YC), tt = t - jn * (32 * OZF_YC);
            const bool inner = jn < my_inner;
            const int job = inner ? tile_n + jn * Mt : oq + (jn - my_inner) * T;
            const int xb = job / nch, ch = job - xb * nch;
            // 32 x's (lanes) by up to OZF_YC y's; inner jobs: x = operand row (block row), y = border inner index; outer jobs:
            // x = any column index, y = border row
            const int y0 = ch * OZF_YC, ny = min(OZF_YC, m0 - y0);
            const int x0 = inner ? m0 + tile_m * OZ_TILE + xb * 32 : xb * 32;
            const int x = x0 + (tt & 31), y = tt >> 5;
            if (x < n && y < ny) {
                const int fr = inner ? x : y0 + y, o = inner ? y0 + y : x;
                double v = oz_border_dot(p, fr, o);
                if (p.phase == 2) {   // D[fr, o] = Sigma'[o, fr]
                    double r6 = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) r6 = fma(Wx[(size_t)o + (size_t)ld * c], Fx[(size_t)fr + (size_t)ld * c], r6);
                    v += r6;
                    if (o == fr) v += Tstep * oz_process_noise(p, o);
                }
                p.Out[p.phase == 1 ? (size_t)fr + (size_t)ld * o : (size_t)o + (size_t)ld * fr] = v;
                if (inner) atomicMax(p.exOut + (fr - m0), oz_exponent(v) - p.h[o]);
            }
        }
        __threadfence();
        epi_bar();
        if (my_inner > 0 && tid == 0) atomicAdd(cnt + tile_m, my_inner);
        epi_bar();   // (the staged column factors are visible to every epilogue warp)
        OZ_STAMP(1);

        double acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.0;
        for (int b = 0; b < nbatch; ++b) {
            const int dmin = OZ_DIAGS_PER_BATCH * b, dmax = min(dmin + OZ_DIAGS_PER_BATCH - 1, S - 1);
            mbar_wait(acc_full, (uint32_t)(b & 1));
            tc_fence_after();
            OZ_STAMP(2 + 2 * b);
            for (int d = dmin; d <= dmax; ++d) {
                const double scale = oz_pow2(-7 * d);
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((d - dmin) * OZ_TILE + cq * 32), v);
                tmem_ld_wait();
                // int32 -> fp64 without the conversion unit: 2^52 + 2^31 + x has x + 2^31 in its low mantissa word
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = fma(__hiloint2double(0x43300000, (int)(v[j] ^ 0x80000000u)) - 4503601774854144.0, scale, acc[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
            OZ_STAMP(3 + 2 * b);
        }
        // ---- this tile in fp64: D[row, col], row = operand row (phase 1: row i of W; phase 2: column j of Sigma'), col = inner index
        const int r = q * 32 + lane;                       // tile-local row
        const int row = tile_m * OZ_TILE + r, grow = m0 + row;
        const int lc0 = cq * 32;                           // tile-local first column of this thread
        const int gcol0 = m0 + tile_n * OZ_TILE + lc0;     // its index in Sigma
        double* tr = reinterpret_cast<double*>(smem) + (size_t)warp * 32 * OZF_TR_LD;   // (phase 2) the pipeline stages are idle now
        {
            const double ra = oz_pow2(max(p.exA[row], -900) - 12);
            double fx[6] = {0, 0, 0, 0, 0, 0};
            double tp = 0.0;
            if (p.phase == 2) {
#pragma unroll
                for (int c = 0; c < 6; ++c) fx[c] = Fx[(size_t)grow + (size_t)ld * c];
                tp = Tstep * oz_process_noise(p, grow);
            }
            int emax = 0;   // largest high word (sign cleared) of the entries scaled by 2^-h: its exponent field is the row's maximum
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                double v = (acc[j] * ra) * s_cb[lc0 + j];
                if (p.phase == 2) {   // Sigma'[i = gcol, j = grow] += (T B_b R B_b^T)[i, j] + T P on the diagonal (VIOFilter.cpp:188-189)
                    double r6 = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) r6 = fma(s_wx[c * 128 + lc0 + j], fx[c], r6);
                    v += r6;
                    if (grow == gcol0 + j) v += tp;
                    tr[lane * OZF_TR_LD + j] = v;
                } else if (gcol0 + j < m0 + 2) {
                    // W's block is consumed as int8 slices only; in fp64 the second product's border dots read W's border columns (written
                    // by the border jobs) and — for a landmark row that ends the border — at most the block's first two columns
                    p.Out[(size_t)grow + (size_t)ld * (gcol0 + j)] = v;
                }
                const double vs = v * s_ch[lc0 + j];
                emax = max(emax, __double2hiint(vs) & 0x7fffffff);
                acc[j] = vs;
            }
            const int be = emax >> 20;
            atomicMax(p.exOut + row, be == 0 ? -2000 : be - 1022);
        }
        if (p.phase == 2) {   // Sigma'[gcol, grow] through shared memory so that the lanes of a warp are 32 consecutive gcol
            __syncwarp();
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr)
                p.Out[(size_t)(gcol0 + lane) + (size_t)ld * (m0 + tile_m * OZ_TILE + q * 32 + rr)] = tr[rr * OZF_TR_LD + lane];
        }
        // ---- tile-row barrier: every tile of this row of tiles and its inner border jobs have published their maxima
        __threadfence();
        OZ_STAMP(6);
        epi_bar();
        if (tid == 0) {
            atomicAdd(cnt + tile_m, 1);
            oz_wait_ge(cnt + tile_m, Mt + nIB, p.err);
        }
        epi_bar();
        OZ_STAMP(7);
        // ---- emission: this tile as int8 slices of the next product's operand (rows = operand rows, inner index = block-local column)
        {
            const int e = max(__ldcg(p.exOut + row), -900);
            const double up = oz_pow2(6 - e);
            const int sw = (r >> 2) & 1;
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                const int kk = tile_n * OZ_TILE + lc0 + 16 * c2;   // rotated inner index k' of the chunk's first column
                double x[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) x[c] = acc[c2 * 16 + c] * up;
                const size_t tile = ((size_t)tile_m * KB + (kk >> 5)) * S;
                const int in_tile = r * OZ_KBLOCK + ((((kk & 31) >> 4) ^ sw) << 4);
                oz_emit16<S>(x, reinterpret_cast<uint4*>(p.slOut + tile * OZ_SLICE_TILE_BYTES + in_tile));
            }
        }
    }
    OZ_STAMP(8);
    tc_fence_before();
    __syncthreads();
    if (warp == OZF_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if (p.stamps && threadIdx.x == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); p.stamps[(size_t)ticket * OZ_STAMPS + 15] = (long long)g; }
}

// F's landmark rows as int8 slices without reading the zeros: row g >= 11 of F = I + T A_b (EqFMatrices.cpp:292-312, VIOFilter.cpp:177-185)
// is zero outside the 11 base columns (-T B_i in 0..2, T A_iv in 8..10 are what k_feature_step writes there) and its landmark's own
// 3 x 3 block (I + T A_ii), so the split of the rows [m0, n) is fourteen bytes per row and slice at positions that depend on the row only;
// every other byte of the (zero-initialised) slice array stays zero.  ex[row] = the row's exponent (entries scaled by 2^(+h[column])).
__global__ void __launch_bounds__(128) k_oz_split_F_rows(const double* F, int ld, int n, int m0, int Mc, int KB, int S, const int* h, int8_t* slices, int* ex) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;   // block-local row
    if (row >= Mc) return;
    const int g = m0 + row;
    const int b0 = 11 + (g - 11) / 3 * 3;
    int cols[14] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, b0, b0 + 1, b0 + 2};
    double v[14];
    int e = -2000;
#pragma unroll
    for (int c = 0; c < 14; ++c) {
        v[c] = g < n ? F[(size_t)g + (size_t)ld * cols[c]] * oz_pow2(h[cols[c]]) : 0.0;
        e = max(e, oz_exponent(v[c]));
    }
    ex[row] = e;
    const double up = oz_pow2(6 - max(e, -900));
    const double magic = 6755399441055744.0;
    const int sw = (row >> 2) & 1;
#pragma unroll
    for (int c = 0; c < 14; ++c) {
        const int kk = cols[c] >= m0 ? cols[c] - m0 : Mc + cols[c];   // rotated inner index
        const size_t tile = ((size_t)(row / OZ_TILE) * KB + (kk >> 5)) * S;
        int8_t* dst = slices + tile * OZ_SLICE_TILE_BYTES + (row % OZ_TILE) * OZ_KBLOCK + ((((kk & 31) >> 4) ^ sw) << 4) + (kk & 15);
        double x = v[c] * up;
        for (int s = 0; s < S; ++s) {
            const double t = x + magic;
            const double d = t - magic;
            x = (x - d) * 128.0;
            dst[(size_t)s * OZ_SLICE_TILE_BYTES] = (int8_t)(__double2loint(t) & 0xff);
        }
    }
}

// C = [0, C0] (VIOFilter.cpp:272-273, EqFMatrices.cpp:332-338) holds one 2 x 3 block per landmark: row a (landmark a / 2) is zero outside
// columns 11 + 3 (a / 2) .. + 2, column c >= 11 is zero outside rows 2 i, 2 i + 1 (i = (c - 11) / 3).  Like F's rows, its rows and its
// columns are split from those entries alone — a few bytes per row and slice at positions that depend on the row index only, into
// slice arrays that are zero elsewhere — instead of a pass over 12.7 MB of mostly zeros (25 us -> 5 us, on the update's critical path).
//   rows:    operand row a in [0, m), inner index = column (rotated like every n-long inner index, scaled by 2^(hsign h[column]))
//   columns: operand row = column c - m0 in [0, Mc), inner index = row a (no rotation, no scale)
__global__ void __launch_bounds__(128) k_oz_split_C_rows(const double* C, int ldm, int m, int m0, int Mc, int KB, int S, const int* h, int hsign,
                                                         int8_t* slices, int* ex) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= m) return;
    const int c0 = 11 + 3 * (a >> 1);
    double v[3];
    int e = -2000;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        v[t] = C[(size_t)a + (size_t)ldm * (c0 + t)] * (h ? oz_pow2(hsign * h[c0 + t]) : 1.0);
        e = max(e, oz_exponent(v[t]));
    }
    ex[a] = e;
    const double up = oz_pow2(6 - max(e, -900));
    const double magic = 6755399441055744.0;
    const int sw = (a >> 2) & 1;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int col = c0 + t;
        const int kk = col >= m0 ? col - m0 : Mc + col;
        const size_t tile = ((size_t)(a / OZ_TILE) * KB + (kk >> 5)) * S;
        int8_t* dst = slices + tile * OZ_SLICE_TILE_BYTES + (a % OZ_TILE) * OZ_KBLOCK + ((((kk & 31) >> 4) ^ sw) << 4) + (kk & 15);
        double x = v[t] * up;
        for (int s = 0; s < S; ++s) {
            const double tt = x + magic;
            const double d = tt - magic;
            x = (x - d) * 128.0;
            dst[(size_t)s * OZ_SLICE_TILE_BYTES] = (int8_t)(__double2loint(tt) & 0xff);
        }
    }
}
__global__ void __launch_bounds__(128) k_oz_split_C_cols(const double* C, int ldm, int m0, int Mc, int KB, int S, int8_t* slices, int* ex) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;   // operand row = column m0 + r of C
    if (r >= Mc) return;
    const int c = m0 + r, a0 = 2 * ((c - 11) / 3);
    double v[2];
    int e = -2000;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        v[t] = C[(size_t)(a0 + t) + (size_t)ldm * c];
        e = max(e, oz_exponent(v[t]));
    }
    ex[r] = e;
    const double up = oz_pow2(6 - max(e, -900));
    const double magic = 6755399441055744.0;
    const int sw = (r >> 2) & 1;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int kk = a0 + t;
        const size_t tile = ((size_t)(r / OZ_TILE) * KB + (kk >> 5)) * S;
        int8_t* dst = slices + tile * OZ_SLICE_TILE_BYTES + (r % OZ_TILE) * OZ_KBLOCK + ((((kk & 31) >> 4) ^ sw) << 4) + (kk & 15);
        double x = v[t] * up;
        for (int s = 0; s < S; ++s) {
            const double tt = x + magic;
            const double d = tt - magic;
            x = (x - d) * 128.0;
            dst[(size_t)s * OZ_SLICE_TILE_BYTES] = (int8_t)(__double2loint(tt) & 0xff);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline int oz_round_up(int a, int b) { return (a + b - 1) / b * b; }

size_t oz_slices_bytes(int rows, int k, int S) { return (size_t)S * oz_round_up(rows, OZ_TILE) * oz_round_up(k, OZ_KBLOCK); }

cudaError_t oz_init_device() {
    cudaError_t e = cudaFuncSetAttribute(k_oz_gemm<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, OzGemmCfg<7>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_oz_gemm<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, OzGemmCfg<8>::SMEM_BYTES)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_oz_gemm<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, OzGemmCfg<9>::SMEM_BYTES)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_oz_riccati<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, OzFusedCfg<7>::SMEM_BYTES)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_oz_riccati<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, OzFusedCfg<8>::SMEM_BYTES)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_oz_riccati<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, OzFusedCfg<9>::SMEM_BYTES)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_oz_split, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SPLIT_SMEM);
}

cudaError_t oz_split(const double* X, long stride_r, long stride_k, int rows, int k, int S, OzOperand* op, int8_t* slices, int* ex,
                     cudaStream_t stream, const OzKScale* ks, bool ex_ready, int ex_margin, int k_rot) {
    if (S < 1 || S > OZ_MAX_SLICES || rows < 1 || k < 1 || k_rot < 0 || k_rot >= k || (k_rot > 0 && (k - k_rot) % 128 != 0)) return cudaErrorInvalidValue;
    op->slices = slices; op->ex = ex; op->rows = rows; op->k = k; op->S = S; op->ex_margin = ex_margin;
    op->rows_pad = oz_round_up(rows, OZ_TILE); op->k_pad = oz_round_up(k, OZ_KBLOCK);
    const int* h = ks ? ks->h : nullptr;
    const int hsign = ks ? ks->sign : 0;
    if (!ex_ready) {
        cudaError_t e = oz_reset_exponents(ex, op->rows_pad, stream);
        if (e != cudaSuccess) return e;
        k_oz_rowmax<<<dim3((k + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, stream>>>(X, stride_r, stride_k, rows, k, h, hsign, ex);
    }
    k_oz_split<<<dim3((op->k_pad + 127) / 128, (rows + 31) / 32), dim3(32, 8), OZ_SPLIT_SMEM, stream>>>(X, stride_r, stride_k, rows, k, h, hsign, ex, ex_margin,
                                                                                                   slices, op->rows_pad, op->k_pad, S, k_rot);
    return cudaGetLastError();
}
cudaError_t oz_reset_exponents(int* ex, int count, cudaStream_t stream) {   // a very negative exponent everywhere
    return cudaMemsetAsync(ex, 0xC0, (size_t)count * sizeof(int), stream);
}
cudaError_t oz_rowmax(const double* X, long stride_r, long stride_k, int rows, int k, int* ex, cudaStream_t stream, const OzKScale* ks) {
    if (rows < 1 || k < 1) return cudaSuccess;
    k_oz_rowmax<<<dim3((k + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, stream>>>(X, stride_r, stride_k, rows, k, ks ? ks->h : nullptr, ks ? ks->sign : 0, ex);
    return cudaGetLastError();
}
cudaError_t oz_diag_scale(const double* Sigma, int ld, int n, int* h, cudaStream_t stream) {
    k_oz_diag_scale<<<(n + 255) / 256, 256, 0, stream>>>(Sigma, ld, n, h);
    return cudaGetLastError();
}

bool oz_fused_supported(int S, int Mt) { return S >= 7 && S <= 9 && Mt >= 2 && 1 + Mt <= OZ_FUSED_SYNC_INTS; }
cudaError_t oz_riccati_fused(const OzFusedParams& p, int S, cudaStream_t stream, bool pdl) {
    if (!oz_fused_supported(S, p.Mt) || p.Mc != p.Mt * OZ_TILE || p.n != p.m0 + p.Mc || p.m0 < 1 || p.KB * OZ_KBLOCK < p.n) return cudaErrorInvalidValue;
    if ((long long)p.KB * OZ_KBLOCK * 64 * 64 * S >= (1LL << 31)) return cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(p.Mt * p.Mt); cfg.blockDim = dim3(OZF_THREADS); cfg.stream = stream;
    cfg.dynamicSmemBytes = S == 7 ? OzFusedCfg<7>::SMEM_BYTES : S == 8 ? OzFusedCfg<8>::SMEM_BYTES : OzFusedCfg<9>::SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return S == 7 ? cudaLaunchKernelEx(&cfg, k_oz_riccati<7>, p) : S == 8 ? cudaLaunchKernelEx(&cfg, k_oz_riccati<8>, p) : cudaLaunchKernelEx(&cfg, k_oz_riccati<9>, p);
}
cudaError_t oz_split_F_rows(const double* F, int ld, int n, int m0, int S, const int* h, int8_t* slices, int* ex, cudaStream_t stream) {
    const int Mc = n - m0, KB = (n + OZ_KBLOCK - 1) / OZ_KBLOCK;
    if (Mc < OZ_TILE || Mc % OZ_TILE != 0 || m0 < 11) return cudaErrorInvalidValue;
    k_oz_split_F_rows<<<(Mc + 127) / 128, 128, 0, stream>>>(F, ld, n, m0, Mc, KB, S, h, slices, ex);
    return cudaGetLastError();
}

static void oz_fill_operand(OzOperand* op, int8_t* slices, int* ex, int rows, int k, int S) {
    op->slices = slices; op->ex = ex; op->rows = rows; op->k = k; op->S = S; op->ex_margin = 0;
    op->rows_pad = oz_round_up(rows, OZ_TILE); op->k_pad = oz_round_up(k, OZ_KBLOCK);
}
cudaError_t oz_split_C_rows(const double* C, int ldm, int m, int n, int m0, int S, const OzKScale* ks, OzOperand* op, int8_t* slices, int* ex, cudaStream_t stream) {
    const int Mc = n - m0;
    if (m < 2 || (m & 1) || Mc % OZ_TILE != 0 || m0 < 11 || 11 + 3 * (m / 2) != n) return cudaErrorInvalidValue;
    oz_fill_operand(op, slices, ex, m, n, S);
    k_oz_split_C_rows<<<(m + 127) / 128, 128, 0, stream>>>(C, ldm, m, m0, Mc, op->k_pad / OZ_KBLOCK, S, ks ? ks->h : nullptr, ks ? ks->sign : 0, slices, ex);
    return cudaGetLastError();
}
cudaError_t oz_split_C_cols(const double* C, int ldm, int m, int n, int m0, int S, OzOperand* op, int8_t* slices, int* ex, cudaStream_t stream) {
    const int Mc = n - m0;
    if (m < 2 || (m & 1) || Mc % OZ_TILE != 0 || m0 < 11 || 11 + 3 * (m / 2) != n) return cudaErrorInvalidValue;
    oz_fill_operand(op, slices, ex, Mc, m, S);
    k_oz_split_C_cols<<<(Mc + 127) / 128, 128, 0, stream>>>(C, ldm, m0, Mc, op->k_pad / OZ_KBLOCK, S, slices, ex);
    return cudaGetLastError();
}

cudaError_t oz_gemm(const OzOperand& A, const OzOperand& B, int M, int N, double alpha, double beta, const double* Cin, int ldcin, double* D,
                    int ldd, cudaStream_t stream, const OzRiccatiEpilogue* ric, const OzExponentsOut* exo, int ksplit) {
    if (ksplit < 1 || ksplit > 2 || (ksplit == 2 && (Cin || ric || exo))) return cudaErrorInvalidValue;   // (a split product only adds into a zeroed D)
    if (A.k_pad != B.k_pad || A.S != B.S || M > A.rows_pad || N > B.rows_pad || M < 1 || N < 1) return cudaErrorInvalidValue;
    if ((long long)A.k_pad * 64 * 64 * A.S >= (1LL << 31)) return cudaErrorInvalidValue;   // int32 accumulation of one diagonal must be exact
    OzParams p;
    p.M = M; p.N = N; p.KB = A.k_pad / OZ_KBLOCK; p.S = A.S;
    p.exA = A.ex; p.exB = B.ex; p.marginA = A.ex_margin; p.marginB = B.ex_margin;
    p.alpha = alpha; p.beta = Cin ? beta : 0.0; p.Cin = Cin; p.ldcin = ldcin; p.D = D; p.ldd = ldd;
    if (ric) p.ric = *ric; else { memset(&p.ric, 0, sizeof p.ric); }
    if (exo) p.exo = *exo; else { memset(&p.exo, 0, sizeof p.exo); }
    const int Mt = (M + OZ_TILE - 1) / OZ_TILE, Nt = (N + OZ_TILE - 1) / OZ_TILE;
    p.ksplit = ksplit;
    const dim3 grid(Mt * Nt * ksplit);
    switch (A.S) {
        case 7: k_oz_gemm<7><<<grid, OZG_THREADS, OzGemmCfg<7>::SMEM_BYTES, stream>>>(A.slices, B.slices, p); break;
        case 8: k_oz_gemm<8><<<grid, OZG_THREADS, OzGemmCfg<8>::SMEM_BYTES, stream>>>(A.slices, B.slices, p); break;
        case 9: k_oz_gemm<9><<<grid, OZG_THREADS, OzGemmCfg<9>::SMEM_BYTES, stream>>>(A.slices, B.slices, p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace eqvio
