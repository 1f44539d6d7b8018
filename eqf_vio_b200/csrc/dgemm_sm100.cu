// dgemm_sm100.cu — see dgemm_sm100.cuh.  Hand-written sm_100a kernel:
//   * operands staged by TMA (cp.async.bulk.tensor.{2d,3d}, SASS UTMALDG) with SWIZZLE_128B into a
//     STAGES-deep shared-memory ring; full/empty mbarriers; one producer warp, N consumer warps;
//   * consumers read conflict-free 8-byte fragments and issue DMMA.8x8x4, accumulators in registers;
//   * fused epilogue: alpha/beta, optional diagonal add (the T*P term of the Riccati step,
//     reference eqf_vio/src/VIOFilter.cpp:162-167,188).
//
// Shared-memory tile layouts (BK = 16 doubles = 128 B, one swizzle row):
//   "MN-major" operand (A always; B when transB): TMA 3-D box {16 (m_in), 16 (k), BM/16 (m_out)}
//      byte(m,k) = (m>>4)*2048 + k*128 + ((((m&15)>>1) ^ (k&7)) << 4) + (m&1)*8
//   "K-major" operand (B when !transB): TMA 2-D box {16 (k), BN (n)}
//      byte(k,n) = n*128 + (((k>>1) ^ (n&7)) << 4) + (k&1)*8
// Within a 16-deep k tile the four k4 MMAs (j = 0..3) use k(t,j) with t = lane&3:
//      k = ((t1^j1)<<3) | (t1<<2) | (t0<<1) | (t0^j0)
// which makes every fragment LDS.64 hit 16 distinct 8-byte slots per half-warp for BOTH layouts.
#include "dgemm_sm100.cuh"

#include <cudaTypedefs.h>  // PFN_cuTensorMapEncodeTiled
#include <stdio.h>
#include <stdlib.h>

namespace eqvio {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// One poll of the barrier (no loop): used where the wait can be long and the warp should sleep between polls.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
struct KernelParams {
    int M, N, K;
    double* D;
    int ldd;
    const double* Cin;
    int ldcin;
    double alpha, beta;
    int add_diag;  // D[m][m] += T * Pd(class of m)
    int no_edge_skip;  // debug A/B switch: treat partial tiles like full ones
    int skip_m, skip_n;  // output box [0, skip_m) x [0, skip_n) is not written
    double T;
    const double* T_dev;  // non-null: T is read from device memory
    double Pd[5];
};

template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int MINB_>
struct TileCfg {
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_, MINB = MINB_;
    static constexpr int BK = 16;
    static constexpr int NWM = BM / WM, NWN = BN / WN;
    static constexpr int CONSUMER_WARPS = NWM * NWN;
    static constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
    static constexpr int A_BYTES = BM * BK * 8;
    static constexpr int B_BYTES = BN * BK * 8;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;  // + barriers + align slack
    static_assert(WM % 16 == 0 && WN % 8 == 0, "warp tile granularity");
    static_assert(BM % 16 == 0 && BN % 16 == 0, "CTA tile granularity");
};

__device__ __forceinline__ double process_diag(const KernelParams& p, int m) {
    // index map of Sigma (eqf_vio/src/VIOFilter.cpp:163-167)
    return m < 3 ? p.Pd[0] : m < 6 ? p.Pd[1] : m < 8 ? p.Pd[2] : m < 11 ? p.Pd[3] : p.Pd[4];
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The four k4 MMA groups of one 16-deep k tile for this warp's WM x WN sub-tile: fragments from the swizzled stage
// buffers (layout comment at the top of the file), accumulators in registers.
template <class Cfg, bool TRANSB>
__device__ __forceinline__ void mma_ktile(double (&acc)[Cfg::WM / 8][Cfg::WN / 8][2], const uint32_t sa, const uint32_t sb,
                                          const uint32_t mn_off0, const uint32_t km_off0, const int b_half) {
    constexpr int MB = Cfg::WM / 8, NB = Cfg::WN / 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int j1 = j >> 1, j0 = j & 1;
        // XOR constants move k by (j1<<3 | j0): k*128 bits [7,10], chunk bit j0 for MN-major;
        // chunk bit (j1<<2) and half bit j0 for K-major
        const uint32_t mn_x = (uint32_t)((((j1 << 3) | j0) << 7) | (j0 << 4));
        const uint32_t km_x = (uint32_t)((j1 << 6) | (j0 << 3));
        double af[MB], bf[NB];
#pragma unroll
        for (int i = 0; i < MB; ++i) {
            const uint32_t off = (mn_off0 ^ mn_x ^ (uint32_t)((i & 1) << 6)) + (uint32_t)((i >> 1) * 2048);
            af[i] = lds_f64(sa + off);
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if (TRANSB) {
                const int blk = i + b_half;
                const uint32_t off = (mn_off0 ^ mn_x ^ (uint32_t)((blk & 1) << 6)) + (uint32_t)((blk >> 1) * 2048);
                bf[i] = lds_f64(sb + off);
            } else {
                const uint32_t off = (km_off0 ^ km_x) + (uint32_t)(i * 1024);
                bf[i] = lds_f64(sb + off);
            }
        }
#pragma unroll
        for (int i = 0; i < MB; ++i)
#pragma unroll
            for (int jn = 0; jn < NB; ++jn) dmma884(acc[i][jn][0], acc[i][jn][1], af[i], bf[jn]);
    }
}

// Fused epilogue of one warp sub-tile: D = alpha acc + beta Cin (+ T P on the diagonal), bounds- and skip-box-checked.
// m_base / n_base: global row of this thread's first accumulator row, global column of its first accumulator column.
template <class Cfg>
__device__ __forceinline__ void store_tile(const double (&acc)[Cfg::WM / 8][Cfg::WN / 8][2], const KernelParams& p, const int m_base,
                                           const int n_base) {
    constexpr int MB = Cfg::WM / 8, NB = Cfg::WN / 8;
    const double Tstep = (p.add_diag && p.T_dev != nullptr) ? *p.T_dev : p.T;
#pragma unroll
    for (int jn = 0; jn < NB; ++jn) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int n = n_base + jn * 8 + c;
            if (n >= p.N) continue;
#pragma unroll
            for (int i = 0; i < MB; ++i) {
                const int m = m_base + i * 8;
                if (m >= p.M) continue;
                if (m < p.skip_m && n < p.skip_n) continue;
                double v = p.alpha * acc[i][jn][c];
                if (p.beta != 0.0) v += p.beta * p.Cin[(size_t)m + (size_t)p.ldcin * n];
                if (p.add_diag && m == n) v += Tstep * process_diag(p, m);
                p.D[(size_t)m + (size_t)p.ldd * n] = v;
            }
        }
    }
}

// One CTA tile of D = alpha A op(B) + beta Cin.  `wait_flag` (may be null): the producer does not touch A before
// *wait_flag >= wait_count (A's rows of this tile are being written by other CTAs of the same launch);
// `signal_flag` (may be null): incremented once this tile's output is globally visible.
// Split-K (sk.ks > 1): the CTA computes k-tiles [chunk KT / ks, (chunk + 1) KT / ks) of the tile, leaves its partial accumulators in
// the tile's workspace slots and counts itself in; the chunk that arrives last adds the ks partials in ascending chunk order (its own
// from registers, at its position: the same sum whoever is last, so results are bit-stable), runs the epilogue and signals.
struct SplitK {
    double* ws;      // [tile][chunk][BM * BN] partial accumulators, fragment order
    int* cnt;        // [tile] arrival counters, zero before the first use and left zero
    int ks, chunk, tile;
};

template <class Cfg, bool TRANSB>
__device__ __forceinline__ void gemm_tile(const CUtensorMap* tmAp, const CUtensorMap* tmBp, const KernelParams& p, const int tile_m,
                                          const int tile_n, const int* wait_flag, const int wait_count, int* signal_flag,
                                          const SplitK sk = SplitK{nullptr, nullptr, 1, 0, 0}) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, WM = Cfg::WM, WN = Cfg::WN, STAGES = Cfg::STAGES;
    constexpr int MB = WM / 8, NB = WN / 8;
    const CUtensorMap& tmA = *tmAp;
    const CUtensorMap& tmB = *tmBp;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-B alignment
    uint8_t* smem = smem_raw + (base - raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KTall = (p.K + 15) >> 4;
    const int kt_begin = sk.ks > 1 ? (int)((long long)KTall * sk.chunk / sk.ks) : 0;
    const int KT = (sk.ks > 1 ? (int)((long long)KTall * (sk.chunk + 1) / sk.ks) : KTall) - kt_begin;   // k-tiles of this CTA
    if ((tile_m + 1) * BM <= p.skip_m && (tile_n + 1) * BN <= p.skip_n) return;  // whole tile inside the skipped box

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], Cfg::CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == Cfg::CONSUMER_WARPS) {
        // ===== producer warp: one lane drives TMA =====
        if (lane == 0) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
            if (wait_flag) {
                while (ld_acquire_gpu(wait_flag) < wait_count) __nanosleep(64);
                // the rows were written with generic-proxy stores by other CTAs; TMA reads them through the async proxy
                asm volatile("fence.proxy.async.global;" ::: "memory");
            }
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % STAGES;
                const uint32_t ph = (kt / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
                uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
                uint8_t* sb = sa + Cfg::A_BYTES;
                tma_load_3d(sa, &tmA, 0, (kt_begin + kt) * 16, tile_m * (BM / 16), &full[s]);
                if (TRANSB)
                    tma_load_3d(sb, &tmB, 0, (kt_begin + kt) * 16, tile_n * (BN / 16), &full[s]);
                else
                    tma_load_2d(sb, &tmB, (kt_begin + kt) * 16, tile_n * BN, &full[s]);
            }
        }
        return;
    }

    // ===== consumer warps =====
    const int wm = warp % Cfg::NWM, wn = warp / Cfg::NWM;
    const int g = lane >> 2, t = lane & 3, t1 = t >> 1, t0 = t & 1;

    double acc[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // per-thread fragment offsets at j = 0 (see layout comment at the top of the file)
    const int k0 = (t1 << 3) | (t1 << 2) | (t0 << 1) | t0;
    // MN-major: m = 16*mo + mi, mi = g (+8 for odd 8-row blocks)
    const uint32_t mn_off0 = (uint32_t)(k0 * 128 + (((g >> 1) ^ (k0 & 7)) << 4) + (g & 1) * 8);
    // K-major: row n (n&7 == g), chunk (k>>1)^g
    const uint32_t km_off0 = (uint32_t)(g * 128 + ((((k0 >> 1) ^ g) & 7) << 4) + (k0 & 1) * 8);

    // valid 8-row / 8-column blocks of this warp's sub-tile (edge tiles only compute these)
    const int m_rem = p.M - (tile_m * BM + wm * WM), n_rem = p.N - (tile_n * BN + wn * WN);
    const int mbv = m_rem <= 0 ? 0 : (m_rem >= WM ? MB : (m_rem + 7) >> 3);
    const int nbv = n_rem <= 0 ? 0 : (n_rem >= WN ? NB : (n_rem + 7) >> 3);
    // warp-granular skipping only: a warp whose whole sub-tile lies outside the matrix just keeps the
    // barriers moving; per-block predication of mma.sync costs a convergence barrier per DMMA (measured: slower)
    const bool idle = ((mbv == 0) || (nbv == 0)) && !p.no_edge_skip;

    const uint32_t a_warp = (uint32_t)((wm * WM / 16) * 2048);
    const uint32_t b_warp = TRANSB ? (uint32_t)((wn * WN / 16) * 2048) : (uint32_t)(wn * WN * 128);
    // note: for TRANSB with WN % 16 == 8 the warp's first 8-column block may start mid-atom
    const int b_half = TRANSB ? ((wn * WN) & 8) >> 3 : 0;

    if (wait_flag) {
        // a second-phase tile may sit in its slot for a while before its rows are complete: the DMMA warps sleep between
        // polls instead of spinning next to the first phase's working warps
        while (!mbar_test(&full[0], 0)) __nanosleep(256);
    }
    for (int kt = 0; kt < KT; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&full[s], ph);
        const uint32_t sa = base + s * Cfg::STAGE_BYTES + a_warp;
        const uint32_t sb = base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES + b_warp;
        if (!idle) mma_ktile<Cfg, TRANSB>(acc, sa, sb, mn_off0, km_off0, b_half);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    // ===== epilogue: registers -> global =====
    if (sk.ks > 1) {
        constexpr int CT = Cfg::CONSUMER_WARPS * 32;
        __shared__ int s_last;
        double* slot = sk.ws + ((size_t)sk.tile * sk.ks + sk.chunk) * (BM * BN) + threadIdx.x;
#pragma unroll
        for (int a = 0; a < MB; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                __stcg(slot + (size_t)((a * NB + b) * 2) * CT, acc[a][b][0]);
                __stcg(slot + (size_t)((a * NB + b) * 2 + 1) * CT, acc[a][b][1]);
            }
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory");   // consumer warps only (the producer has left)
        if (threadIdx.x == 0) {
            const int last = atomicAdd(sk.cnt + sk.tile, 1) == sk.ks - 1;
            if (last) sk.cnt[sk.tile] = 0;      // re-armed for the next launch
            s_last = last;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory");
        if (!s_last) return;
        __threadfence();
        double sum[MB][NB][2];
#pragma unroll
        for (int a = 0; a < MB; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b) sum[a][b][0] = sum[a][b][1] = 0.0;
        for (int ch = 0; ch < sk.ks; ++ch) {
            const double* src = sk.ws + ((size_t)sk.tile * sk.ks + ch) * (BM * BN) + threadIdx.x;
            const bool own = ch == sk.chunk;
#pragma unroll
            for (int a = 0; a < MB; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    sum[a][b][0] += own ? acc[a][b][0] : __ldcg(src + (size_t)((a * NB + b) * 2) * CT);
                    sum[a][b][1] += own ? acc[a][b][1] : __ldcg(src + (size_t)((a * NB + b) * 2 + 1) * CT);
                }
        }
#pragma unroll
        for (int a = 0; a < MB; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b) { acc[a][b][0] = sum[a][b][0]; acc[a][b][1] = sum[a][b][1]; }
    }
    store_tile<Cfg>(acc, p, tile_m * BM + wm * WM + g, tile_n * BN + wn * WN + 2 * t);
    if (signal_flag) {
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(Cfg::CONSUMER_WARPS * 32) : "memory");   // consumer warps only (the producer has left)
        if (threadIdx.x == 0) atomicAdd(signal_flag, 1);
    }
}

struct SplitKArgs { double* ws; int* cnt; int ks; };

template <class Cfg, bool TRANSB>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
dgemm_dmma_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KernelParams p) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN;
    // 1-D grid, full tiles first and the partial edge tiles last: in an edge tile the warps whose
    // sub-tile is entirely outside the matrix do no MMAs, so it is cheap, and cheap work dispatched last
    // fills the tail of the last wave (n = 11 + 3N is never a multiple of the tile size).
    int tile_m, tile_n;
    {
        const int Tn = (p.N + BN - 1) / BN, Fm = p.M / BM, Fn = p.N / BN;
        const int id = blockIdx.x, interior = Fm * Fn;
        if (id < interior) { tile_m = id % Fm; tile_n = id / Fm; }
        else {
            const int e = id - interior;
            if (Tn > Fn && e < Fm) { tile_m = e; tile_n = Fn; }          // partial last tile column
            else { tile_m = Fm; tile_n = e - (Tn > Fn ? Fm : 0); }       // partial last tile row (incl. corner)
        }
    }
    gemm_tile<Cfg, TRANSB>(&tmA, &tmB, p, tile_m, tile_n, nullptr, 0, nullptr);
}

// One product with every tile cut into ks k-ranges (SplitK): for thin products (a few rows or columns, a long k loop) whose tile count
// cannot fill the GPU — the strips in front of the int8 core block (ozaki_sm100.cuh) — ks CTAs share a tile's k loop.
template <class Cfg, bool TRANSB>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
dgemm_splitk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KernelParams p, const SplitKArgs ska) {
    const int Tn = (p.N + Cfg::BN - 1) / Cfg::BN;
    const int id = blockIdx.x / ska.ks, chunk = blockIdx.x - id * ska.ks;
    const int tile_m = id / Tn, tile_n = id - tile_m * Tn;
    gemm_tile<Cfg, TRANSB>(&tmA, &tmB, p, tile_m, tile_n, nullptr, 0, nullptr, SplitK{ska.ws, ska.cnt, ska.ks, chunk, id});
}

// Two dependent GEMMs in one launch — the reference's left-to-right products (A B) C (eqf_vio/src/VIOFilter.cpp:188-189
// (F Sigma) F^T, :276 (C Sigma) C^T, :297 (K C) Sigma):
//   phase 1:  W = A1 B1                (p1, NN)   tiles in row-major order, each one counted in rows[tile_m]
//   phase 2:  D = W op(B2) (+ epilogue) (p2)       tile (i, j) starts once rows[i] == all tiles of row-block i of W
// Phase-2 CTAs fill the SM slots that the tail of phase 1 leaves empty (one drain per step instead of two) and have
// their barriers / descriptors set up by the time their rows are complete.  Logical CTA ids are tickets taken in
// start order, so a CTA only ever waits for CTAs that are already running; the CTA that finishes last resets the
// counters for the next launch.   sync[0] ticket, sync[1] finished CTAs, sync[8 + i] row-block counters.
template <class Cfg, bool TRANSB2>
__global__ void __launch_bounds__(Cfg::THREADS, (Cfg::STAGES == 4 ? 6 : Cfg::MINB))   // 6 CTAs / SM like the single-product kernel (64 registers)
dgemm_pair_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                          const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const KernelParams p1,
                          const KernelParams p2, int* sync, const SplitKArgs ska) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN;
    __shared__ int s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(&sync[0], 1);
    __syncthreads();
    const int ks = ska.ks;
    const int id = s_ticket / ks, chunk = s_ticket - id * ks;    // the ks chunks of a tile hold consecutive tickets
    const int Tm1 = (p1.M + BM - 1) / BM, Tn1 = (p1.N + BN - 1) / BN, T1 = Tm1 * Tn1;
    const int Tm2 = (p2.M + BM - 1) / BM, Tn2 = (p2.N + BN - 1) / BN, Fn2 = p2.N / BN;
    if (id < T1) {
        const int tile_m = id / Tn1, tile_n = id - tile_m * Tn1;
        gemm_tile<Cfg, false>(&tmA1, &tmB1, p1, tile_m, tile_n, nullptr, 0, &sync[8 + tile_m], SplitK{ska.ws, ska.cnt, ks, chunk, id});
    } else {
        // row-major over the full-width tile columns (rows of W become ready in that order), the partial last tile
        // column (cheap tiles) at the very end
        const int e = id - T1, body = Tm2 * Fn2;
        int tile_m, tile_n;
        if (e < body) { tile_m = e / Fn2; tile_n = e - tile_m * Fn2; }
        else { tile_m = e - body; tile_n = Fn2; }
        (void)Tn2;
        gemm_tile<Cfg, TRANSB2>(&tmA2, &tmB2, p2, tile_m, tile_n, &sync[8 + tile_m], Tn1, nullptr, SplitK{ska.ws, ska.cnt, ks, chunk, id});
    }
    if (threadIdx.x == 0) {   // a consumer thread: its tile is complete (and its wait / signal on the counters behind it)
        const int total = gridDim.x;
        if (atomicAdd(&sync[1], 1) == total - 1) {
            for (int i = 0; i < Tm1; ++i) sync[8 + i] = 0;
            sync[0] = 0;
            sync[1] = 0;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Stream-K form of the same two-product step, for the sizes where one product is a single partial wave of tiles
// (N = 256: 625 tiles of 32 x 32 on 148 SMs = 4.2 per SM, so some SMs carry 5 tiles and some 4, and launch, pipeline
// fill and drain are paid per product with nothing behind them to hide in).  G = 148 k persistent CTAs, all
// co-resident; the work of a product is its T tiles x KT k-tiles laid end to end ("units"), and CTA c takes the
// contiguous range [c U / G, (c + 1) U / G): every SM issues the same number of DMMAs.  A range is cut at tile
// boundaries; the piece that does not reach its tile's end is computed FIRST and posted (workspace slot + flag) for
// the CTA that finishes that tile, then come the whole tiles, last the piece that finishes the tile the range starts
// in, which adds the partials of the lower-numbered CTAs (already posted, or being computed without waiting on
// anyone: no cycle) in ascending-k order and then its own: a fixed order, so results are bit-stable run to run.  One grid-wide barrier separates the products (W must be complete), after which the
// producer warps fence the async proxy (W was written with generic stores, TMA reads it).  The producer warp runs
// ahead across tile and segment boundaries, so the pipeline never drains inside a product.
//   sync[0] barrier arrivals, sync[1] finished CTAs (the last one re-arms both), sync[8 + c] partial-ready flag of CTA c
//   ws: G slots of Cfg::BM * Cfg::BN doubles (fragment order: [register][consumer thread])
// ------------------------------------------------------------------------------------------------
struct StreamKSeg { int tile, k0, k1; };   // k-tiles [k0, k1) of one tile; k1 < KT: post a partial; k0 > 0 && k1 == KT: gather the partials

// The range [u0, u1) of CTA `c` cut at tile boundaries, in processing order: the piece that does not reach its tile's end
// (at most one: the last) FIRST — it is posted for the CTA that finishes that tile and never waits on anyone — then the
// whole tiles, then the piece that finishes the tile the range starts in (it gathers what lower-numbered CTAs posted).
struct StreamKPlan {
    long long u0, u1;
    int KT, t_first, t_last, count;
    bool open_end;          // the range ends inside tile t_last
    __device__ StreamKPlan(int c, int G, int T, int KT_) : KT(KT_) {
        const long long U = (long long)T * KT;
        u0 = U * c / G; u1 = U * (c + 1) / G;
        t_first = (int)(u0 / KT);
        t_last = (int)((u1 - 1) / KT);
        open_end = (u1 % KT) != 0;
        count = u1 > u0 ? t_last - t_first + 1 : 0;
    }
    __device__ StreamKSeg seg(int i) const {
        // natural order: tile t_first + j, j = 0 .. count - 1; processing order: [last if open_end], 1 .. , 0
        int j;
        if (open_end) j = (i == 0) ? count - 1 : (i < count - 1 ? i : 0);
        else j = (i < count - 1) ? i + 1 : 0;
        if (count == 1) j = 0;
        StreamKSeg s;
        s.tile = t_first + j;
        s.k0 = (j == 0) ? (int)(u0 % KT) : 0;
        s.k1 = (j == count - 1 && open_end) ? (int)(u1 % KT) : KT;
        return s;
    }
};

template <class Cfg>
__device__ __forceinline__ void streamk_tile_coords(const KernelParams& p, int tile, int& tile_m, int& tile_n) {
    const int Tn = (p.N + Cfg::BN - 1) / Cfg::BN;
    tile_m = tile / Tn;            // row-major: a row block of the first product's output completes early and in order
    tile_n = tile - tile_m * Tn;
}

template <class Cfg, bool TRANSB>
__device__ __forceinline__ void streamk_produce(const CUtensorMap* tmA, const CUtensorMap* tmB, const KernelParams& p, const int c, const int G,
                                                uint8_t* smem, uint64_t* full, uint64_t* empty, uint32_t& it) {
    constexpr int STAGES = Cfg::STAGES;
    const int KT = (p.K + 15) >> 4, T = ((p.M + Cfg::BM - 1) / Cfg::BM) * ((p.N + Cfg::BN - 1) / Cfg::BN);
    const StreamKPlan plan(c, G, T, KT);
    for (int i = 0; i < plan.count; ++i) {
        const StreamKSeg sg = plan.seg(i);
        int tile_m, tile_n;
        streamk_tile_coords<Cfg>(p, sg.tile, tile_m, tile_n);
        for (int kt = sg.k0; kt < sg.k1; ++kt, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
            uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + Cfg::A_BYTES;
            tma_load_3d(sa, tmA, 0, kt * 16, tile_m * (Cfg::BM / 16), &full[s]);
            if (TRANSB)
                tma_load_3d(sb, tmB, 0, kt * 16, tile_n * (Cfg::BN / 16), &full[s]);
            else
                tma_load_2d(sb, tmB, kt * 16, tile_n * Cfg::BN, &full[s]);
        }
    }
}

template <class Cfg, bool TRANSB>
__device__ __forceinline__ void streamk_consume(const KernelParams& p, const int c, const int G, const uint32_t base, uint64_t* full,
                                                uint64_t* empty, uint32_t& it, double* ws, int* flags) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, WM = Cfg::WM, WN = Cfg::WN, STAGES = Cfg::STAGES;
    constexpr int MB = WM / 8, NB = WN / 8, NREG = MB * NB * 2, CT = Cfg::CONSUMER_WARPS * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp % Cfg::NWM, wn = warp / Cfg::NWM;
    const int g = lane >> 2, t = lane & 3, t1 = t >> 1, t0 = t & 1;
    const int k0f = (t1 << 3) | (t1 << 2) | (t0 << 1) | t0;
    const uint32_t mn_off0 = (uint32_t)(k0f * 128 + (((g >> 1) ^ (k0f & 7)) << 4) + (g & 1) * 8);
    const uint32_t km_off0 = (uint32_t)(g * 128 + ((((k0f >> 1) ^ g) & 7) << 4) + (k0f & 1) * 8);
    const uint32_t a_warp = (uint32_t)((wm * WM / 16) * 2048);
    const uint32_t b_warp = TRANSB ? (uint32_t)((wn * WN / 16) * 2048) : (uint32_t)(wn * WN * 128);
    const int b_half = TRANSB ? ((wn * WN) & 8) >> 3 : 0;
    const int KT = (p.K + 15) >> 4, T = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
    const StreamKPlan plan(c, G, T, KT);
    const long long U = (long long)T * KT;
    for (int i = 0; i < plan.count; ++i) {
        const StreamKSeg sg = plan.seg(i);
        int tile_m, tile_n;
        streamk_tile_coords<Cfg>(p, sg.tile, tile_m, tile_n);
        const int m_rem = p.M - (tile_m * BM + wm * WM), n_rem = p.N - (tile_n * BN + wn * WN);
        const bool idle = (m_rem <= 0 || n_rem <= 0) && !p.no_edge_skip;
        double acc[MB][NB][2];
#pragma unroll
        for (int a = 0; a < MB; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        for (int kt = sg.k0; kt < sg.k1; ++kt, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&full[s], ph);
            const uint32_t sa = base + s * Cfg::STAGE_BYTES + a_warp;
            const uint32_t sb = base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES + b_warp;
            if (!idle) mma_ktile<Cfg, TRANSB>(acc, sa, sb, mn_off0, km_off0, b_half);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (sg.k1 < KT) {
            // this piece does not finish its tile: post the partial accumulators (fragment order, coalesced) for the CTA that does
            double* slot = ws + (size_t)c * (BM * BN) + threadIdx.x;
#pragma unroll
            for (int a = 0; a < MB; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    __stcg(slot + (size_t)((a * NB + b) * 2) * CT, acc[a][b][0]);
                    __stcg(slot + (size_t)((a * NB + b) * 2 + 1) * CT, acc[a][b][1]);
                }
            __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory");   // consumer warps only
            if (threadIdx.x == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flags + c), "r"(1) : "memory");
            continue;
        }
        if (sg.k0 > 0) {
            // this piece finishes a tile other CTAs started: add their partials in ascending-k order (the lowest-numbered CTA holds
            // the tile's first k-tiles), then this CTA's own — a fixed order, so the result is bit-stable
            const long long tile_u0 = (long long)sg.tile * KT;
            int cf = c - 1;
            while (cf > 0 && U * cf / G > tile_u0) --cf;
            double sum[MB][NB][2];
#pragma unroll
            for (int a = 0; a < MB; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) sum[a][b][0] = sum[a][b][1] = 0.0;
            for (int cc = cf; cc < c; ++cc) {
                if (threadIdx.x == 0) {
                    while (ld_acquire_gpu(flags + cc) == 0) __nanosleep(32);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory");
                const double* slot = ws + (size_t)cc * (BM * BN) + threadIdx.x;
#pragma unroll
                for (int a = 0; a < MB; ++a)
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        sum[a][b][0] += __ldcg(slot + (size_t)((a * NB + b) * 2) * CT);
                        sum[a][b][1] += __ldcg(slot + (size_t)((a * NB + b) * 2 + 1) * CT);
                    }
                asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory");   // every thread has read the slot
                if (threadIdx.x == 0) flags[cc] = 0;                      // re-armed for the next product / launch
            }
#pragma unroll
            for (int a = 0; a < MB; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) { acc[a][b][0] = sum[a][b][0] + acc[a][b][0]; acc[a][b][1] = sum[a][b][1] + acc[a][b][1]; }
        }
        store_tile<Cfg>(acc, p, tile_m * BM + wm * WM + g, tile_n * BN + wn * WN + 2 * t);
    }
    (void)NREG;
}

template <class Cfg, bool TRANSB2>
__global__ void __launch_bounds__(Cfg::THREADS, 6)
dgemm_streamk_pair_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                          const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const KernelParams p1,
                          const KernelParams p2, const int two_products, int* sync, double* ws) {
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x, G = gridDim.x;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], Cfg::CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t it = 0;   // k-tiles through the ring so far: producer and consumers walk the same sequence
    int* flags = sync + 8;
    if (warp == Cfg::CONSUMER_WARPS) {
        if (lane == 0) {
            tma_prefetch_desc(&tmA1);
            tma_prefetch_desc(&tmB1);
            streamk_produce<Cfg, false>(&tmA1, &tmB1, p1, c, G, smem, full, empty, it);
            if (two_products) {
                tma_prefetch_desc(&tmA2);
                tma_prefetch_desc(&tmB2);
                // grid-wide barrier: thread 0 of every CTA arrives once its first-product stores are visible
                while (ld_acquire_gpu(sync) < G) __nanosleep(64);
                asm volatile("fence.proxy.async.global;" ::: "memory");
                streamk_produce<Cfg, TRANSB2>(&tmA2, &tmB2, p2, c, G, smem, full, empty, it);
            }
        }
        return;
    }
    streamk_consume<Cfg, false>(p1, c, G, base, full, empty, it, ws, flags);
    if (two_products) {
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(Cfg::CONSUMER_WARPS * 32) : "memory");
        if (threadIdx.x == 0) atomicAdd(sync, 1);
        streamk_consume<Cfg, TRANSB2>(p2, c, G, base, full, empty, it, ws, flags);
    }
    if (threadIdx.x == 0) {
        if (atomicAdd(&sync[1], 1) == G - 1) {   // everyone is past the barrier: re-arm for the next launch
            sync[0] = 0;
            sync[1] = 0;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
using Cfg128x128 = TileCfg<128, 128, 64, 32, 4, 1>;
using Cfg128x64 = TileCfg<128, 64, 32, 32, 4, 2>;
using Cfg64x64 = TileCfg<64, 64, 32, 32, 4, 3>;
using Cfg32x32 = TileCfg<32, 32, 16, 16, 4, 4>;
using Cfg32x64 = TileCfg<32, 64, 16, 32, 4, 4>;
using Cfg64x32 = TileCfg<64, 32, 32, 16, 4, 4>;
using Cfg32x32s6 = TileCfg<32, 32, 16, 16, 6, 5>;
using Cfg48x48 = TileCfg<48, 48, 16, 48, 4, 4>;
// 64-deep in-place panel solves of the Schur chains: the tile spans the whole 64-wide side it shares with its neighbours (so
// a CTA reads exactly the rows / columns it later overwrites) and two stages keep it within one freed 32x32-GEMM slot (25 KB)
using Cfg32x64s2 = TileCfg<32, 64, 16, 32, 2, 6>;
using Cfg64x32s2 = TileCfg<64, 32, 32, 16, 2, 6>;

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

// 3-D map of an MN-major operand X (R x K, element (r,k) at X[r + k*ld]): dims {16, K, ceil(R/16)}.
static CUresult encode_mn_major(CUtensorMap* map, const double* X, int R, int K, int ld, int box_r) {
    cuuint64_t dims[3] = {16, (cuuint64_t)K, (cuuint64_t)((R + 15) / 16)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8, 128};
    cuuint32_t box[3] = {16, 16, (cuuint32_t)(box_r / 16)};
    cuuint32_t estr[3] = {1, 1, 1};
    return get_encode()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(X), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}
// 2-D map of a K-major operand B (K x N, element (k,n) at B[k + n*ld]): dims {K, N}.
static CUresult encode_k_major(CUtensorMap* map, const double* X, int K, int N, int ld, int box_n) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {16, (cuuint32_t)box_n};
    cuuint32_t estr[2] = {1, 1};
    return get_encode()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(X), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

static void fill_params(KernelParams& p, const GemmProblem& g) {
    p.M = g.M; p.N = g.N; p.K = g.K;
    p.D = g.D; p.ldd = g.ldd;
    p.Cin = g.epi.Cin; p.ldcin = g.epi.ldcin;
    p.alpha = g.epi.alpha;
    p.beta = g.epi.Cin ? g.epi.beta : 0.0;
    p.add_diag = g.epilogue == EPI_RICCATI;
    {
        static int noedge = -1;
        if (noedge < 0) { const char* e = getenv("EQVIO_GEMM_NOEDGE"); noedge = (e && e[0] == '1') ? 1 : 0; }
        p.no_edge_skip = noedge;
    }
    p.T = g.epi.T;
    p.skip_m = g.skip_m; p.skip_n = g.skip_n;
    p.T_dev = g.epi.T_dev;
    for (int i = 0; i < 5; ++i) p.Pd[i] = g.epi.Pd[i];
}

template <class Cfg>
static cudaError_t launch_cfg(const GemmProblem& g, cudaStream_t stream) {
    if (!get_encode()) return cudaErrorNotSupported;
    CUtensorMap tmA, tmB;
    if (encode_mn_major(&tmA, g.A, g.M, g.K, g.lda, Cfg::BM) != CUDA_SUCCESS) return cudaErrorInvalidValue;
    CUresult rb = g.transB ? encode_mn_major(&tmB, g.B, g.N, g.K, g.ldb, Cfg::BN)
                           : encode_k_major(&tmB, g.B, g.K, g.N, g.ldb, Cfg::BN);
    if (rb != CUDA_SUCCESS) return cudaErrorInvalidValue;
    KernelParams p;
    fill_params(p, g);
    dim3 grid(((g.M + Cfg::BM - 1) / Cfg::BM) * ((g.N + Cfg::BN - 1) / Cfg::BN));
    // the dynamic shared-memory opt-in is per device: dgemm_init_device() has set it for the current one
    if (g.transB)
        dgemm_dmma_tma_kernel<Cfg, true><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
    else
        dgemm_dmma_tma_kernel<Cfg, false><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
    return cudaGetLastError();
}

cudaError_t dgemm_splitk_launch(const GemmProblem& g, int ks, double* ws, int* cnt, cudaStream_t stream) {
    using Cfg = Cfg32x32;
    if (g.M <= 0 || g.N <= 0) return cudaSuccess;
    if (!get_encode()) return cudaErrorNotSupported;
    const int tiles = ((g.M + Cfg::BM - 1) / Cfg::BM) * ((g.N + Cfg::BN - 1) / Cfg::BN);
    if (ks < 2 || g.skip_m || g.skip_n || !ws || !cnt || (long)tiles * ks > DGEMM_SPLITK_MAX_SLOTS) return cudaErrorInvalidValue;
    CUtensorMap tmA, tmB;
    if (encode_mn_major(&tmA, g.A, g.M, g.K, g.lda, Cfg::BM) != CUDA_SUCCESS) return cudaErrorInvalidValue;
    const CUresult rb = g.transB ? encode_mn_major(&tmB, g.B, g.N, g.K, g.ldb, Cfg::BN) : encode_k_major(&tmB, g.B, g.K, g.N, g.ldb, Cfg::BN);
    if (rb != CUDA_SUCCESS) return cudaErrorInvalidValue;
    KernelParams p;
    fill_params(p, g);
    const SplitKArgs ska{ws, cnt, ks};
    if (g.transB)
        dgemm_splitk_kernel<Cfg, true><<<dim3(tiles * ks), Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p, ska);
    else
        dgemm_splitk_kernel<Cfg, false><<<dim3(tiles * ks), Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p, ska);
    return cudaGetLastError();
}

template <class Cfg, bool TB2>
static cudaError_t launch_pair_cfg(const GemmProblem& g1, const GemmProblem& g2, int* sync, cudaStream_t stream, int ks = 1, double* ws = nullptr) {
    if (!get_encode()) return cudaErrorNotSupported;
    CUtensorMap tmA1, tmB1, tmA2, tmB2;
    if (encode_mn_major(&tmA1, g1.A, g1.M, g1.K, g1.lda, Cfg::BM) != CUDA_SUCCESS) return cudaErrorInvalidValue;
    if (encode_k_major(&tmB1, g1.B, g1.K, g1.N, g1.ldb, Cfg::BN) != CUDA_SUCCESS) return cudaErrorInvalidValue;
    if (encode_mn_major(&tmA2, g2.A, g2.M, g2.K, g2.lda, Cfg::BM) != CUDA_SUCCESS) return cudaErrorInvalidValue;
    const CUresult rb = TB2 ? encode_mn_major(&tmB2, g2.B, g2.N, g2.K, g2.ldb, Cfg::BN) : encode_k_major(&tmB2, g2.B, g2.K, g2.N, g2.ldb, Cfg::BN);
    if (rb != CUDA_SUCCESS) return cudaErrorInvalidValue;
    KernelParams p1, p2;
    fill_params(p1, g1);
    fill_params(p2, g2);
    const int T1 = ((g1.M + Cfg::BM - 1) / Cfg::BM) * ((g1.N + Cfg::BN - 1) / Cfg::BN);
    const int T2 = ((g2.M + Cfg::BM - 1) / Cfg::BM) * ((g2.N + Cfg::BN - 1) / Cfg::BN);
    // split-K: the arrival counters of the (T1 + T2) tiles live behind the row-block counters of `sync`
    const SplitKArgs ska{ws, sync + 8 + DGEMM_PAIR_MAX_ROW_BLOCKS, ks};
    dgemm_pair_kernel<Cfg, TB2><<<dim3((T1 + T2) * ks), Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA1, tmB1, tmA2, tmB2, p1, p2, sync, ska);
    return cudaGetLastError();
}

// When the single launch pays (measured on B200, bench.py at N = 64 / 256 / 512): the first product must run for more
// than one wave of the 148 x 6 CTA slots, so that second-phase CTAs are dispatched into its tail and find their rows
// complete (N = 512: +5.4 % on the Riccati step); with a single partial wave (N = 256: 625 tiles) the second phase's CTAs
// would occupy the free slots at once, spin, and skew the SM load (-10 %).  Tiny problems whose two phases fit the GPU
// one CTA per SM save the second launch's set-up.
// Split-K factor of the pair launch: shapes whose products are a single partial wave of 32 x 32 tiles (N = 256 has 625 on 888 CTA slots)
// leave SMs with 4 or 5 whole tiles and too few warps to keep the DMMA pipe busy (77 % while resident against 90 % at n = 1547); cut
// into ks k-ranges the same work is more than a wave of finer-grained CTAs, 6 resident on every SM, and the ticketed pair launch pays
// again.  1 elsewhere.
int dgemm_pair_splitk(const GemmProblem& g1, const GemmProblem& g2) {
    const char* env = getenv("EQVIO_SPLITK");   // 0 = never, k >= 2 = that factor for every single-wave shape; default: by shape
    const int force = env ? atoi(env) : 1;
    const long t1 = (long)((g1.M + 31) / 32) * ((g1.N + 31) / 32), t2 = (long)((g2.M + 31) / 32) * ((g2.N + 31) / 32);
    const long tmin = t1 < t2 ? t1 : t2, tmax = t1 < t2 ? t2 : t1;
    // measured (profiles/r02_streamk_splitk.md, us per pair, two launches / ks = 2 / 3 / 4): n = 395 (169 tiles) 27.0 / 31.8 / 28.9 / 28.9,
    // n = 587 (361) 45.4 / 47.5 / 43.8 / 42.9, n = 779 (625) 84.6 / 77.4 / 75.8 / 76.9; from 888 tiles (a full wave) the plain pair wins
    if (force == 0 || tmin < 324 || t1 >= 888 || g1.skip_m || g2.skip_m || g1.skip_n || g2.skip_n) return 1;
    int ks = force >= 2 ? force : (tmin < 500 ? 4 : 3);
    const int kt = ((g1.K < g2.K ? g1.K : g2.K) + 15) / 16;
    while (ks > 1 && (kt / ks < 8 || tmax * ks > DGEMM_SPLITK_MAX_SLOTS / 2)) --ks;
    return ks < 1 ? 1 : ks > 4 ? 4 : ks;
}

bool dgemm_pair_pays(const GemmProblem& g1, const GemmProblem& g2) {
    const long t1 = (long)((g1.M + 31) / 32) * ((g1.N + 31) / 32), t2 = (long)((g2.M + 31) / 32) * ((g2.N + 31) / 32);
    if ((g1.M + 31) / 32 > DGEMM_PAIR_MAX_ROW_BLOCKS) return false;
    static int force = -1;   // EQVIO_PAIR_FORCE=1: pair whatever the shape (A/B measurements)
    if (force < 0) { const char* e = getenv("EQVIO_PAIR_FORCE"); force = (e && e[0] == '1') ? 1 : 0; }
    if (force) return true;
    return t1 + t2 <= 148 || t1 >= 888;   // n = 971 (961 tiles): 132 us as a pair against 147 us as two launches
}

// Per-device opt-in to more than 48 KB of dynamic shared memory for every instantiation this file can launch.
template <class Cfg>
static cudaError_t opt_in_cfg() {
    cudaError_t e = cudaFuncSetAttribute(dgemm_dmma_tma_kernel<Cfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(dgemm_dmma_tma_kernel<Cfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
}
template <class Cfg>
static cudaError_t opt_in_pair_cfg() {
    cudaError_t e = cudaFuncSetAttribute(dgemm_pair_kernel<Cfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(dgemm_pair_kernel<Cfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
}
cudaError_t dgemm_init_device() {
    cudaError_t e;
#define EQVIO_OPT_IN(call) if ((e = (call)) != cudaSuccess) return e
    EQVIO_OPT_IN(opt_in_cfg<Cfg128x128>()); EQVIO_OPT_IN(opt_in_cfg<Cfg128x64>()); EQVIO_OPT_IN(opt_in_cfg<Cfg64x64>());
    EQVIO_OPT_IN(opt_in_cfg<Cfg32x32>()); EQVIO_OPT_IN(opt_in_cfg<Cfg32x64>()); EQVIO_OPT_IN(opt_in_cfg<Cfg64x32>());
    EQVIO_OPT_IN(opt_in_cfg<Cfg32x32s6>()); EQVIO_OPT_IN(opt_in_cfg<Cfg48x48>()); EQVIO_OPT_IN(opt_in_cfg<Cfg32x64s2>());
    EQVIO_OPT_IN(opt_in_cfg<Cfg64x32s2>());
    EQVIO_OPT_IN(opt_in_pair_cfg<Cfg32x32>()); EQVIO_OPT_IN(opt_in_pair_cfg<Cfg32x32s6>());
    EQVIO_OPT_IN(cudaFuncSetAttribute(dgemm_splitk_kernel<Cfg32x32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg32x32::SMEM_BYTES));
    EQVIO_OPT_IN(cudaFuncSetAttribute(dgemm_splitk_kernel<Cfg32x32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg32x32::SMEM_BYTES));
    EQVIO_OPT_IN(cudaFuncSetAttribute(dgemm_streamk_pair_kernel<Cfg32x32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg32x32::SMEM_BYTES));
    EQVIO_OPT_IN(cudaFuncSetAttribute(dgemm_streamk_pair_kernel<Cfg32x32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg32x32::SMEM_BYTES));
#undef EQVIO_OPT_IN
    return cudaSuccess;
}

cudaError_t dgemm_pair_launch(const GemmProblem& g1, const GemmProblem& g2, int* sync, cudaStream_t stream) {
    if (g1.M <= 0 || g1.N <= 0 || g2.N <= 0) return cudaSuccess;
    if (g1.transB || g1.D != g2.A || g1.M != g2.M || g1.skip_m || g2.skip_m || (g1.M + 31) / 32 > DGEMM_PAIR_MAX_ROW_BLOCKS)
        return cudaErrorInvalidValue;
    // the second product's tiles are stored while first-product tiles of later row blocks may still be reading their
    // operands: its output must not alias anything the first product reads
    if (g2.D == g1.B || g2.D == g1.A) return cudaErrorInvalidValue;
    const long tiles32 = (long)((g1.M + 31) / 32) * ((g1.N + 31) / 32);
    if (tiles32 <= 148)
        return g2.transB ? launch_pair_cfg<Cfg32x32s6, true>(g1, g2, sync, stream) : launch_pair_cfg<Cfg32x32s6, false>(g1, g2, sync, stream);
    return g2.transB ? launch_pair_cfg<Cfg32x32, true>(g1, g2, sync, stream) : launch_pair_cfg<Cfg32x32, false>(g1, g2, sync, stream);
}

cudaError_t dgemm_pair_splitk_launch(const GemmProblem& g1, const GemmProblem& g2, int ks, int* sync, double* ws, cudaStream_t stream) {
    if (g1.M <= 0 || g1.N <= 0 || g2.N <= 0) return cudaSuccess;
    if (g1.transB || g1.D != g2.A || g1.M != g2.M || g1.skip_m || g2.skip_m || (g1.M + 31) / 32 > DGEMM_PAIR_MAX_ROW_BLOCKS || g2.D == g1.B || g2.D == g1.A)
        return cudaErrorInvalidValue;
    const long t1 = (long)((g1.M + 31) / 32) * ((g1.N + 31) / 32), t2 = (long)((g2.M + 31) / 32) * ((g2.N + 31) / 32);
    if (ks < 2 || (t1 + t2) * ks > DGEMM_SPLITK_MAX_SLOTS || !ws) return cudaErrorInvalidValue;
    return g2.transB ? launch_pair_cfg<Cfg32x32, true>(g1, g2, sync, stream, ks, ws) : launch_pair_cfg<Cfg32x32, false>(g1, g2, sync, stream, ks, ws);
}

// ---- stream-K launch -------------------------------------------------------------------------------------------
static const int STREAMK_MAX_PER_SM = 6;
int dgemm_streamk_ctas(int tiles) {
    // k persistent CTAs on every SM, all co-resident (6 fit: 64 registers x 160 threads, 34 KB of shared memory each)
    static int per_sm = -1;   // EQVIO_STREAMK_PER_SM: experiments
    if (per_sm < 0) { const char* e = getenv("EQVIO_STREAMK_PER_SM"); per_sm = e ? atoi(e) : 6; if (per_sm < 1 || per_sm > STREAMK_MAX_PER_SM) per_sm = 6; }
    return tiles >= 148 ? 148 * per_sm : 0;
}
size_t dgemm_streamk_ws_doubles() { return (size_t)148 * STREAMK_MAX_PER_SM * 32 * 32; }

template <bool TB2>
static cudaError_t launch_streamk(const GemmProblem& g1, const GemmProblem* g2, int* sync, double* ws, cudaStream_t stream) {
    using Cfg = Cfg32x32;
    if (!get_encode()) return cudaErrorNotSupported;
    CUtensorMap tmA1, tmB1, tmA2, tmB2;
    if (encode_mn_major(&tmA1, g1.A, g1.M, g1.K, g1.lda, Cfg::BM) != CUDA_SUCCESS) return cudaErrorInvalidValue;
    if (encode_k_major(&tmB1, g1.B, g1.K, g1.N, g1.ldb, Cfg::BN) != CUDA_SUCCESS) return cudaErrorInvalidValue;
    KernelParams p1, p2;
    fill_params(p1, g1);
    p2 = p1;
    tmA2 = tmA1; tmB2 = tmB1;
    if (g2) {
        if (encode_mn_major(&tmA2, g2->A, g2->M, g2->K, g2->lda, Cfg::BM) != CUDA_SUCCESS) return cudaErrorInvalidValue;
        const CUresult rb = TB2 ? encode_mn_major(&tmB2, g2->B, g2->N, g2->K, g2->ldb, Cfg::BN) : encode_k_major(&tmB2, g2->B, g2->K, g2->N, g2->ldb, Cfg::BN);
        if (rb != CUDA_SUCCESS) return cudaErrorInvalidValue;
        fill_params(p2, *g2);
    }
    const int T1 = ((g1.M + Cfg::BM - 1) / Cfg::BM) * ((g1.N + Cfg::BN - 1) / Cfg::BN);
    const int T2 = g2 ? ((g2->M + Cfg::BM - 1) / Cfg::BM) * ((g2->N + Cfg::BN - 1) / Cfg::BN) : T1;
    const int G = dgemm_streamk_ctas(T1 < T2 ? T1 : T2);
    if (G == 0) return cudaErrorInvalidValue;
    dgemm_streamk_pair_kernel<Cfg, TB2><<<dim3(G), Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA1, tmB1, tmA2, tmB2, p1, p2, g2 ? 1 : 0, sync, ws);
    return cudaGetLastError();
}

// Whether the stream-K form is the better choice for a two-product step of these shapes: a single partial wave of
// 32 x 32 tiles (more than one tile per SM, fewer than the 1110 from which the ticketed pair launch pays).
bool dgemm_streamk_pays(const GemmProblem& g1, const GemmProblem& g2) {
    // EQVIO_STREAMK=1: whenever legal.  Default: never — measured on B200 (profiles/r02_streamk_splitk.md) the persistent form
    // issues DMMAs ~10 % slower than hardware-dispatched CTAs at every size, which eats what the even split gains.
    const char* e = getenv("EQVIO_STREAMK");   // read per call (tests switch it inside one process)
    const int mode = e ? atoi(e) : 0;
    const long t1 = (long)((g1.M + 31) / 32) * ((g1.N + 31) / 32), t2 = (long)((g2.M + 31) / 32) * ((g2.N + 31) / 32);
    const long tmin = t1 < t2 ? t1 : t2;
    if (mode == 0 || tmin < 148) return false;
    if (g1.skip_m || g2.skip_m || g1.skip_n || g2.skip_n) return false;
    if (mode == 1) return true;
    return t1 < 1110;
}

cudaError_t dgemm_streamk_pair_launch(const GemmProblem& g1, const GemmProblem& g2, int* sync, double* ws, cudaStream_t stream) {
    if (g1.M <= 0 || g1.N <= 0 || g2.N <= 0) return cudaSuccess;
    if (g1.transB || g1.D != g2.A || g1.M != g2.M || g2.D == g1.A || g2.D == g1.B) return cudaErrorInvalidValue;
    return g2.transB ? launch_streamk<true>(g1, &g2, sync, ws, stream) : launch_streamk<false>(g1, &g2, sync, ws, stream);
}

int dgemm_num_configs() { return 10; }
const char* dgemm_config_name(int cfg) {
    switch (cfg) {
        case 0: return "128x128x16 (8 DMMA warps 64x32, 4 stages)";
        case 1: return "128x64x16 (8 DMMA warps 32x32, 4 stages)";
        case 2: return "64x64x16 (4 DMMA warps 32x32, 4 stages)";
        case 3: return "32x32x16 (4 DMMA warps 16x16, 4 stages)";
        case 4: return "32x64x16 (4 DMMA warps 16x32, 4 stages)";
        case 5: return "64x32x16 (4 DMMA warps 32x16, 4 stages)";
        case 6: return "32x32x16 (4 DMMA warps 16x16, 6 stages, 5 CTA/SM)";
        case 7: return "48x48x16 (3 DMMA warps 16x48, 4 stages)";
        case 8: return "32x64x16 (4 DMMA warps 16x32, 2 stages, 6 CTA/SM: in-place column-panel solves)";
        case 9: return "64x32x16 (4 DMMA warps 32x16, 2 stages, 6 CTA/SM: in-place row-panel solves)";
    }
    return "?";
}

// Tile-shape choice, from measurement on B200 (profiles/r01_gemm_tiles.md): an SM retires 64 fp64 FMA/clk
// however many CTAs share it and fp64 needs so little operand bandwidth that the smallest tile wins at
// every filter size (n = 203 ... 3083) — many small CTAs per SM balance the 148 SMs best (n = 11 + 3N is
// never a multiple of a big tile) and the DMMA pipe stays ~87% busy either way.  Larger tiles stay
// selectable (EQVIO_GEMM_CONFIG / force_config) for experiments.
int dgemm_pick_config(int M, int N, int K) {
    (void)K;
    const long tiles32 = (long)((M + 31) / 32) * ((N + 31) / 32);
    return tiles32 <= 148 ? 6 : 3;  // tiny problems: deeper pipeline, 5 CTAs/SM (latency-bound)
}

cudaError_t dgemm_launch(const GemmProblem& g, cudaStream_t stream, int force_config) {
    if (g.M <= 0 || g.N <= 0) return cudaSuccess;
    int cfg = force_config >= 0 ? force_config : dgemm_pick_config(g.M, g.N, g.K);
    switch (cfg) {
        case 0: return launch_cfg<Cfg128x128>(g, stream);
        case 1: return launch_cfg<Cfg128x64>(g, stream);
        case 2: return launch_cfg<Cfg64x64>(g, stream);
        case 4: return launch_cfg<Cfg32x64>(g, stream);
        case 5: return launch_cfg<Cfg64x32>(g, stream);
        case 6: return launch_cfg<Cfg32x32s6>(g, stream);
        case 7: return launch_cfg<Cfg48x48>(g, stream);
        case 8: return launch_cfg<Cfg32x64s2>(g, stream);
        case 9: return launch_cfg<Cfg64x32s2>(g, stream);
        default: return launch_cfg<Cfg32x32>(g, stream);
    }
}

}  // namespace eqvio
