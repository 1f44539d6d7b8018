// fp64_latency.cu — dependent-issue latencies (cycles) of the instructions on the per-pivot critical path of the
// diagonal-block LU kernel (k_getrf_diag_inv): DFMA, DMUL, F2F, MUFU.RCP, MUFU.RCP64H, SHFL, LDS, named barriers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu ; run on the B200.
#include <cstdio>
#include <cuda_runtime.h>

#define N_IT 512

template <int WHICH>
__global__ void lat_kernel(double* out, long long* cycles, double seed, int nthreads_bar) {
    __shared__ double sm[64];
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001, z = 0.5;
    float xf = (float)seed;
    sm[threadIdx.x & 63] = x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT / 8; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (WHICH == 0) x = fma(x, y, z);                                  // DFMA
            if (WHICH == 1) x = x * y;                                         // DMUL
            if (WHICH == 2) x = x + z;                                         // DADD
            if (WHICH == 3) { float f = (float)x; x = (double)f + 0.0; }       // F2F.F32.F64 + F2F.F64.F32 (+DADD)
            if (WHICH == 4) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(xf)); }   // MUFU.RCP
            if (WHICH == 5) { asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x)); }    // MUFU.RCP64H
            if (WHICH == 6) { int lo = __double2loint(x), hi = __double2hiint(x);
                              lo = __shfl_sync(0xffffffffu, lo, (j * 7 + 3) & 31); hi = __shfl_sync(0xffffffffu, hi, (j * 7 + 3) & 31);
                              x = __hiloint2double(hi, lo); }                  // 64-bit shuffle (2 x SHFL.IDX)
            if (WHICH == 7) { x = sm[(__double2loint(x) + j) & 63]; }          // dependent LDS.64
            if (WHICH == 8) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads_bar) : "memory"); }   // named barrier
            if (WHICH == 9) { sm[threadIdx.x & 63] = x; asm volatile("bar.sync 1, %0;" ::"r"(nthreads_bar) : "memory"); x = sm[(threadIdx.x + 1) & 63]; }  // STS + bar + LDS
            if (WHICH == 10) x = 1.0 / x;                                      // IEEE double division
            if (WHICH == 11) { x = fma(x, y, z); x = x * y; }                  // DFMA + DMUL
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    out[threadIdx.x] = x + xf;
}

template <int WHICH>
static void run(const char* name, int threads, double* out, long long* cyc) {
    lat_kernel<WHICH><<<1, threads>>>(out, cyc, 1.25, threads);
    cudaDeviceSynchronize();
    lat_kernel<WHICH><<<1, threads>>>(out, cyc, 1.25, threads);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-42s threads=%3d  %7.1f cycles per op\n", name, threads, (double)c / N_IT);
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    run<0>("DFMA dependent", 32, out, cyc);
    run<0>("DFMA dependent (4 warps)", 128, out, cyc);
    run<1>("DMUL dependent", 32, out, cyc);
    run<2>("DADD dependent", 32, out, cyc);
    run<11>("DFMA+DMUL dependent pair", 32, out, cyc);
    run<3>("F2F f64->f32->f64 (+DADD)", 32, out, cyc);
    run<4>("MUFU.RCP f32", 32, out, cyc);
    run<5>("MUFU.RCP64H", 32, out, cyc);
    run<6>("64-bit SHFL.IDX pair", 32, out, cyc);
    run<7>("LDS.64 dependent", 32, out, cyc);
    run<8>("bar.sync 1, 128", 128, out, cyc);
    run<8>("bar.sync 1, 512", 512, out, cyc);
    run<8>("bar.sync 1, 32", 32, out, cyc);
    run<9>("STS + bar.sync(128) + LDS", 128, out, cyc);
    run<9>("STS + bar.sync(32) + LDS", 32, out, cyc);
    run<10>("IEEE 1.0/x double", 32, out, cyc);
    return 0;
}
