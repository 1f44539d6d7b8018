// Microbenchmark (not product code): measures (1) raw DMMA.8x8x4 and DFMA issue throughput on
// the B200 and (2) cuBLAS DGEMM TFLOP/s at the filter's sizes.  The cuBLAS number is the
// "fp64 tensor peak" denominator SURVEY.md §8(d) asks the builder to measure.
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(1024) dmma_chain(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) dfma_chain(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(a, c[i], b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA and DFMA interleaved: are they separate pipes?
template <int NACC>
__global__ void __launch_bounds__(512) mixed_chain(double* out, const double* in, int iters, int nfma, long long* clk) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c[NACC][2];
  double d[8];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
#pragma unroll
  for (int i = 0; i < 8; i++) d[i] = i;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
      if (i < nfma) d[i & 7] = fma(a, d[i & 7], b);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 8; i++) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

// distinct A/B operand registers per DMMA (register-bank / reuse effects)
template <int NA, int NB>
__global__ void __launch_bounds__(256) dmma_tile(double* out, const double* in, int iters, long long* clk) {
  double a[NA], b[NB];
#pragma unroll
  for (int i = 0; i < NA; i++) a[i] = in[(threadIdx.x + i) & 63];
#pragma unroll
  for (int i = 0; i < NB; i++) b[i] = in[(threadIdx.x + 7 * i) & 63];
  double c[NA][NB][2];
#pragma unroll
  for (int i = 0; i < NA; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) { c[i][j][0] = 0; c[i][j][1] = 0; }
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NA; i++)
#pragma unroll
      for (int j = 0; j < NB; j++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NA; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; i++) f();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main(int argc, char** argv) {
  bool only_cublas = argc > 1 && argv[1][0] == 'c';
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  int sms = p.multiProcessorCount;
  double *in, *out;
  CK(cudaMalloc(&in, 1024)); CK(cudaMemset(in, 0, 1024));
  CK(cudaMalloc(&out, sizeof(double) * 1024 * sms * 8));
  int iters = 4096;
  long long* dclk; CK(cudaMalloc(&dclk, 8)); long long hclk;
  if (!only_cublas) {
  for (int warps : {4, 8, 16}) {
    for (int nfma : {0, 4, 8, 16}) {
      float ms = time_ms([&] { mixed_chain<16><<<sms, warps * 32>>>(out, in, iters, nfma, dclk); }, 5);
      CK(cudaMemcpy(&hclk, dclk, 8, cudaMemcpyDeviceToHost));
      double fma_mma = double(sms) * warps * iters * 16 * 256, fma_v = double(sms) * warps * iters * nfma * 32;
      printf("mixed 16 DMMA + %2d DFMA warps/SM=%2d : %.2f TFLOP/s total (%.2f mma + %.2f fma)  clk/iter=%.1f  SMclk=%.0f MHz\n", nfma, warps,
             2 * (fma_mma + fma_v) / ms / 1e9, 2 * fma_mma / ms / 1e9, 2 * fma_v / ms / 1e9, double(hclk) / iters, hclk / (ms * 1e3));
    }
  }
  for (int warps : {4, 8}) {
    float ms = time_ms([&] { dmma_tile<8, 4><<<sms, warps * 32>>>(out, in, iters / 2, dclk); }, 5);
    CK(cudaMemcpy(&hclk, dclk, 8, cudaMemcpyDeviceToHost));
    double fma = double(sms) * warps * (iters / 2) * 32 * 256;
    printf("dmma_tile<8,4> warps/SM=%2d : %.2f TFLOP/s  clk/DMMA/warp=%.2f SMclk=%.0f MHz -> %.1f FMA/clk/SM\n", warps, 2 * fma / ms / 1e9,
           double(hclk) / (iters / 2) / 32, hclk / (ms * 1e3), fma / sms / (double(hclk) * 1.0));
    ms = time_ms([&] { dmma_tile<4, 4><<<sms, warps * 32>>>(out, in, iters, dclk); }, 5);
    CK(cudaMemcpy(&hclk, dclk, 8, cudaMemcpyDeviceToHost));
    fma = double(sms) * warps * iters * 16 * 256;
    printf("dmma_tile<4,4> warps/SM=%2d : %.2f TFLOP/s  clk/DMMA/warp=%.2f SMclk=%.0f MHz -> %.1f FMA/clk/SM\n", warps, 2 * fma / ms / 1e9,
           double(hclk) / iters / 16, hclk / (ms * 1e3), fma / sms / (double(hclk) * 1.0));
  }
  if (!only_cublas)
  for (int warps : {4, 8, 16, 32}) {
    {
      float ms = time_ms([&] { dmma_chain<16><<<sms, warps * 32>>>(out, in, iters); }, 5);
      double fma = double(sms) * warps * iters * 16 * 256;
      printf("DMMA.8x8x4 x16acc  warps/SM=%2d : %.2f TFLOP/s  (%.1f FMA/clk/SM @1.965GHz)\n", warps, 2 * fma / ms / 1e9,
             fma / (ms * 1e-3) / sms / 1.965e9);
    }
    {
      float ms = time_ms([&] { dmma_chain<32><<<sms, warps * 32>>>(out, in, iters); }, 5);
      double fma = double(sms) * warps * iters * 32 * 256;
      printf("DMMA.8x8x4 x32acc  warps/SM=%2d : %.2f TFLOP/s\n", warps, 2 * fma / ms / 1e9);
    }
    {
      float ms = time_ms([&] { dfma_chain<16><<<sms, warps * 32>>>(out, in, iters * 8); }, 5);
      double fma = double(sms) * warps * iters * 8 * 16 * 32;
      printf("DFMA       x16acc  warps/SM=%2d : %.2f TFLOP/s\n", warps, 2 * fma / ms / 1e9);
    }
  }
  }
  if (!only_cublas)
  // latency of a dependent DMMA chain: 1 warp, 1 accumulator
  {
    float ms = time_ms([&] { dmma_chain<1><<<1, 32>>>(out, in, 1 << 16); }, 3);
    printf("dependent DMMA latency: %.1f ns\n", ms * 1e6 / (1 << 16));
  }

  cublasHandle_t h; cublasCreate(&h);
  size_t nmax = 8192;
  double *A, *B, *C;
  CK(cudaMalloc(&A, nmax * nmax * 8)); CK(cudaMalloc(&B, nmax * nmax * 8)); CK(cudaMalloc(&C, nmax * nmax * 8));
  std::vector<double> hbuf(nmax * nmax);
  for (auto& v : hbuf) v = rand() / double(RAND_MAX) - 0.5;
  CK(cudaMemcpy(A, hbuf.data(), nmax * nmax * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(B, hbuf.data(), nmax * nmax * 8, cudaMemcpyHostToDevice));
  double one = 1, zero = 0;
  for (int n : {779, 1547, 3083, 8192}) {
    for (int t = 0; t < 2; t++) {
      int reps = n >= 4096 ? 5 : 20;
      float ms = time_ms([&] {
        cublasDgemm(h, CUBLAS_OP_N, t ? CUBLAS_OP_T : CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
      }, reps);
      printf("cublasDgemm N%c n=%5d : %8.3f ms  %.2f TFLOP/s\n", t ? 'T' : 'N', n, ms, 2.0 * n * n * n / ms / 1e9);
    }
  }
  // sustained: 8192 for ~3 s
  {
    int n = 8192;
    float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 100);
    printf("cublasDgemm NN n=8192 sustained(100 reps): %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * n * n * n / ms / 1e9);
  }
  return 0;
}
