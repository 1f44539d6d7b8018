// smid_probe.cu — which %smid values a full-GPU grid sees on this part (the GEMM's SM reservation filters on %smid).
#include <cstdio>
#include <cuda_runtime.h>
#include <set>
#include <vector>
__global__ void probe(unsigned* out, unsigned* nsm) {
    unsigned s, n;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    asm volatile("mov.u32 %0, %%nsmid;" : "=r"(n));
    if (threadIdx.x == 0) { out[blockIdx.x] = s; *nsm = n; }
    // hold the SM for a while so that the grid spreads over every SM
    long long t0 = clock64();
    while (clock64() - t0 < 200000) { }
}
int main() {
    const int G = 148 * 8;
    unsigned *d, *dn;
    cudaMalloc(&d, G * 4); cudaMalloc(&dn, 4);
    probe<<<G, 128, 32 * 1024>>>(d, dn);
    cudaDeviceSynchronize();
    std::vector<unsigned> h(G); unsigned n;
    cudaMemcpy(h.data(), d, G * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&n, dn, 4, cudaMemcpyDeviceToHost);
    std::set<unsigned> u(h.begin(), h.end());
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("multiProcessorCount %d, %%nsmid %u, distinct smid seen %zu, min %u max %u\n", p.multiProcessorCount, n, u.size(), *u.begin(), *u.rbegin());
    printf("first 16 CTAs -> smid:"); for (int i = 0; i < 16; ++i) printf(" %u", h[i]); printf("\n");
    return 0;
}
