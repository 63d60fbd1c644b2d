// Latency micro-benchmarks (one warp): dependent FP64 add/mul/fma chains, division, LDS chains.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double seed, int n)
{
    __shared__ double sh[1024];
    __shared__ unsigned short idx[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) { sh[i] = 1.0 + 1e-9 * i; idx[i] = (unsigned short)((i * 8 + 8) & 8191); }
    __syncthreads();
    double a = seed, b = seed * 0.5 + 1.0, c = 1.0000001;
    long long t0, t1;
    // DADD chain
    t0 = clock64();
    for (int i = 0; i < n; ++i) a = __dadd_rn(a, c);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DMUL chain
    t0 = clock64();
    for (int i = 0; i < n; ++i) a = __dmul_rn(a, c);
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // DFMA chain
    t0 = clock64();
    for (int i = 0; i < n; ++i) a = __fma_rn(a, c, b);
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // 4 independent DADD chains (throughput)
    double a1 = a, a2 = a + 1, a3 = a + 2, a4 = a + 3;
    t0 = clock64();
    for (int i = 0; i < n; ++i) { a1 = __dadd_rn(a1, c); a2 = __dadd_rn(a2, c); a3 = __dadd_rn(a3, c); a4 = __dadd_rn(a4, c); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    a = a1 + a2 + a3 + a4;
    // division chain
    t0 = clock64();
    for (int i = 0; i < n; ++i) a = __ddiv_rn(b, a) + 1.5;
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // dependent LDS.64 chain through index: x = sh[idx-derived]
    unsigned off = (threadIdx.x * 8) & 8191;
    t0 = clock64();
    for (int i = 0; i < n; ++i) off = idx[off >> 3];
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // LDS.U16 -> LDS.64 -> DMUL -> DADD (one slot, dependent across iterations through acc only)
    double acc = 0;
    t0 = clock64();
    for (int i = 0; i < n; ++i) { unsigned o = idx[(threadIdx.x + i) & 1023]; acc = __dadd_rn(acc, __dmul_rn(c, sh[o >> 3])); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // FP32 FFMA chain for reference
    float f = (float)seed;
    t0 = clock64();
    for (int i = 0; i < n; ++i) f = __fmaf_rn(f, 1.0000001f, 0.5f);
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    out[threadIdx.x] = a + acc + off + f;
}
int main()
{
    double *o; long long *c, h[8];
    cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 8 * 8);
    const int n = 4096;
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(o, c, 1.25, n); cudaDeviceSynchronize(); }
    cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
    const char *nm[8] = {"DADD chain", "DMUL chain", "DFMA chain", "4x DADD indep (per add)", "DDIV+DADD chain", "LDS.U16 chain", "slot LDS16->LDS64->DMUL->DADD", "FFMA chain"};
    for (int i = 0; i < 8; ++i) printf("%-32s %.2f cycles/op\n", nm[i], (double)h[i] / n / (i == 3 ? 4 : 1));
    // throughput: many warps
    return 0;
}
