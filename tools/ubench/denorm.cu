// Are FP64 ops on subnormal operands slow on B200?  One warp, dependent chains.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double seed, double tiny, int n)
{
    long long t0, t1;
    double a = seed, c = 1.0;
    // normal: DMUL chain by 1.0
    t0 = clock64(); for (int i = 0; i < n; ++i) a = __dmul_rn(a, c); t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double d = tiny;   // subnormal
    t0 = clock64(); for (int i = 0; i < n; ++i) d = __dmul_rn(d, c); t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    double e = tiny;
    t0 = clock64(); for (int i = 0; i < n; ++i) e = __dadd_rn(e, tiny); t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // normal * normal -> subnormal result
    double f = 1e-200, g = 0;
    t0 = clock64(); for (int i = 0; i < n; ++i) g = __dadd_rn(g, __dmul_rn(f, 1e-110)); t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // compare on subnormal
    int cnt = 0; double h = tiny;
    t0 = clock64(); for (int i = 0; i < n; ++i) { if (h > seed * 1e-300) cnt++; h = __dadd_rn(h, 0.0); } t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // division with tiny numerator (slow path) and normal
    double q = 0;
    t0 = clock64(); for (int i = 0; i < n; ++i) q = __dadd_rn(__ddiv_rn(tiny, seed + q), 0.0); t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    double q2 = 0;
    t0 = clock64(); for (int i = 0; i < n; ++i) q2 = __dadd_rn(__ddiv_rn(1e-280, seed + q2), 0.0); t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    double q3 = 0;
    t0 = clock64(); for (int i = 0; i < n; ++i) q3 = __dadd_rn(__ddiv_rn(0.0 * seed, seed + q3), 0.0); t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    out[threadIdx.x] = a + d + e + g + cnt + h + q + q2 + q3;
}
int main()
{
    double *o; long long *c, h[8];
    cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 8 * 8);
    const int n = 2048;
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(o, c, 1.25, 1e-310, n); cudaDeviceSynchronize(); }
    cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
    const char *nm[8] = {"DMUL normal", "DMUL subnormal operand", "DADD subnormal", "DMUL->subnormal result + DADD", "DSETP subnormal + DADD", "DDIV tiny(1e-310)/normal", "DDIV 1e-280/normal", "DDIV 0/normal"};
    for (int i = 0; i < 8; ++i) printf("%-36s %.2f cycles/iter\n", nm[i], (double)h[i] / n);
    return 0;
}
