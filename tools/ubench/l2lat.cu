// Global-memory round-trip latency on B200 as seen by one warp per SM: pointer chase over an
// L2-resident buffer (a) alone, (b) while one thread per CTA polls a flag (grid-barrier style),
// (c) plain vs .cg vs volatile loads; and the cost of an atomicAdd round trip.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chase(const unsigned *next, unsigned *out, long long *cyc, volatile unsigned *flag, int n, int mode)
{
    if (threadIdx.x == 32 && mode == 1) { while (*flag == 0) { __nanosleep(32); } return; }
    if (threadIdx.x >= 32) return;
    unsigned p = (blockIdx.x * 9973u + threadIdx.x * 131u) & ((1u << 22) - 1);
    long long t0 = clock64();
    if (mode == 2) for (int i = 0; i < n; ++i) p = __ldcg(next + p);
    else for (int i = 0; i < n; ++i) p = next[p];
    long long t1 = clock64();
    unsigned a = 0;
    long long t2 = clock64();
    for (int i = 0; i < 64; ++i) a += atomicAdd(out + 1 + ((a + i) & 3), 1u) & 1;
    long long t3 = clock64();
    if (threadIdx.x == 0) { cyc[2 * blockIdx.x] = (t1 - t0) / n; cyc[2 * blockIdx.x + 1] = (t3 - t2) / 64; }
    out[0] = p + a;
    if (blockIdx.x == 0 && threadIdx.x == 0) *flag = 1;
}
int main()
{
    const int N = 1 << 22;   // 16 MB of indices
    unsigned *h = new unsigned[N];
    for (int i = 0; i < N; ++i) h[i] = (unsigned)(((long long)i * 1103515245LL + 12345) & (N - 1));
    unsigned *d, *out, *flag; long long *cyc;
    cudaMalloc(&d, N * 4); cudaMalloc(&out, 64); cudaMalloc(&flag, 4); cudaMalloc(&cyc, 148 * 16);
    cudaMemcpy(d, h, N * 4, cudaMemcpyHostToDevice);
    const char *names[3] = {"alone", "with one poller per CTA", "ld.cg"};
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(flag, 0, 4);
            chase<<<148, 64>>>(d, out, cyc, flag, 2000, mode);
            cudaDeviceSynchronize();
        }
        long long hc[296];
        cudaMemcpy(hc, cyc, sizeof hc, cudaMemcpyDeviceToHost);
        double s = 0, a = 0; for (int i = 0; i < 148; ++i) { s += hc[2 * i]; a += hc[2 * i + 1]; }
        printf("%-26s load round trip %.0f cycles, atomicAdd round trip %.0f cycles (mean over 148 SMs)\n", names[mode], s / 148, a / 148);
    }
    return 0;
}
