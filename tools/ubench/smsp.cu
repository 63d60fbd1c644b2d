// Do 64-thread CTAs use only two of the four SM sub-partitions?  Issue-bound FFMA loop,
// same total warps per SM, different CTA shapes.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float *out, int n)
{
    float a = threadIdx.x, b = 1.0001f, c = 0.5f, d = a + 1, e = a + 2, f = a + 3;
    for (int i = 0; i < n; ++i) { a = a * b + c; d = d * b + c; e = e * b + c; f = f * b + c; }
    if (a + d + e + f == 12345.f) out[0] = a;
}
int main()
{
    float *o; cudaMalloc(&o, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int n = 200000;
    int shapes[5] = {32, 64, 128, 256, 512};
    for (int s = 0; s < 5; ++s) {
        int T = shapes[s];
        int ctas_per_sm = 1024 / T;                 // 32 warps per SM in every case
        int grid = 148 * ctas_per_sm;
        k<<<grid, T>>>(o, 1000); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<<<grid, T>>>(o, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double inst = (double)grid * (T / 32) * n * 4.0;
        printf("T=%4d ctas/SM=%2d  %.3f ms  %.2f warp-FFMA/cycle/SM (at 1.965 GHz)\n", T, ctas_per_sm, ms, inst / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
