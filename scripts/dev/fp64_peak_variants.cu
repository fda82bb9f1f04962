#include <cstdio>
#include <cuda_runtime.h>
template <int CH, int UN, bool REGOPS>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b)
{
    double x[CH];
    for (int c = 0; c < CH; c++) x[c] = threadIdx.x + c;
    double ra = a, rb = b;
    if (REGOPS) { ra += threadIdx.x * 1e-30; rb += threadIdx.x * 1e-30; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < UN; u++)
#pragma unroll
            for (int c = 0; c < CH; c++) x[c] = fma(x[c], ra, rb);
    }
    double s = 0; for (int c = 0; c < CH; c++) s += x[c];
    if (s == 123.456) out[0] = s;
}
template <int CH, int UN, bool REGOPS>
void run(const char* name, int bps)
{
    double* d; cudaMalloc(&d, 64);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * bps, iters = 200000 / UN;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CH, UN, REGOPS><<<blocks, 256>>>(d, iters / 10, 0.999999, 1e-9);
    float best = 1e30;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); k<CH, UN, REGOPS><<<blocks, 256>>>(d, iters, 0.999999, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double fl = 2.0 * blocks * 256.0 * iters * UN * CH;
    printf("%-28s bps=%d  %.3f ms  %.2f TFLOP/s\n", name, bps, best, fl / (best * 1e-3) / 1e12);
    cudaFree(d);
}
int main()
{
    run<8, 16, false>("ch8 un16 const", 8);
    run<8, 16, true>("ch8 un16 reg", 8);
    run<8, 64, true>("ch8 un64 reg", 8);
    run<4, 64, true>("ch4 un64 reg", 8);
    run<16, 32, true>("ch16 un32 reg", 4);
    run<8, 64, true>("ch8 un64 reg bps4", 4);
    run<8, 64, true>("ch8 un64 reg bps2", 2);
    run<2, 128, true>("ch2 un128 reg", 8);
    return 0;
}
