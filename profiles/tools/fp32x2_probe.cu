// Throughput of packed fp32 (FFMA2 / FADD2 / FMUL2) against scalar FFMA / FADD on sm_100a: element operations per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_probe fp32x2_probe.cu && ./fp32x2_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(float2 *out, int iters, float2 seed)
{
    float2 a[8], b = seed, c = make_float2(seed.y, seed.x);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x + i, i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }          // 2 FFMA
                if (MODE == 1) a[i] = __ffma2_rn(a[i], b, c);                                                   // 1 FFMA2
                if (MODE == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }                               // 2 FADD
                if (MODE == 3) a[i] = __fadd2_rn(a[i], c);                                                      // 1 FADD2
                if (MODE == 4) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i] = __ffma2_rn(a[i], b, c); }            // 1 FFMA + 1 FFMA2
                if (MODE == 5) { a[i].x = a[i].x * b.x; a[i].y = a[i].y * b.y; }                               // 2 FMUL
                if (MODE == 6) a[i] = __fmul2_rn(a[i], b);                                                      // 1 FMUL2
            }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) s = __fadd2_rn(s, a[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char *name, int elem_ops_per_inner, float2 *out, int sms, int clock_khz)
{
    const int iters = 4096, grid = sms * 8;
    probe<MODE><<<grid, 256>>>(out, 16, make_float2(1.0001f, 0.9999f));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<grid, 256>>>(out, iters, make_float2(1.0001f, 0.9999f));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)grid * 256 * iters * 32 * elem_ops_per_inner;      // element operations
    const double clocks = ms * 1e-3 * clock_khz * 1e3;
    printf("%-28s %8.3f ms  %7.1f element-ops / clk / SM  (%.1f Tops/s)\n", name, ms, ops / clocks / sms, ops / ms * 1e-9);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float2 *out;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * sizeof(float2));
    printf("%s, %d SMs, %d kHz nominal (element-ops/clk assume that clock)\n", p.name, p.multiProcessorCount, khz);
    run<0>("2 x FFMA (scalar)", 2, out, p.multiProcessorCount, khz);
    run<1>("1 x FFMA2 (packed)", 2, out, p.multiProcessorCount, khz);
    run<2>("2 x FADD (scalar)", 2, out, p.multiProcessorCount, khz);
    run<3>("1 x FADD2 (packed)", 2, out, p.multiProcessorCount, khz);
    run<4>("1 x FFMA + 1 x FFMA2", 3, out, p.multiProcessorCount, khz);
    run<5>("2 x FMUL (scalar)", 2, out, p.multiProcessorCount, khz);
    run<6>("1 x FMUL2 (packed)", 2, out, p.multiProcessorCount, khz);
    return 0;
}
