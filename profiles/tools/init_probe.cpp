// init_probe.cpp — where the first cudaProcessing() call spends its time: CUDA context creation vs the library's own
// lazy initialisation vs a steady-state call.   g++ -O2 init_probe.cpp -I../../include <lib> -L/usr/local/cuda/lib64 -lcudart
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "mmw_legacy.h"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main()
{
    setenv("MMW_LEGACY_QUIET", "1", 1);
    short *frame = (short *)calloc(102400, sizeof(short));
    Complex_t *base = (Complex_t *)calloc(12800, sizeof(Complex_t));
    for (int i = 0; i < 102400; ++i) frame[i] = (short)((i * 37) % 201 - 100);
    double t0 = now();
    cudaFree(0);
    double t1 = now();
    double a = 0, b = 0, c = 0, d = 0;
    cudaProcessing(frame, base, 102400, &a, &b, &c, &d);
    double t2 = now();
    cudaProcessing(frame, base, 102400, &a, &b, &c, &d);
    double t3 = now();
    for (int i = 0; i < 100; ++i) cudaProcessing(frame, base, 102400, &a, &b, &c, &d);
    double t4 = now();
    printf("context creation (cudaFree(0)) %.1f ms | first cudaProcessing %.1f ms | second %.3f ms | steady state %.3f ms/call\n",
           1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3) / 100);
    return 0;
}
