// Host-side check of the in-register butterfly networks of csrc/fft_regs.cuh: every radix the kernels instantiate is run on
// the CPU (the header spells its packed arithmetic with fmaf there, i.e. the device's roundings) and compared with a direct
// fp64 DFT, forward e^{-j 2 pi k n / R}, unnormalised (cudaBenchMarking.cpp:88-104 in the reference).  Prints one line per
// radix: "R max_abs_err/max_abs_out"; exit code 1 if any exceeds 2e-6.  Built and run by tests/test_host_logic.py (no GPU).
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../cuda-based-mmwave-radar-object-detection-acceleration_b200/csrc/fft_regs.cuh"

template <int R>
static double check(unsigned seed)
{
    double worst = 0;
    srand(seed);
    for (int trial = 0; trial < 64; ++trial) {
        float2 x[R];
        double xr[R], xi[R];
        for (int n = 0; n < R; ++n) {
            // trial 0: impulse train; trial 1: ramp (the reference's fftTest vector, acceleration.cu:361-365); else random int16-like
            float re, im;
            if (trial == 0) { re = n == 1 ? 1.f : 0.f; im = 0.f; }
            else if (trial == 1) { re = (float)(n + 1); im = 0.f; }
            else { re = (float)(rand() % 65536 - 32768); im = (float)(rand() % 65536 - 32768); }
            x[n] = make_float2(re, im);
            xr[n] = re; xi[n] = im;
        }
        mmw::dft_regs<R>(x);
        double maxo = 0, maxe = 0;
        for (int k = 0; k < R; ++k) {
            double sr = 0, si = 0;
            for (int n = 0; n < R; ++n) {
                const double ang = -2.0 * M_PI * (double)((k * n) % R) / R;
                sr += xr[n] * cos(ang) - xi[n] * sin(ang);
                si += xr[n] * sin(ang) + xi[n] * cos(ang);
            }
            const float2 got = x[mmw::bitrev(k, mmw::ilog2(R))];
            maxo = fmax(maxo, hypot(sr, si));
            maxe = fmax(maxe, hypot(got.x - sr, got.y - si));
        }
        worst = fmax(worst, maxe / maxo);
    }
    printf("%d %.3e\n", R, worst);
    return worst;
}

int main()
{
    double w = 0;
    w = fmax(w, check<2>(1));
    w = fmax(w, check<4>(2));
    w = fmax(w, check<8>(3));
    w = fmax(w, check<16>(4));
    w = fmax(w, check<32>(5));
    return w < 2e-6 ? 0 : 1;
}
