// fft_regs.cuh — in-register DFTs of 2..32 points with compile-time twiddles.
//
// Every FFT in the pipeline is "lane = batch element": the 32 lanes of a warp
// hold 32 independent transforms (32 chirps in the range stage, 32 range bins
// in the Doppler stage) and each thread runs whole radix-R butterflies on
// registers.  The butterfly index is therefore warp-uniform, so the inner
// twiddles are literals that ptxas folds into FMUL/FFMA immediates and no
// cross-lane exchange is ever needed.
//
// Math convention = the reference's: forward transform, e^{-j 2 pi k n / N},
// unnormalised (cudaBenchMarking.cpp:88-104, acceleration.cu:202-247).
#pragma once
#include <cuda_runtime.h>

namespace mmw {

// cos(2 pi k / 32), k = 0..16
__host__ __device__ constexpr float w32_cos(int k)
{
    switch (k) {
    case 0:  return 1.0f;
    case 1:  return 0.98078528040323044913f;
    case 2:  return 0.92387953251128675613f;
    case 3:  return 0.83146961230254523708f;
    case 4:  return 0.70710678118654752440f;
    case 5:  return 0.55557023301960222474f;
    case 6:  return 0.38268343236508977173f;
    case 7:  return 0.19509032201612826785f;
    case 8:  return 0.0f;
    case 9:  return -0.19509032201612826785f;
    case 10: return -0.38268343236508977173f;
    case 11: return -0.55557023301960222474f;
    case 12: return -0.70710678118654752440f;
    case 13: return -0.83146961230254523708f;
    case 14: return -0.92387953251128675613f;
    case 15: return -0.98078528040323044913f;
    default: return -1.0f;
    }
}
// sin(2 pi k / 32) = cos(2 pi (k - 8) / 32), k = 0..16
__host__ __device__ constexpr float w32_sin(int k)
{
    return k >= 8 ? w32_cos(k - 8) : w32_cos(8 - k);
}

__host__ __device__ constexpr int bitrev(int v, int bits)
{
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((v >> b) & 1) << (bits - 1 - b);
    return r;
}
__host__ __device__ constexpr int ilog2(int n)
{
    int l = 0;
    while ((1 << l) < n) ++l;
    return l;
}

// Complex add / subtract are single packed-fp32 instructions on sm_100 (FADD2 / FFMA2): a float2 lives in
// an aligned register pair, so one issue slot does both components.  The FFT kernels are issue-bound,
// not FP-pipe-bound, so this is worth ~1/3 of the butterfly instruction count.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }   // exact: a - b
__device__ __forceinline__ float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
// a * w in two packed instructions: FMUL2 (w.y, w.x) * (-a.y, a.y), then FFMA2 w * (a.x, a.x) + that.  Written in this operand
// order because ptxas folds swap and sign into the FIRST operand's modifiers (Rn.F32x2.LO_HI.NP) and the lane broadcast into the
// second (Rm.F32): no register is moved, and a twiddle that lives in registers across a loop is kept in ONE (swapped) copy.
// The roundings are those of fmaf(a.x, w.x, -(a.y * w.y)), fmaf(a.x, w.y, a.y * w.x).
__device__ __forceinline__ float2 cmul(float2 a, float2 w)
{
    const float2 t = __fmul2_rn(make_float2(w.y, w.x), make_float2(-a.y, a.y));
    return __ffma2_rn(w, make_float2(a.x, a.x), t);
}
// d * (c - j s), c and s compile-time constants: (d.x c + d.y s, d.y c - d.x s) in two packed FMAs whose constants are
// immediates (FFMA2 takes a broadcast 32-bit immediate, FMUL2 does not: hence the FMA with a zero addend for the product);
// roundings: fmaf(d.x, c, d.y * s), fmaf(d.y, c, -(d.x * s)).
__device__ __forceinline__ float2 crot(float2 d, float c, float s)
{
    const float2 t = __ffma2_rn(make_float2(d.y, -d.x), make_float2(s, s), make_float2(0.f, 0.f));
    return __ffma2_rn(d, make_float2(c, c), t);
}

// (a - b) * exp(-j 2 pi K / 32), K in [0, 16), K known at compile time
template <int K>
__device__ __forceinline__ float2 sub_mul_w32(float2 a, float2 b)
{
    if constexpr (K == 0) {
        return csub(a, b);
    } else if constexpr (K == 8) {                       // * -j : (d.y, -d.x)
        return make_float2(a.y - b.y, b.x - a.x);
    } else {                                             // (K = 4, 12 included: c = +-s = 1/sqrt2)
        constexpr float c = w32_cos(K);
        constexpr float s = w32_sin(K);                  // w = c - j s
        return crot(csub(a, b), c, s);
    }
}

// one radix-2 decimation-in-frequency level over x[BASE .. BASE+N), butterfly I
template <int N, int BASE, int I, int RTOT>
__device__ __forceinline__ void dif_level(float2 (&x)[RTOT])
{
    if constexpr (I < N / 2) {
        const float2 a = x[BASE + I], b = x[BASE + I + N / 2];
        x[BASE + I] = cadd(a, b);
        x[BASE + I + N / 2] = sub_mul_w32<I * (32 / N)>(a, b);
        dif_level<N, BASE, I + 1, RTOT>(x);
    }
}

template <int N, int BASE, int RTOT>
__device__ __forceinline__ void dif_rec(float2 (&x)[RTOT])
{
    if constexpr (N >= 2) {
        dif_level<N, BASE, 0, RTOT>(x);
        dif_rec<N / 2, BASE, RTOT>(x);
        dif_rec<N / 2, BASE + N / 2, RTOT>(x);
    }
}

// In-place R-point DFT (R = 2,4,8,16,32).  On return X[k] sits in x[bitrev(k, log2 R)];
// callers index with a compile-time bit-reversed subscript, which costs nothing
// once the loops are unrolled.
template <int R>
__device__ __forceinline__ void dft_regs(float2 (&x)[R])
{
    static_assert(R == 1 || R == 2 || R == 4 || R == 8 || R == 16 || R == 32, "radix");
    dif_rec<R, 0, R>(x);
}

}  // namespace mmw
