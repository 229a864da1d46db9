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

// the same in double (the factored butterflies divide them)
__host__ __device__ constexpr double w32_cos_d(int k)
{
    switch (k) {
    case 0:  return 1.0;
    case 1:  return 0.98078528040323044913;
    case 2:  return 0.92387953251128675613;
    case 3:  return 0.83146961230254523708;
    case 4:  return 0.70710678118654752440;
    case 5:  return 0.55557023301960222474;
    case 6:  return 0.38268343236508977173;
    case 7:  return 0.19509032201612826785;
    case 8:  return 0.0;
    default: return k <= 16 ? -w32_cos_d(16 - k) : 0.0;
    }
}
__host__ __device__ constexpr double w32_sin_d(int k) { return k >= 8 ? w32_cos_d(k - 8) : w32_cos_d(8 - k); }

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

// Packed fp32 (sm_100: FADD2 / FFMA2 / FMUL2): a float2 lives in an aligned register pair and one issue slot does both
// components.  The pipe still takes two cycles for a packed instruction (profiles/r2/fp32x2_probe.log: 128 lane-ops/clk/SM
// either way), so packing halves the ISSUE cost of complex arithmetic, not its pipe time.  On the host (fft_regs is also
// compiled for the CPU by tests/dft_regs_host_check.cu, which checks the butterfly networks without a GPU) the same roundings
// are spelled with fmaf.
__host__ __device__ __forceinline__ float2 pk_add(float2 a, float2 b)
{
#ifdef __CUDA_ARCH__
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__host__ __device__ __forceinline__ float2 pk_mul(float2 a, float2 b)
{
#ifdef __CUDA_ARCH__
    return __fmul2_rn(a, b);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
__host__ __device__ __forceinline__ float2 pk_fma(float2 a, float2 b, float2 c)
{
#ifdef __CUDA_ARCH__
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return pk_add(a, b); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return pk_fma(b, make_float2(-1.f, -1.f), a); }   // exact: a - b
__host__ __device__ __forceinline__ float2 cscale(float2 a, float s) { return pk_mul(a, make_float2(s, s)); }
// a * w in two packed instructions: FMUL2 (w.y, w.x) * (-a.y, a.y), then FFMA2 w * (a.x, a.x) + that.  Written in this operand
// order because ptxas folds swap and sign into the FIRST operand's modifiers (Rn.F32x2.LO_HI.NP) and the lane broadcast into the
// second (Rm.F32): no register is moved, and a twiddle that lives in registers across a loop is kept in ONE (swapped) copy.
// The roundings are those of fmaf(a.x, w.x, -(a.y * w.y)), fmaf(a.x, w.y, a.y * w.x).
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 w)
{
    const float2 t = pk_mul(make_float2(w.y, w.x), make_float2(-a.y, a.y));
    return pk_fma(w, make_float2(a.x, a.x), t);
}

#ifndef MMW_FFT_DIF
// ---------------------------------------------------------------------------
// Decimation in time with the twiddle's cosine (or sine) factored out (Linzer-Feig): a butterfly with a general twiddle
// w = c - j s is THREE packed FMAs whose constants are immediates,
//     t = b + (s/c) (b.y, -b.x)        = w b / c
//     a' = a + c t,   b' = a - c t
// (for |s| > |c| the same with c/s and the product turned by -j, so the factored-out constant never exceeds 1 in magnitude),
// w = 1 is an add and a subtract, w = -j two FMAs by +-1 on the swapped operand — the swap and the sign ride in the first
// operand's modifiers, nothing is moved.  A 16-point transform is 74 packed instructions (the decimation-in-frequency form
// this replaced, kept below under MMW_FFT_DIF: 77 packed + 14-22 scalar), a 32-point one 194 (213 + 30).
// ---------------------------------------------------------------------------
// (a, b) <- (a + w b, a - w b), w = exp(-j 2 pi K / 32), K in [0, 16) known at compile time
template <int K>
__host__ __device__ __forceinline__ void dit_bfly(float2 &a, float2 &b)
{
    if constexpr (K == 0) {
        const float2 t = b;
        b = csub(a, t);
        a = cadd(a, t);
    } else if constexpr (K == 8) {                       // w b = (b.y, -b.x)
        const float2 u = make_float2(b.y, -b.x);
        b = pk_fma(u, make_float2(-1.f, -1.f), a);
        a = pk_fma(u, make_float2(1.f, 1.f), a);
    } else {
        constexpr double c = w32_cos_d(K), s = w32_sin_d(K);                 // s > 0 on (0, 16)
        if constexpr ((c < 0 ? -c : c) >= s) {
            constexpr float tau = (float)(s / c), cf = (float)c;
            const float2 t = pk_fma(make_float2(b.y, -b.x), make_float2(tau, tau), b);      // (b.x + tau b.y, b.y - tau b.x)
            b = pk_fma(t, make_float2(-cf, -cf), a);
            a = pk_fma(t, make_float2(cf, cf), a);
        } else {
            constexpr float sig = (float)(c / s), sf = (float)s;
            const float2 t = pk_fma(make_float2(-b.y, b.x), make_float2(sig, sig), b);      // (b.x - sig b.y, b.y + sig b.x) = j w b / s
            const float2 u = make_float2(t.y, -t.x);                                        // w b / s
            b = pk_fma(u, make_float2(-sf, -sf), a);
            a = pk_fma(u, make_float2(sf, sf), a);
        }
    }
}

// In place over the registers x[BASE + STRIDE i], i < N: the half-size transforms of the even and the odd elements, then
// X[k] = E[k] + W_N^k O[k], X[k + N/2] = E[k] - W_N^k O[k].  A transform leaves X[k] in slot bitrev(k) of its own sequence, so
// E[k] and O[k] sit in slots 2 rev(k) and 2 rev(k) + 1 of this one — which are bitrev(k) and bitrev(k + N/2): where the
// outputs belong.
template <int N, int BASE, int STRIDE, int K, int RTOT>
__host__ __device__ __forceinline__ void dit_combine(float2 (&x)[RTOT])
{
    if constexpr (K < N / 2) {
        constexpr int slot = 2 * bitrev(K, ilog2(N) - 1);
        dit_bfly<K * (32 / N)>(x[BASE + STRIDE * slot], x[BASE + STRIDE * (slot + 1)]);
        dit_combine<N, BASE, STRIDE, K + 1, RTOT>(x);
    }
}
template <int N, int BASE, int STRIDE, int RTOT>
__host__ __device__ __forceinline__ void dit_rec(float2 (&x)[RTOT])
{
    if constexpr (N >= 2) {
        dit_rec<N / 2, BASE, STRIDE * 2, RTOT>(x);
        dit_rec<N / 2, BASE + STRIDE, STRIDE * 2, RTOT>(x);
        dit_combine<N, BASE, STRIDE, 0, RTOT>(x);
    }
}

// In-place R-point DFT (R = 2,4,8,16,32), natural order in.  On return X[k] sits in x[bitrev(k, log2 R)];
// callers index with a compile-time bit-reversed subscript, which costs nothing
// once the loops are unrolled.
template <int R>
__host__ __device__ __forceinline__ void dft_regs(float2 (&x)[R])
{
    static_assert(R == 1 || R == 2 || R == 4 || R == 8 || R == 16 || R == 32, "radix");
    dit_rec<R, 0, 1, R>(x);
}

#else   // MMW_FFT_DIF: the radix-2 decimation-in-frequency network of rounds 1-2a (A/B builds only)
// d * (c - j s), c and s compile-time constants: (d.x c + d.y s, d.y c - d.x s) in two packed FMAs whose constants are
// immediates (FFMA2 takes a broadcast 32-bit immediate, FMUL2 does not: hence the FMA with a zero addend for the product);
// roundings: fmaf(d.x, c, d.y * s), fmaf(d.y, c, -(d.x * s)).
__host__ __device__ __forceinline__ float2 crot(float2 d, float c, float s)
{
    const float2 t = pk_fma(make_float2(d.y, -d.x), make_float2(s, s), make_float2(0.f, 0.f));
    return pk_fma(d, make_float2(c, c), t);
}

// (a - b) * exp(-j 2 pi K / 32), K in [0, 16), K known at compile time
template <int K>
__host__ __device__ __forceinline__ float2 sub_mul_w32(float2 a, float2 b)
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
__host__ __device__ __forceinline__ void dif_level(float2 (&x)[RTOT])
{
    if constexpr (I < N / 2) {
        const float2 a = x[BASE + I], b = x[BASE + I + N / 2];
        x[BASE + I] = cadd(a, b);
        x[BASE + I + N / 2] = sub_mul_w32<I * (32 / N)>(a, b);
        dif_level<N, BASE, I + 1, RTOT>(x);
    }
}

template <int N, int BASE, int RTOT>
__host__ __device__ __forceinline__ void dif_rec(float2 (&x)[RTOT])
{
    if constexpr (N >= 2) {
        dif_level<N, BASE, 0, RTOT>(x);
        dif_rec<N / 2, BASE, RTOT>(x);
        dif_rec<N / 2, BASE + N / 2, RTOT>(x);
    }
}

// In-place R-point DFT (R = 2,4,8,16,32).  On return X[k] sits in x[bitrev(k, log2 R)].
template <int R>
__host__ __device__ __forceinline__ void dft_regs(float2 (&x)[R])
{
    static_assert(R == 1 || R == 2 || R == 4 || R == 8 || R == 16 || R == 32, "radix");
    dif_rec<R, 0, R>(x);
}
#endif

}  // namespace mmw
