// mmw_legacy.cu — drop-in for the reference's cudaProcessing() (acceleration.cu:417-572).
//
// The reference runs, per 200 KB frame, 6 cudaMalloc + 6 cudaFree + 5 cudaMemcpy + 4 device syncs
// + 19 launches (unpack, reshape, extension, bit reversal, 14 global-memory radix-2 stages) and then
// copies the whole fp64 spectrum back to search it on the host.  Here one frame is ONE launch of
// ONE kernel: rx0 gather + int16 unpack + base-frame subtraction + zero pad + a 16 384-point FFT + the
// arg-max, with 4 bytes going back to the host (into mapped pinned memory: no copy).  Batches run one
// CTA per frame with the FFT held entirely in one SM's shared memory (16 x 32 x 32 register
// butterflies, legacy_frame_kernel); a single frame per call — the reference's calling pattern — runs
// on a thread-block cluster of 8 CTAs that exchange their partial spectra over distributed shared
// memory (legacy_cluster_kernel).  Only rx0's rows of a capture are uploaded.  Device state is created
// once and reused.
//
// Numerics: the FFT runs in fp32 (inputs are int16 differences, exactly representable).  So that the
// returned distance is the reference's even when two bins are within fp32 rounding of each other,
// every bin within 1e-4 (relative, power) of the fp32 maximum is re-evaluated in fp64 with a direct
// DFT and the reference's rule (strict >, ascending bin => first maximum wins, cudaBenchMarking.cpp:
// 191-206) is applied to the fp64 values.  The near-tie bins are kept as a bit map of the search range,
// so there is no cap on their number and they are visited in ascending order by construction (a flat
// spectrum re-evaluates every bin: slow — ~20 us per bin — but never a different answer).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>

#include "../../include/mmw_legacy.h"
#include "../../include/mmw_radar.h"
#include "fft_regs.cuh"

namespace mmw {
void set_last_error(const char *fmt, ...);
}

namespace {

using namespace mmw;

constexpr int kS = 100, kC = 128, kA = 4;          // acceleration.cu:8-11
constexpr int kValid = kS * kC;                    // 12 800
constexpr int kN = 16384;                          // nextPow2(12 800), acceleration.cu:465
constexpr int kSearch = 6553;                      // floor(0.4 * 16384), acceleration.cu:522
constexpr int kFrameShorts = kS * kC * kA * 2;     // 102 400
constexpr int kNT = 512;
constexpr int kCandWords = (kSearch + 31) / 32;   // near-tie bins as a bit map of the search range [0, kSearch)
constexpr int kRowShorts = 2 * kS;                 // rx0's I/Q of one chirp: the only part of a frame this chain reads
constexpr int kFullRowShorts = kA * 2 * kS;        // one chirp of all receivers in the capture
constexpr int kPackedFrameShorts = kC * kRowShorts;   // 25 600: what the host path uploads per frame (a quarter of the capture)

__device__ __forceinline__ int phys(int i) { return i + (i >> 5); }   // one pad slot per 32: conflict-free in all three passes

struct LegacyArgs {
    const int16_t *frames;      // [n][frame_stride]: chirp rows of row_stride shorts, rx0's 2 * kS shorts first in each row
    const double2 *base;        // [kValid]
    const float2 *tw;           // [kN] exp(-2 pi i k / kN)
    float2 *spectrum;           // [kN] of the LAST frame of the launch, or nullptr
    int *raw;                   // [n]
    int size;                   // valid shorts per frame, counted in the capture layout ([chirp][rx][sample], kFrameShorts per frame)
    int row_stride;             // kA * 2 * kS for frames in the capture layout, 2 * kS for the packed rx0 rows the host path uploads
    int frame_stride;           // shorts between frames
};

__device__ __forceinline__ void load_sample(const LegacyArgs &a, const int16_t *frame, int n, double &re, double &im)
{
    // rx0 of chirp c, sample s sits at element c*(kA*kS) + s of the [chirp][rx][sample] stream (acceleration.cu:117-150),
    // i.e. at short c * (kA * 2 * kS) + 4 * (s >> 1) + (s & 1) (and + 2 for Q) of the IIQQ-packed frame
    const int c = n / kS, s = n - c * kS;
    const int in_row = 4 * (s >> 1) + (s & 1);
    double i16 = 0, q16 = 0;
    // whole IIQQ groups only: the reference unpacks size / 4 groups (acceleration.cu:94-97), an incomplete last group is not read
    if (c * (kA * 2 * kS) + 4 * (s >> 1) + 4 <= a.size) {
        i16 = (double)frame[c * a.row_stride + in_row];
        q16 = (double)frame[c * a.row_stride + in_row + 2];
    }
    const double2 b = a.base[n];
    re = i16 - b.x;
    im = q16 - b.y;
}

__global__ void __launch_bounds__(kNT, 1) legacy_frame_kernel(LegacyArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *sm = reinterpret_cast<float2 *>(smem_raw);                     // [kN + kN/32]
    __shared__ unsigned long long red_key[kNT / 32];
    __shared__ double red_d[2][kNT / 32];
    __shared__ uint32_t cand_bits[kCandWords];
    __shared__ int n_cand;
    __shared__ unsigned long long best_key_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int16_t *frame = a.frames + (size_t)blockIdx.x * a.frame_stride;
    const bool last = blockIdx.x == gridDim.x - 1;

    // ---- gather rx0, subtract the base frame, zero-pad (acceleration.cu:152-166 with the CPU path's padding) ----
    for (int n = tid; n < kN; n += kNT) {
        float2 v = make_float2(0.f, 0.f);
        if (n < kValid) {
            double re, im;
            load_sample(a, frame, n, re, im);
            v = make_float2((float)re, (float)im);
        }
        sm[phys(n)] = v;
    }
    if (tid == 0) n_cand = 0;
    if (tid < kCandWords) cand_bits[tid] = 0u;
    __syncthreads();

    // ---- pass 1: 1024 radix-16 butterflies, stride 1024 ----
#pragma unroll 1
    for (int j = tid; j < 1024; j += kNT) {
        float2 x[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) x[m] = sm[phys(j + 1024 * m)];
        dft_regs<16>(x);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            float2 v = x[bitrev(q, 4)];
            if (q > 0) v = cmul(v, a.tw[j * q]);
            sm[phys(j + 1024 * q)] = v;
        }
    }
    __syncthreads();
    // ---- pass 2: 16 blocks x 32 radix-32 butterflies, stride 32 ----
    {
        const int blk = tid >> 5, j = tid & 31;
        float2 x[32];
#pragma unroll
        for (int m = 0; m < 32; ++m) x[m] = sm[phys(blk * 1024 + j + 32 * m)];
        dft_regs<32>(x);
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            float2 v = x[bitrev(q, 5)];
            if (q > 0) v = cmul(v, a.tw[16 * j * q]);
            sm[phys(blk * 1024 + j + 32 * q)] = v;
        }
    }
    __syncthreads();
    // ---- pass 3: 512 radix-32 butterflies on contiguous runs; bin k = q1 + 16 q2 + 512 q3 ----
    float mag[13];
    unsigned long long key = 0;
    {
        const int q1 = tid >> 5, q2 = tid & 31;
        float2 x[32];
#pragma unroll
        for (int m = 0; m < 32; ++m) x[m] = sm[tid * 33 + m];
        dft_regs<32>(x);
#pragma unroll
        for (int q3 = 0; q3 < 32; ++q3) {
            const float2 v = x[bitrev(q3, 5)];
            const int k = q1 + 16 * q2 + 512 * q3;
            if (a.spectrum != nullptr && last) a.spectrum[k] = v;
            if (q3 < 13) {
                const float m2 = v.x * v.x + v.y * v.y;
                mag[q3] = m2;
                if (k < kSearch) {
                    const unsigned long long kk = ((unsigned long long)__float_as_uint(m2) << 32) | (unsigned)(0xffffffffu - (unsigned)k);
                    key = kk > key ? kk : key;
                }
            }
        }
    }
    // block arg-max: larger power wins, equal power -> smaller bin (first maximum, strict >)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if (lane == 0) red_key[warp] = key;
    __syncthreads();
    if (tid == 0) {
        unsigned long long b = 0;
        for (int w = 0; w < kNT / 32; ++w) b = red_key[w] > b ? red_key[w] : b;
        best_key_s = b;
    }
    __syncthreads();
    const float best_m = __uint_as_float((unsigned)(best_key_s >> 32));
    int best_k = (int)(0xffffffffu - (unsigned)(best_key_s & 0xffffffffu));
    if (best_key_s == 0ull) best_k = 0;

    // ---- near ties: collect every bin within 1e-4 of the fp32 maximum ----
    {
        const int q1 = tid >> 5, q2 = tid & 31;
        const float thr = best_m * (1.0f - 1e-4f);
#pragma unroll
        for (int q3 = 0; q3 < 13; ++q3) {
            const int k = q1 + 16 * q2 + 512 * q3;
            if (k < kSearch && best_m > 0.f && mag[q3] >= thr) {
                atomicAdd(&n_cand, 1);
                atomicOr(&cand_bits[k >> 5], 1u << (k & 31));
            }
        }
    }
    __syncthreads();
    if (n_cand > 1) {
        double best64 = 0.0;
        int best64_k = 0;
        for (int wi = 0; wi < kCandWords; ++wi)           // ascending bin order; block-uniform control flow
        for (uint32_t bits = cand_bits[wi]; bits; bits &= bits - 1) {
            const int k = 32 * wi + __ffs(bits) - 1;
            double sr = 0.0, si = 0.0;
            for (int n = tid; n < kValid; n += kNT) {
                double xr, xi, s, c;
                load_sample(a, frame, n, xr, xi);
                sincospi((double)((k * n) & (kN - 1)) / (double)(kN / 2), &s, &c);
                sr += xr * c + xi * s;                     // (xr + j xi)(c - j s)
                si += xi * c - xr * s;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sr += __shfl_xor_sync(0xffffffffu, sr, o);
                si += __shfl_xor_sync(0xffffffffu, si, o);
            }
            if (lane == 0) { red_d[0][warp] = sr; red_d[1][warp] = si; }
            __syncthreads();
            if (tid == 0) {
                double r = 0, i = 0;
                for (int w = 0; w < kNT / 32; ++w) { r += red_d[0][w]; i += red_d[1][w]; }
                const double m = r * r + i * i;
                if (m > best64) { best64 = m; best64_k = k; }
            }
            __syncthreads();
        }
        if (tid == 0) a.raw[blockIdx.x] = best64_k;
    } else if (tid == 0) {
        a.raw[blockIdx.x] = best_k;
    }
}

// ---------------------------------------------------------------------------
// The same frame on a thread-block cluster of 8 CTAs (8 SMs, distributed shared memory): 16 384 = 8 x 2048, decimation in
// time.  CTA q gathers the samples n = 8 m + q (1/8 of the unpack, base subtraction and conversion work), runs a
// 2048-point FFT (8 x 16 x 16 register butterflies) in its own shared memory and leaves Y_q[k2] * W_N^(q k2) there; after one
// cluster barrier CTA j reads its 256 values of k2 from all eight CTAs over DSMEM and finishes with the radix-8 butterfly
// X[k2 + 2048 k1] = sum_q W_8^(q k1) Y_q[k2] W_N^(q k2), arg-max included.  The per-CTA maxima and the near-tie candidates meet
// in CTA 0's shared memory; the fp64 re-check of near ties (rare) is CTA 0's.  One frame takes ~1/4 of the single-CTA kernel's time.
// ---------------------------------------------------------------------------
constexpr int kCl = 8;                    // CTAs per cluster = frames are split 8 ways
constexpr int kN2 = kN / kCl;             // 2048-point FFT per CTA
constexpr int kNTc = 256;
__device__ __forceinline__ int phys16(int i) { return i + (i >> 4); }   // one pad slot per 16: conflict-free runs of 16 in the last pass

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `p` (a shared-memory object of this CTA) in CTA `rank` of the cluster, as a shared::cluster address
__device__ __forceinline__ uint32_t dsmem_addr(const void *p, uint32_t rank)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ float2 dsmem_ld_f2(uint32_t addr)
{
    float2 v;
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long dsmem_ld_u64(uint32_t addr)
{
    unsigned long long v;
    asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void dsmem_st_u64(uint32_t addr, unsigned long long v)
{
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void dsmem_atom_or(uint32_t addr, uint32_t v)
{
    asm volatile("red.shared::cluster.or.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t dsmem_atom_inc(uint32_t addr)
{
    uint32_t old;
    asm volatile("atom.shared::cluster.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
    return old;
}

__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kNTc, 1) legacy_cluster_kernel(LegacyArgs a)
{
    __shared__ __align__(16) float2 sm[kN2 + kN2 / 16];        // the CTA's 2048-point transform, padded
    __shared__ __align__(16) float2 yb[kN2];                   // Y_q[k2] W_N^(q k2), natural order: read by the whole cluster
    __shared__ unsigned long long red_key[kNTc / 32];
    __shared__ unsigned long long cta_keys[kCl];               // CTA 0's copy collects the eight maxima
    __shared__ double red_d[2][kNTc / 32];
    __shared__ uint32_t cand_bits[kCandWords];
    __shared__ int n_cand;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t q = cluster_ctarank();
    const int fidx = blockIdx.x / kCl;
    const int16_t *frame = a.frames + (size_t)fidx * a.frame_stride;
    const bool last = fidx == (int)(gridDim.x / kCl) - 1;

    // ---- gather x[8 m + q]: rx0, minus the base frame, zero-padded ----
    for (int m = tid; m < kN2; m += kNTc) {
        const int n = kCl * m + (int)q;
        float2 v = make_float2(0.f, 0.f);
        if (n < kValid) {
            double re, im;
            load_sample(a, frame, n, re, im);
            v = make_float2((float)re, (float)im);
        }
        sm[phys16(m)] = v;
    }
    if (tid == 0) n_cand = 0;
    if (tid < kCandWords) cand_bits[tid] = 0u;
    __syncthreads();

    // ---- pass 1: 256 radix-8 butterflies, stride 256 ----
    {
        const int j = tid;
        float2 x[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) x[m] = sm[phys16(j + 256 * m)];
        dft_regs<8>(x);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float2 v = x[bitrev(k, 3)];
            if (k > 0) v = cmul(v, a.tw[kCl * j * k]);                 // W_2048^(j k)
            sm[phys16(j + 256 * k)] = v;
        }
    }
    __syncthreads();
    // ---- pass 2: 8 blocks x 16 radix-16 butterflies, stride 16 ----
    if (tid < 128) {
        const int blk = tid >> 4, j = tid & 15;
        float2 x[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) x[m] = sm[phys16(blk * 256 + j + 16 * m)];
        dft_regs<16>(x);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float2 v = x[bitrev(k, 4)];
            if (k > 0) v = cmul(v, a.tw[64 * j * k]);                  // W_256^(j k)
            sm[phys16(blk * 256 + j + 16 * k)] = v;
        }
    }
    __syncthreads();
    // ---- pass 3: 128 radix-16 butterflies on contiguous runs; k2 = q1 + 8 q2 + 128 q3; times the cluster twiddle W_N^(q k2) ----
    if (tid < 128) {
        const int q1 = tid >> 4, q2 = tid & 15;
        float2 x[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) x[m] = sm[tid * 17 + m];
        dft_regs<16>(x);
#pragma unroll
        for (int q3 = 0; q3 < 16; ++q3) {
            const int k2 = q1 + 8 * q2 + 128 * q3;
            float2 v = x[bitrev(q3, 4)];
            if (q > 0) v = cmul(v, a.tw[(int)q * k2]);                  // q k2 < 8 * 2048 = kN
            yb[k2] = v;
        }
    }
    cluster_sync_all();

    // ---- radix-8 across the cluster: this CTA finishes k2 in [256 q, 256 q + 256) ----
    float mag[4];
    unsigned long long key = 0;
    {
        const int k2 = 256 * (int)q + tid;
        float2 x[8];
#pragma unroll
        for (int r = 0; r < kCl; ++r) x[r] = dsmem_ld_f2(dsmem_addr(&yb[k2], (uint32_t)r));
        dft_regs<8>(x);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
            const float2 v = x[bitrev(k1, 3)];
            const int k = k2 + kN2 * k1;
            if (a.spectrum != nullptr && last) a.spectrum[k] = v;
            if (k1 < 4) {
                const float m2 = v.x * v.x + v.y * v.y;
                mag[k1] = m2;
                if (k < kSearch) {
                    const unsigned long long kk = ((unsigned long long)__float_as_uint(m2) << 32) | (unsigned)(0xffffffffu - (unsigned)k);
                    key = kk > key ? kk : key;
                }
            }
        }
    }
    // arg-max: larger power wins, equal power -> smaller bin (first maximum, strict >); CTA, then cluster
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if (lane == 0) red_key[warp] = key;
    __syncthreads();
    if (tid == 0) {
        unsigned long long b = 0;
        for (int w = 0; w < kNTc / 32; ++w) b = red_key[w] > b ? red_key[w] : b;
        dsmem_st_u64(dsmem_addr(&cta_keys[q], 0), b);
    }
    cluster_sync_all();
    unsigned long long best_key = 0;
#pragma unroll
    for (int r = 0; r < kCl; ++r) {
        const unsigned long long kk = dsmem_ld_u64(dsmem_addr(&cta_keys[r], 0));
        best_key = kk > best_key ? kk : best_key;
    }
    const float best_m = __uint_as_float((unsigned)(best_key >> 32));
    int best_k = (int)(0xffffffffu - (unsigned)(best_key & 0xffffffffu));
    if (best_key == 0ull) best_k = 0;

    // ---- near ties: every bin within 1e-4 of the fp32 maximum goes to CTA 0's candidate list ----
    {
        const float thr = best_m * (1.0f - 1e-4f);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            const int k = 256 * (int)q + tid + kN2 * k1;
            if (k < kSearch && best_m > 0.f && mag[k1] >= thr) {
                dsmem_atom_inc(dsmem_addr(&n_cand, 0));
                dsmem_atom_or(dsmem_addr(&cand_bits[k >> 5], 0), 1u << (k & 31));
            }
        }
    }
    cluster_sync_all();
    if (q != 0) return;                                        // nothing reads the other CTAs' shared memory any more

    if (n_cand > 1) {
        double best64 = 0.0;
        int best64_k = 0;
        for (int wi = 0; wi < kCandWords; ++wi)           // ascending bin order; block-uniform control flow
        for (uint32_t bits = cand_bits[wi]; bits; bits &= bits - 1) {
            const int k = 32 * wi + __ffs(bits) - 1;
            double sr = 0.0, si = 0.0;
            for (int n = tid; n < kValid; n += kNTc) {
                double xr, xi, sn, cs;
                load_sample(a, frame, n, xr, xi);
                sincospi((double)((k * n) & (kN - 1)) / (double)(kN / 2), &sn, &cs);
                sr += xr * cs + xi * sn;                   // (xr + j xi)(c - j s)
                si += xi * cs - xr * sn;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sr += __shfl_xor_sync(0xffffffffu, sr, o);
                si += __shfl_xor_sync(0xffffffffu, si, o);
            }
            if (lane == 0) { red_d[0][warp] = sr; red_d[1][warp] = si; }
            __syncthreads();
            if (tid == 0) {
                double r = 0, i = 0;
                for (int w = 0; w < kNTc / 32; ++w) { r += red_d[0][w]; i += red_d[1][w]; }
                const double m = r * r + i * i;
                if (m > best64) { best64 = m; best64_k = k; }
            }
            __syncthreads();
        }
        if (tid == 0) a.raw[fidx] = best64_k;
    } else if (tid == 0) {
        a.raw[fidx] = best_k;
    }
}

// ---------------------------------------------------------------------------
// host state (lazy singleton; the reference API has no init/teardown call)
// ---------------------------------------------------------------------------
struct LegacyState {
    bool ready = false;
    cudaStream_t stream = nullptr;
    int16_t *d_frames = nullptr;   // [frames_cap][kC][kRowShorts]: rx0 rows only
    int16_t *h_stage = nullptr;    // pinned, same shape: pageable captures are packed here by the CPU (allocated on first use)
    int stage_cap = 0;
    int frames_cap = 0;
    double2 *d_base = nullptr;
    float2 *d_tw = nullptr;
    float2 *d_spec = nullptr;
    int *d_raw = nullptr;          // device alias of h_raw
    int *h_raw = nullptr;          // pinned, mapped
    int raw_cap = 0;
    double *h_base_copy = nullptr; // last base frame uploaded
    bool base_valid = false;
    bool quiet = false;
    int device = 0;                // the device the state lives on: the one that was current at first use
    int variant = 0;               // 0: kernel picked by batch size, 1: always one CTA per frame, 2: always the 8-CTA cluster
};
LegacyState g;
std::mutex g_mu;                   // every entry point below (shutdown included) holds it
bool g_atexit_registered = false;
bool g_env_read = false;

constexpr int kSmemBytes = (kN + kN / 32) * 8;

#define LCHECK(call)                                                                                      \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            mmw::set_last_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            (void)cudaGetLastError();   /* reported: do not leave it for the next launch's check */       \
            return e_;                                                                                    \
        }                                                                                                 \
    } while (0)

cudaError_t ensure_frames(int n)
{
    if (n <= g.frames_cap) return cudaSuccess;
    if (g.d_frames) cudaFree(g.d_frames);
    if (g.h_raw) cudaFreeHost(g.h_raw);
    g.d_frames = nullptr; g.d_raw = nullptr; g.h_raw = nullptr; g.frames_cap = 0;
    LCHECK(cudaMalloc(&g.d_frames, (size_t)n * kPackedFrameShorts * sizeof(int16_t)));
    // the 4-byte result per frame is written by the kernel straight into mapped pinned host memory: no D2H copy on the
    // single-frame path (one DMA operation less per cudaProcessing() call); stream synchronisation makes it visible
    LCHECK(cudaHostAlloc(&g.h_raw, (size_t)n * sizeof(int), cudaHostAllocMapped));
    LCHECK(cudaHostGetDevicePointer(&g.d_raw, g.h_raw, 0));
    g.frames_cap = n;
    return cudaSuccess;
}

void shutdown_locked();

// Called with g_mu held, first thing in every entry point: binds the state to the device that is current at first use and
// makes that device current on every later call (the caller's thread may have another one current by then).
cudaError_t ensure_init()
{
    if (g.ready) {
        LCHECK(cudaSetDevice(g.device));
        return cudaSuccess;
    }
    if (!g_env_read) {               // the environment is read once per process; mmw_legacy_configure overrides it
        const char *q = getenv("MMW_LEGACY_QUIET");
        g.quiet = q && q[0] && q[0] != '0';
        const char *v = getenv("MMW_LEGACY_VARIANT");
        g.variant = v ? atoi(v) : 0;
        g_env_read = true;
    }
    LCHECK(cudaGetDevice(&g.device));
    LCHECK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    LCHECK(cudaMalloc(&g.d_base, kValid * sizeof(double2)));
    LCHECK(cudaMalloc(&g.d_tw, kN * sizeof(float2)));
    LCHECK(cudaMalloc(&g.d_spec, kN * sizeof(float2)));
    float2 *tw = (float2 *)malloc(kN * sizeof(float2));
    for (int k = 0; k < kN; ++k) {
        const double th = -2.0 * M_PI * (double)k / (double)kN;
        tw[k] = make_float2((float)cos(th), (float)sin(th));
    }
    cudaError_t e = cudaMemcpy(g.d_tw, tw, kN * sizeof(float2), cudaMemcpyHostToDevice);
    free(tw);
    LCHECK(e);
    g.h_base_copy = (double *)malloc(kValid * 2 * sizeof(double));
    LCHECK(cudaFuncSetAttribute(legacy_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    LCHECK(ensure_frames(1));
    g.ready = true;
    if (!g_atexit_registered) {
        atexit(mmw_legacy_shutdown);
        g_atexit_registered = true;
    }
    return cudaSuccess;
}

cudaError_t upload_base(const double *base)
{
    if (g.base_valid && memcmp(g.h_base_copy, base, kValid * 2 * sizeof(double)) == 0) return cudaSuccess;
    memcpy(g.h_base_copy, base, kValid * 2 * sizeof(double));
    LCHECK(cudaMemcpyAsync(g.d_base, g.h_base_copy, kValid * sizeof(double2), cudaMemcpyHostToDevice, g.stream));
    g.base_valid = true;
    return cudaSuccess;
}

// distance formula, operation for operation as acceleration.cu:521-523 / cudaBenchMarking.cpp:301-303
double distance_from_raw(int raw)
{
    const double fs = 2.0e6, lightSpeed = 3.0e08, mu = 5.987e12;
    const int extendedSize = kN;
    double Fs_extend = fs * extendedSize / (kC * kS);
    int maxDisIdx = raw * (kC * kS) / extendedSize;
    return lightSpeed * (((double)maxDisIdx / extendedSize) * Fs_extend) / (2 * mu);
}

// A few frames per launch (the drop-in's one frame per call) are latency-bound: the 8-CTA cluster finishes a frame in about a
// quarter of the single-CTA kernel's time.  Many frames per launch are throughput-bound, and there one CTA per frame wins
// (5.0 M against 2.2 M frames/s device-resident: no cluster barriers, no idle half of the CTA in passes 2 and 3); the two
// lines cross near 60 frames per launch.  mmw_legacy_configure (or MMW_LEGACY_VARIANT at first use) = 1 / 2 forces the
// single-CTA / cluster kernel (profiles/, tests).
cudaError_t launch_frames(const LegacyArgs &a, int n)
{
    const int var = g.variant;
    const bool cluster = var == 2 || (var != 1 && n <= 48);
    if (cluster)
        legacy_cluster_kernel<<<n * kCl, kNTc, 0, g.stream>>>(a);
    else
        legacy_frame_kernel<<<n, kNT, kSmemBytes, g.stream>>>(a);
    return cudaGetLastError();
}

// Only rx0's rows go up: 51 200 of the frame's 204 800 bytes.  A pinned capture is read by one strided (2-D) DMA; a pageable
// one (the reference's malloc'ed `input`, cudaBenchMarking.cpp:350) is packed into a pinned staging buffer by the CPU first.
cudaError_t upload_rx0(const short *frames, int n, int spacing, int per)
{
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, frames) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned && per == kFrameShorts && spacing == kFrameShorts) {
        LCHECK(cudaMemcpy2DAsync(g.d_frames, kRowShorts * sizeof(int16_t), frames, kFullRowShorts * sizeof(int16_t), kRowShorts * sizeof(int16_t),
                                 (size_t)n * kC, cudaMemcpyHostToDevice, g.stream));
        return cudaSuccess;
    }
    if (n > g.stage_cap) {
        if (g.h_stage) cudaFreeHost(g.h_stage);
        g.h_stage = nullptr; g.stage_cap = 0;
        LCHECK(cudaMallocHost(&g.h_stage, (size_t)n * kPackedFrameShorts * sizeof(int16_t)));
        g.stage_cap = n;
    }
    for (int f = 0; f < n; ++f) {
        const short *src = frames + (size_t)f * spacing;
        int16_t *dst = g.h_stage + (size_t)f * kPackedFrameShorts;
        for (int c = 0; c < kC; ++c) {
            const int left = per - c * kFullRowShorts;           // valid shorts from the start of this chirp row on
            if (left <= 0) break;                                // the kernel never reads past `size` (load_sample)
            memcpy(dst + c * kRowShorts, src + c * kFullRowShorts, (size_t)(left < kRowShorts ? left : kRowShorts) * sizeof(int16_t));
        }
    }
    LCHECK(cudaMemcpyAsync(g.d_frames, g.h_stage, (size_t)n * kPackedFrameShorts * sizeof(int16_t), cudaMemcpyHostToDevice, g.stream));
    return cudaSuccess;
}

// The capture goes on the bus first; comparing the base frame with the one already in HBM (204 800 bytes of host memory,
// several microseconds) then runs under that DMA.  *t_prepared: host clock once everything before the launch is queued.
double now_s();
cudaError_t run_frames(const short *frames, int n, const double *base, int size, bool want_spec, double *t_prepared = nullptr)
{
    LCHECK(ensure_init());
    LCHECK(ensure_frames(n));
    const int per = size < kFrameShorts ? size : kFrameShorts;
    LCHECK(upload_rx0(frames, n, per == kFrameShorts ? kFrameShorts : size, per));
    LCHECK(upload_base(base));
    if (t_prepared) *t_prepared = now_s();
    LegacyArgs a;
    a.frames = g.d_frames;
    a.base = g.d_base;
    a.tw = g.d_tw;
    a.spectrum = want_spec ? g.d_spec : nullptr;
    a.raw = g.d_raw;
    a.size = per;
    a.row_stride = kRowShorts;
    a.frame_stride = kPackedFrameShorts;
    LCHECK(launch_frames(a, n));
    LCHECK(cudaStreamSynchronize(g.stream));             // a.raw is mapped host memory (ensure_frames): the results are in g.h_raw
    return cudaSuccess;
}

void shutdown_locked()
{
    if (!g.ready) return;
    const bool quiet = g.quiet;
    const int variant = g.variant;
    // at process exit the context may already be gone; ignore errors
    cudaSetDevice(g.device);
    cudaFree(g.d_frames); cudaFree(g.d_base); cudaFree(g.d_tw); cudaFree(g.d_spec);
    cudaFreeHost(g.h_raw); cudaFreeHost(g.h_stage);
    if (g.stream) cudaStreamDestroy(g.stream);
    free(g.h_base_copy);
    g = LegacyState();
    g.quiet = quiet;                 // configuration survives a shutdown / re-init cycle
    g.variant = variant;
}

double now_s()
{
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
}

}  // namespace

// --------------------------------------------------------------------------- C++-linkage drop-in
double cudaProcessing(short *input_host, Complex_t *host_baseFrame, int size, double *fftTime, double *preProcessTime,
                      double *findMaxTime, double *totalTime)
{
    std::lock_guard<std::mutex> lk(g_mu);
    const double t0 = now_s();
    double t1 = t0;
    cudaError_t e = run_frames(input_host, 1, reinterpret_cast<const double *>(host_baseFrame), size, true, &t1);
    if (e != cudaSuccess) {
        // the reference exits silently with the CUDA error code (acceleration.cu:19-31); keep the exit, add the message
        fprintf(stderr, "cudaProcessing: %s\n", mmw_last_error());
        exit((int)e);
    }
    const double t2 = now_s();
    const double maxDis = distance_from_raw(g.h_raw[0]);
    const double t3 = now_s();
    if (!g.quiet)
        printf("Inner CUDA Timing:single round processing Time %.5f ms, FFT + findMax %.5f ms Reshape %.5f ms Extension %.5f ms\n",
               1000.0 * (t3 - t0), 1000.0 * (t3 - t1), 1000.0 * (t1 - t0), 0.0);
    // the reference accumulates seconds with += (acceleration.cu:534-537)
    if (totalTime) *totalTime += t3 - t0;
    if (findMaxTime) *findMaxTime += t3 - t2;
    if (preProcessTime) *preProcessTime += t1 - t0;     // rx0 rows packed and queued for upload, base frame checked: everything before the launch
    if (fftTime) *fftTime += t3 - t1;                   // fused kernel (unpack .. FFT .. arg-max), result read-back, formula
    return maxDis;
}

// --------------------------------------------------------------------------- C ABI
extern "C" {

double mmw_legacy_process_frame(const short *frame_host, const double *base_frame_host, int size, int *raw_index)
{
    if (!frame_host || !base_frame_host || size <= 0) {
        mmw::set_last_error("mmw_legacy_process_frame: bad argument");
        return (double)MMW_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (run_frames(frame_host, 1, base_frame_host, size, true) != cudaSuccess) return (double)MMW_ERR_CUDA;
    if (raw_index) *raw_index = g.h_raw[0];
    return distance_from_raw(g.h_raw[0]);
}

int mmw_legacy_process_frames(const short *frames_host, int n_frames, const double *base_frame_host, int size, double *distances,
                              int *raw_indices)
{
    if (!frames_host || !base_frame_host || n_frames <= 0 || size <= 0 || !distances) {
        mmw::set_last_error("mmw_legacy_process_frames: bad argument");
        return MMW_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (run_frames(frames_host, n_frames, base_frame_host, size, true) != cudaSuccess) return MMW_ERR_CUDA;
    for (int f = 0; f < n_frames; ++f) {
        distances[f] = distance_from_raw(g.h_raw[f]);
        if (raw_indices) raw_indices[f] = g.h_raw[f];
    }
    return MMW_OK;
}

int mmw_legacy_process_device(const short *frames_dev, int n_frames, const double *base_frame_host, int *raw_dev)
{
    if (!frames_dev || !base_frame_host || n_frames <= 0 || !raw_dev) {
        mmw::set_last_error("mmw_legacy_process_device: bad argument");
        return MMW_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (ensure_init() != cudaSuccess || upload_base(base_frame_host) != cudaSuccess) return MMW_ERR_CUDA;
    LegacyArgs a;
    a.frames = frames_dev;
    a.base = g.d_base;
    a.tw = g.d_tw;
    a.spectrum = nullptr;
    a.raw = raw_dev;
    a.size = kFrameShorts;
    a.row_stride = kFullRowShorts;          // frames in HBM keep the capture layout
    a.frame_stride = kFrameShorts;
    if (launch_frames(a, n_frames) != cudaSuccess) {
        mmw::set_last_error("legacy_frame_kernel launch failed");
        return MMW_ERR_CUDA;
    }
    return MMW_OK;
}

int mmw_legacy_sync(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g.ready) return MMW_OK;
    cudaSetDevice(g.device);
    if (cudaStreamSynchronize(g.stream) != cudaSuccess) {
        mmw::set_last_error("mmw_legacy_sync: %s", cudaGetErrorString(cudaGetLastError()));
        return MMW_ERR_CUDA;
    }
    return MMW_OK;
}

double mmw_legacy_distance_from_raw(int raw_index) { return distance_from_raw(raw_index); }

// The reference's cudaTiming() (cudaBenchMarking.cpp:334-395) in one call.
int mmw_legacy_process_file(const char *path, double *distances, int *raw_indices, int capacity, int *n_frames)
{
    if (n_frames) *n_frames = 0;
    if (!path || capacity < 0 || (capacity > 0 && !distances)) {
        mmw::set_last_error("mmw_legacy_process_file: bad argument");
        return MMW_ERR_ARG;
    }
    FILE *fp = fopen(path, "rb");
    if (!fp) {
        mmw::set_last_error("unable to read the specified file: %s", path);     // the reference's message (:346)
        return MMW_ERR_ARG;
    }
    constexpr int kBatch = 256;
    short *buf = (short *)malloc((size_t)kBatch * kFrameShorts * sizeof(short));
    double *base = (double *)malloc((size_t)kValid * 2 * sizeof(double));
    int rc = MMW_OK, done = 0;
    // frame 0 -> base frame, rx0 only, [chirp][sample] (ReshapeComplex_t + memmove, cudaBenchMarking.cpp:357-365)
    size_t got = fread(buf, sizeof(short), kFrameShorts, fp);
    if (got == 0) rc = MMW_ERR_ARG, mmw::set_last_error("mmw_legacy_process_file: empty capture %s", path);
    if (rc == MMW_OK) {
        for (int n = 0; n < kValid; ++n) {
            const int c = n / kS, sidx = n - c * kS;
            const int e = c * (kA * kS) + sidx;
            const size_t grp = (size_t)4 * (e >> 1), gidx = grp + (e & 1);
            const bool whole = grp + 4 <= got;                   // whole IIQQ groups only, as in load_sample
            base[2 * n] = whole ? (double)buf[gidx] : 0.0;
            base[2 * n + 1] = whole ? (double)buf[gidx + 2] : 0.0;
        }
        std::lock_guard<std::mutex> lk(g_mu);
        while (rc == MMW_OK && (got = fread(buf, sizeof(short), (size_t)kBatch * kFrameShorts, fp)) > 0) {
            const int whole = (int)(got / kFrameShorts), rest = (int)(got % kFrameShorts);
            if (whole > 0 && run_frames(buf, whole, base, kFrameShorts, false) != cudaSuccess) { rc = MMW_ERR_CUDA; break; }
            for (int f = 0; f < whole; ++f, ++done)
                if (done < capacity) {
                    distances[done] = distance_from_raw(g.h_raw[f]);
                    if (raw_indices) raw_indices[done] = g.h_raw[f];
                }
            if (rest) {       // short final read: the reference hands the short count on and processes the frame (:374-377)
                if (run_frames(buf + (size_t)whole * kFrameShorts, 1, base, rest, false) != cudaSuccess) { rc = MMW_ERR_CUDA; break; }
                if (done < capacity) {
                    distances[done] = distance_from_raw(g.h_raw[0]);
                    if (raw_indices) raw_indices[done] = g.h_raw[0];
                }
                ++done;
            }
        }
    }
    free(buf);
    free(base);
    fclose(fp);
    if (n_frames) *n_frames = done;
    return rc;
}

int mmw_legacy_copy_spectrum(float *out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g.ready || !out) {
        mmw::set_last_error("mmw_legacy_copy_spectrum: no frame processed yet");
        return MMW_ERR_STATE;
    }
    cudaSetDevice(g.device);
    if (cudaMemcpy(out, g.d_spec, kN * sizeof(float2), cudaMemcpyDeviceToHost) != cudaSuccess) return MMW_ERR_CUDA;
    return MMW_OK;
}

void mmw_legacy_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    shutdown_locked();
}

int mmw_legacy_configure(int kernel_variant, int quiet)
{
    if (kernel_variant > 2) { mmw::set_last_error("mmw_legacy_configure: kernel_variant must be 0, 1, 2 or negative (keep)"); return MMW_ERR_ARG; }
    std::lock_guard<std::mutex> lk(g_mu);
    g_env_read = true;               // explicit configuration wins over the environment from here on
    if (kernel_variant >= 0) g.variant = kernel_variant;
    if (quiet >= 0) g.quiet = quiet != 0;
    return MMW_OK;
}

}  // extern "C"
