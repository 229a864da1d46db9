// mmw_pipeline.cu — the four stages of the batched radar chain as sm_100a kernels.
//
//   K1 range_fft_kernel    int16 IIQQ unpack + range window + Sp-point FFT per (chirp, antenna);
//                          output written corner-turned ([ant][range][chirp]) and pre-multiplied
//                          by the Doppler window.                  (replaces acceleration.cu:91-150)
//   K2 doppler_fft_kernel  Cp-point FFT across chirps per (range, antenna), |X|^2 accumulated over
//                          antennas in registers -> power map; Doppler cube optional.
//   K3 cfar_kernel         2-D CA-CFAR on the power map -> bit mask   (no reference counterpart)
//   K4 detect_kernel       ordered compaction of the mask, noise re-evaluation, 3x3 peak grouping,
//                          angle spectrum arg-max -> detection records
//   K5 compact_kernel      per-frame lists -> one dense list + header (what D2H / NCCL moves)
//
// Both FFT kernels use the same scheme: a tile is BT transforms x N points held in shared
// memory, staged in by 1-D TMA bulk copies (one contiguous row per transform), the 32 lanes
// of a warp are 32 different transforms ("lane = batch"), and the transform is a two-pass
// (R1 x R2) decimation-in-frequency whose radix-R1 / radix-R2 butterflies run entirely in
// registers (fft_regs.cuh).  Because lanes are batch elements, the final pass writes the
// corner-turned layout directly with fully coalesced stores: the transpose never exists as a
// separate step and never round-trips HBM.
#include "fft_regs.cuh"
#include "mmw_common.cuh"

namespace mmw {

// ---------------------------------------------------------------------------
// K1: range FFT
// ---------------------------------------------------------------------------
template <int N, int BT>
struct RangeSmem {
    static constexpr int kStageStride = 4 * N + 16;                 // bytes per staged int16 row
    static constexpr int kOffTw = 16;
    static constexpr int kOffWin = kOffTw + 8 * N;
    static constexpr int kOffStage = kOffWin + 4 * N;
    static constexpr int kOffWork = kOffStage + BT * kStageStride;
    static constexpr int kBytes = kOffWork + BT * (N + 1) * 8;
};

template <int N, int R1, int R2, int BT, int NW, bool PAIR>
__global__ void __launch_bounds__(NW * 32) range_fft_kernel(PlanDev p, const int16_t *__restrict__ adc, float2 *__restrict__ rs)
{
    static_assert(R1 * R2 == N, "plan");
    using L = RangeSmem<N, BT>;
    constexpr int NT = NW * 32;
    constexpr int SUBS = 32 / BT;
    constexpr int NSLOT = NW * SUBS;
    constexpr int LR1 = ilog2(R1), LR2 = ilog2(R2);

    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    float2 *tw = reinterpret_cast<float2 *>(smem + L::kOffTw);
    float *win = reinterpret_cast<float *>(smem + L::kOffWin);
    unsigned char *stage = smem + L::kOffStage;
    float2 *work = reinterpret_cast<float2 *>(smem + L::kOffWork);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = lane % BT, sub = lane / BT, slot = warp * SUBS + sub;
    const int S = p.S, C = p.C, A = p.A;

    const int nct = (C + BT - 1) / BT;
    int t = blockIdx.x;
    const int ct = t % nct;  t /= nct;
    const int a = t % A;
    const int f = t / A;
    const int c0 = ct * BT;
    const int nrows = min(BT, C - c0);

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(nrows * S * 4));
        __syncwarp();
        if (lane < nrows) {
            const int16_t *src = adc + (((size_t)f * C + c0 + lane) * A + a) * (size_t)(2 * S);
            bulk_g2s(stage + lane * L::kStageStride, src, (uint32_t)(S * 4), bar);
        }
    }
    for (int i = tid; i < N; i += NT) tw[i] = p.tw_r[i];
    for (int i = tid; i < N; i += NT) win[i] = i < S ? p.win_r[i] : 0.f;
    const bool row_ok = row < nrows;
    const float wdop = row_ok ? p.win_d[c0 + row] : 0.f;
    __syncthreads();
    mbar_wait(bar, 0);

    const unsigned char *srow = stage + row * L::kStageStride;
    float2 *wrow = work + row * (N + 1);

    // ---- pass 1: R2 butterflies of radix R1 over stride R2, reading the staged int16 rows ----
    if constexpr (PAIR) {
#pragma unroll 1
        for (int u = slot; u < R2 / 2; u += NSLOT) {
            const int n2 = 2 * u;
            float2 xa[R1], xb[R1];
#pragma unroll
            for (int m = 0; m < R1; ++m) {
                const int n = n2 + m * R2;
                if (n < S) {
                    const uint2 raw = *reinterpret_cast<const uint2 *>(srow + 4 * n);   // [I(n) I(n+1)] [Q(n) Q(n+1)]
                    const float2 w = *reinterpret_cast<const float2 *>(win + n);
                    xa[m] = make_float2((float)(short)(raw.x & 0xffffu) * w.x, (float)(short)(raw.y & 0xffffu) * w.x);
                    xb[m] = make_float2((float)((int)raw.x >> 16) * w.y, (float)((int)raw.y >> 16) * w.y);
                } else {
                    xa[m] = xb[m] = make_float2(0.f, 0.f);
                }
            }
            dft_regs<R1>(xa);
            dft_regs<R1>(xb);
#pragma unroll
            for (int k1 = 0; k1 < R1; ++k1) {
                float2 va = xa[bitrev(k1, LR1)], vb = xb[bitrev(k1, LR1)];
                if (k1 > 0) {
                    va = cmul(va, tw[n2 * k1]);
                    vb = cmul(vb, tw[(n2 + 1) * k1]);
                }
                wrow[k1 * R2 + n2] = va;
                wrow[k1 * R2 + n2 + 1] = vb;
            }
        }
    } else {
#pragma unroll 1
        for (int u = slot; u < R2; u += NSLOT) {
            const int n2 = u;
            const int odd = n2 & 1;
            float2 x[R1];
#pragma unroll
            for (int m = 0; m < R1; ++m) {
                const int n = n2 + m * R2;
                if (n < S) {
                    const uint2 raw = *reinterpret_cast<const uint2 *>(srow + 4 * (n - odd));
                    const float w = win[n];
                    const int iv = odd ? ((int)raw.x >> 16) : (int)(short)(raw.x & 0xffffu);
                    const int qv = odd ? ((int)raw.y >> 16) : (int)(short)(raw.y & 0xffffu);
                    x[m] = make_float2((float)iv * w, (float)qv * w);
                } else {
                    x[m] = make_float2(0.f, 0.f);
                }
            }
            dft_regs<R1>(x);
#pragma unroll
            for (int k1 = 0; k1 < R1; ++k1) {
                float2 v = x[bitrev(k1, LR1)];
                if (k1 > 0) v = cmul(v, tw[n2 * k1]);
                wrow[k1 * R2 + n2] = v;
            }
        }
    }
    __syncthreads();

    // ---- pass 2: R1 butterflies of radix R2 on contiguous runs; outputs go straight to HBM ----
    float2 *out = rs + ((size_t)f * A + a) * (size_t)N * C + c0 + row;
#pragma unroll 1
    for (int u = slot; u < R1; u += NSLOT) {
        const int k1 = u;
        float2 y[R2];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) y[n2] = wrow[k1 * R2 + n2];
        dft_regs<R2>(y);
        if (row_ok) {
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) {
                const float2 v = y[bitrev(k2, LR2)];
                const int k = k1 + R1 * k2;
                st_global_f2(out + (size_t)k * C, make_float2(v.x * wdop, v.y * wdop));
            }
        }
    }
}

// ---------------------------------------------------------------------------
// K2: Doppler FFT + non-coherent integration
// ---------------------------------------------------------------------------
template <int N, int BT>
struct DopplerSmem {
    static constexpr int kStageRow = N + 2;                          // float2 per staged row (16-byte multiple)
    static constexpr int kOffTw = 16;
    static constexpr int kOffStage = kOffTw + 8 * N;
    static constexpr int kStageBytes = BT * kStageRow * 8;
    static constexpr int kOffWork = kOffStage + 2 * kStageBytes;
    static constexpr int kBytes = kOffWork + BT * (N + 1) * 8;
};

template <int N, int R1, int R2, int BT, int NW>
__global__ void __launch_bounds__(NW * 32) doppler_fft_kernel(PlanDev p, const float2 *__restrict__ rs, float2 *__restrict__ cube,
                                                               float *__restrict__ pmap)
{
    static_assert(R1 * R2 == N, "plan");
    using L = DopplerSmem<N, BT>;
    constexpr int NT = NW * 32;
    constexpr int SUBS = 32 / BT;
    constexpr int NSLOT = NW * SUBS;
    constexpr int LR1 = ilog2(R1), LR2 = ilog2(R2);
    constexpr int UPS2 = (R1 + NSLOT - 1) / NSLOT;                   // pass-2 butterflies per slot

    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);              // two barriers
    float2 *tw = reinterpret_cast<float2 *>(smem + L::kOffTw);
    float2 *stage = reinterpret_cast<float2 *>(smem + L::kOffStage);
    float2 *work = reinterpret_cast<float2 *>(smem + L::kOffWork);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = lane % BT, sub = lane / BT, slot = warp * SUBS + sub;
    const int C = p.C, A = p.A, Sp = p.Sp;

    const int nrt = Sp / BT;
    const int rt = blockIdx.x % nrt;
    const int f = blockIdx.x / nrt;
    const int r0 = rt * BT;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int a) {
        if (warp == 0) {
            uint64_t *b = &bar[a & 1];
            if (lane == 0) mbar_arrive_expect_tx(b, (uint32_t)(BT * C * 8));
            __syncwarp();
            if (lane < BT) {
                const float2 *src = rs + (((size_t)f * A + a) * Sp + r0 + lane) * (size_t)C;
                bulk_g2s(stage + (size_t)(a & 1) * (BT * L::kStageRow) + lane * L::kStageRow, src, (uint32_t)(C * 8), b);
            }
        }
    };
    issue(0);
    for (int i = tid; i < N; i += NT) tw[i] = p.tw_d[i];

    float acc[UPS2][R2];
#pragma unroll
    for (int i = 0; i < UPS2; ++i)
#pragma unroll
        for (int j = 0; j < R2; ++j) acc[i][j] = 0.f;

    float2 *wrow = work + row * (N + 1);
    __syncthreads();

#pragma unroll 1
    for (int a = 0; a < A; ++a) {
        if (a + 1 < A) issue(a + 1);
        mbar_wait(&bar[a & 1], (uint32_t)((a >> 1) & 1));
        const float2 *srow = stage + (size_t)(a & 1) * (BT * L::kStageRow) + row * L::kStageRow;

        // pass 1: R2 butterflies of radix R1 over stride R2 (Doppler window already applied by K1)
#pragma unroll 1
        for (int u = slot; u < R2; u += NSLOT) {
            const int n2 = u;
            float2 x[R1];
#pragma unroll
            for (int m = 0; m < R1; ++m) {
                const int n = n2 + m * R2;
                x[m] = n < C ? srow[n] : make_float2(0.f, 0.f);
            }
            dft_regs<R1>(x);
#pragma unroll
            for (int k1 = 0; k1 < R1; ++k1) {
                float2 v = x[bitrev(k1, LR1)];
                if (k1 > 0) v = cmul(v, tw[n2 * k1]);
                wrow[k1 * R2 + n2] = v;
            }
        }
        __syncthreads();

        // pass 2: radix R2 on contiguous runs; accumulate |X|^2 (ascending antenna order)
#pragma unroll
        for (int ui = 0; ui < UPS2; ++ui) {
            const int k1 = slot + ui * NSLOT;
            if (k1 < R1) {
                float2 y[R2];
#pragma unroll
                for (int n2 = 0; n2 < R2; ++n2) y[n2] = wrow[k1 * R2 + n2];
                dft_regs<R2>(y);
#pragma unroll
                for (int k2 = 0; k2 < R2; ++k2) {
                    const float2 v = y[bitrev(k2, LR2)];
                    acc[ui][k2] += v.x * v.x + v.y * v.y;
                }
                if (cube != nullptr) {
                    float2 *o = cube + ((size_t)f * A + a) * (size_t)N * Sp + r0 + row;
#pragma unroll
                    for (int k2 = 0; k2 < R2; ++k2) st_global_f2(o + (size_t)(k1 + R1 * k2) * Sp, y[bitrev(k2, LR2)]);
                }
            }
        }
        __syncthreads();
    }

    float *po = pmap + (size_t)f * N * Sp + r0 + row;
#pragma unroll
    for (int ui = 0; ui < UPS2; ++ui) {
        const int k1 = slot + ui * NSLOT;
        if (k1 < R1) {
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) po[(size_t)(k1 + R1 * k2) * Sp] = acc[ui][k2];
        }
    }
}

// ---------------------------------------------------------------------------
// K3: 2-D CA-CFAR.  Range axis clamps (training count recounted), Doppler axis wraps.
// The noise estimate is a sum over the training cells themselves, evaluated in one fixed
// order shared with K4 (cfar_row_parts + ascending Doppler offsets), so both kernels get
// bit-identical thresholds.  Plain running sums / prefix differences are not usable here:
// a target cell is up to ~1e9 x the noise floor, and fp32 cancellation would leave an error
// of tens of noise floors behind it.
// ---------------------------------------------------------------------------
template <class Get>
__device__ __forceinline__ void cfar_row_parts(Get get, int guard, int half, float &full, float &ring)
{
    float left = 0.f, right = 0.f, mid = 0.f;
    for (int i = -half; i < -guard; ++i) left += get(i);
    for (int i = guard + 1; i <= half; ++i) right += get(i);
    for (int i = -guard; i <= guard; ++i) mid += get(i);
    ring = left + right;
    full = ring + mid;
}

__device__ __forceinline__ int cfar_train_count(const PlanDev &p, int r)
{
    const int n_full = min(r + p.win_r_half, p.Sp - 1) - max(r - p.win_r_half, 0) + 1;
    const int n_guard = min(r + p.guard_r, p.Sp - 1) - max(r - p.guard_r, 0) + 1;
    return (2 * p.win_d_half + 1) * n_full - (2 * p.guard_d + 1) * n_guard;
}

constexpr int kCfarRT = 64;      // range bins per CFAR tile
constexpr int kCfarNT = 256;

__global__ void __launch_bounds__(kCfarNT) cfar_kernel(PlanDev p, const float *__restrict__ pmap, uint32_t *__restrict__ mask)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int Wr = p.win_r_half, Wd = p.win_d_half, Gr = p.guard_r, Gd = p.guard_d;
    const int Sp = p.Sp, Cp = p.Cp;
    const int tw_ = kCfarRT + 2 * Wr;          // tile width
    const int th_ = 32 + 2 * Wd;               // tile height
    float *tile = reinterpret_cast<float *>(smem);               // [th_][tw_]
    float *fullS = tile + th_ * tw_;                              // [th_][RT+1]
    float *ringS = fullS + th_ * (kCfarRT + 1);                   // [th_][RT+1]
    uint32_t *words = reinterpret_cast<uint32_t *>(ringS + th_ * (kCfarRT + 1));   // [RT]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.x * kCfarRT;
    const int dblk = blockIdx.y;
    const int f = blockIdx.z;
    const int d0 = dblk * 32;
    const float *pf = pmap + (size_t)f * Cp * Sp;

    for (int i = tid; i < th_ * tw_; i += kCfarNT) {
        const int j = i / tw_, x = i - j * tw_;
        const int d = (d0 - Wd + j + Cp) & (Cp - 1);
        const int r = r0 - Wr + x;
        tile[i] = (r >= 0 && r < Sp) ? pf[(size_t)d * Sp + r] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < th_ * kCfarRT; i += kCfarNT) {
        const int j = i / kCfarRT, x = i - j * kCfarRT;
        const float *c = tile + j * tw_ + x + Wr;
        float full, ring;
        cfar_row_parts([&](int o) { return c[o]; }, Gr, Wr, full, ring);
        fullS[j * (kCfarRT + 1) + x] = full;
        ringS[j * (kCfarRT + 1) + x] = ring;
    }
    __syncthreads();
    for (int x = warp; x < kCfarRT; x += kCfarNT / 32) {
        const int r = r0 + x;
        float T = 0.f;
        for (int jj = -Wd; jj <= Wd; ++jj) {
            const int j = lane + Wd + jj;
            const bool g = (jj >= -Gd && jj <= Gd);
            T += g ? ringS[j * (kCfarRT + 1) + x] : fullS[j * (kCfarRT + 1) + x];
        }
        const int n = cfar_train_count(p, r);
        const float cut = tile[(lane + Wd) * tw_ + x + Wr];
        const bool hit = (r < Sp) && (n > 0) && (cut > p.alpha * (T / (float)n));
        const uint32_t w = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) words[x] = w;
    }
    __syncthreads();
    if (tid < kCfarRT && r0 + tid < Sp) mask[((size_t)f * (Cp / 32) + dblk) * Sp + r0 + tid] = words[tid];
}

// ---------------------------------------------------------------------------
// K4: detection records.  One CTA per frame.
// ---------------------------------------------------------------------------
constexpr int kDetNT = 256;
constexpr int kDetMaxA = 256;

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kDetNT) detect_kernel(PlanDev p, const float2 *__restrict__ rs, const float2 *__restrict__ cube,
                                                        const float *__restrict__ pmap, const uint32_t *__restrict__ mask,
                                                        mmw_detection *__restrict__ dets, uint32_t *__restrict__ counts)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *keys = reinterpret_cast<uint32_t *>(smem);                         // [max_det]
    float2 *xs = reinterpret_cast<float2 *>(keys + ((p.max_det + 3) & ~3));                  // [NW][A]
    float2 *twa = xs + (kDetNT / 32) * p.A;                                      // [n_theta]
    __shared__ uint32_t scan[kDetNT / 32];
    __shared__ uint32_t total_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int f = blockIdx.x;
    const int Sp = p.Sp, Cp = p.Cp, A = p.A, C = p.C;
    const int wpr = Cp / 32;                                  // mask words per range bin
    const int nwords = wpr * Sp;
    const uint32_t *mf = mask + (size_t)f * nwords;
    const float *pf = pmap + (size_t)f * Cp * Sp;

    for (int i = tid; i < p.n_theta; i += kDetNT) twa[i] = p.tw_a[i];

    // ---- phase 1: ordered compaction of set bits, order = (range, doppler) ----
    const int wpt = (nwords + kDetNT - 1) / kDetNT;
    const int i0 = tid * wpt, i1 = min(nwords, i0 + wpt);
    uint32_t cnt = 0;
    for (int i = i0; i < i1; ++i) cnt += __popc(mf[(size_t)(i % wpr) * Sp + i / wpr]);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) scan[warp] = incl;
    __syncthreads();
    if (tid == 0) {
        uint32_t s = 0;
        for (int w = 0; w < kDetNT / 32; ++w) {
            const uint32_t v = scan[w];
            scan[w] = s;
            s += v;
        }
        total_s = s;
    }
    __syncthreads();
    uint32_t pos = scan[warp] + incl - cnt;
    for (int i = i0; i < i1 && pos < (uint32_t)p.max_det; ++i) {
        uint32_t w = mf[(size_t)(i % wpr) * Sp + i / wpr];
        const uint32_t r = i / wpr, dbase = (i % wpr) * 32;
        while (w && pos < (uint32_t)p.max_det) {
            const int b = __ffs(w) - 1;
            w &= w - 1;
            keys[pos++] = (r << 16) | (dbase + b);
        }
    }
    __syncthreads();
    const uint32_t total = total_s;
    if (tid == 0) counts[f] = total;
    const int ndet = (int)min(total, (uint32_t)p.max_det);

    // ---- phase 2: one warp per detection ----
    const int Wd = p.win_d_half, Gd = p.guard_d;
    float2 *xw = xs + warp * A;
    for (int i = warp; i < ndet; i += kDetNT / 32) {
        const uint32_t key = keys[i];
        const int r = key >> 16, d = key & 0xffff;

        // noise: same association order as cfar_kernel
        float T = 0.f;
        for (int jb = -Wd; jb <= Wd; jb += 32) {
            const int jj = jb + lane;
            float val = 0.f;
            if (jj <= Wd) {
                const int dd = (d + jj + Cp) & (Cp - 1);
                const float *rowp = pf + (size_t)dd * Sp;
                float full, ring;
                cfar_row_parts([&](int o) { const int rr = r + o; return (rr >= 0 && rr < Sp) ? rowp[rr] : 0.f; }, p.guard_r,
                               p.win_r_half, full, ring);
                val = (jj >= -Gd && jj <= Gd) ? ring : full;
            }
            const int nv = min(32, Wd - jb + 1);
            for (int l = 0; l < nv; ++l) T += __shfl_sync(0xffffffffu, val, l);
        }
        const int n = cfar_train_count(p, r);
        const float noise = T / (float)n;
        const float pw = pf[(size_t)d * Sp + r];

        // 3x3 grouping among detected cells (Doppler wraps, range clamps; ties -> lowest (r,d))
        bool worse = false;
        if (lane < 9 && lane != 4) {
            const int rr = r + lane / 3 - 1;
            const int dd = (d + lane % 3 - 1 + Cp) & (Cp - 1);
            if (rr >= 0 && rr < Sp) {
                const uint32_t w = mf[(size_t)(dd >> 5) * Sp + rr];
                if ((w >> (dd & 31)) & 1u) {
                    const float pn = pf[(size_t)dd * Sp + rr];
                    const uint32_t kn = ((uint32_t)rr << 16) | (uint32_t)dd;
                    worse = (pn > pw) || (pn == pw && kn < key);
                }
            }
        }
        const bool is_peak = __ballot_sync(0xffffffffu, worse) == 0u;

        // antenna snapshot at (r, d)
        if (cube != nullptr) {
            for (int a = lane; a < A; a += 32) xw[a] = cube[(((size_t)f * A + a) * Cp + d) * Sp + r];
        } else {
            // fused mode: evaluate Doppler bin d of each antenna directly from the (windowed) range spectrum
            for (int a = 0; a < A; ++a) {
                const float2 *src = rs + (((size_t)f * A + a) * Sp + r) * (size_t)C;
                float sx = 0.f, sy = 0.f;
                for (int c = lane; c < C; c += 32) {
                    const float2 v = src[c];
                    const float2 w = p.tw_d[(c * d) & (Cp - 1)];
                    sx += v.x * w.x - v.y * w.y;
                    sy += v.x * w.y + v.y * w.x;
                }
                sx = warp_sum(sx);
                sy = warp_sum(sy);
                if (lane == 0) xw[a] = make_float2(sx, sy);
            }
        }
        __syncwarp();

        // angle spectrum arg-max (strict >, first wins)
        float best = -1.f;
        int bestk = 0;
        for (int k = lane; k < p.n_theta; k += 32) {
            float yx = 0.f, yy = 0.f;
            for (int a = 0; a < A; ++a) {
                const float2 v = xw[a];
                const float2 w = twa[(k * a) & (p.n_theta - 1)];
                yx += v.x * w.x - v.y * w.y;
                yy += v.x * w.y + v.y * w.x;
            }
            const float m = yx * yx + yy * yy;
            if (m > best) { best = m; bestk = k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bestk, o);
            if (ob > best || (ob == best && ok < bestk)) { best = ob; bestk = ok; }
        }
        __syncwarp();
        if (lane == 0) {
            const int kw = bestk < p.n_theta / 2 ? bestk : bestk - p.n_theta;
            float s = (float)kw * p.lambda_over_d / (float)p.n_theta;
            s = fminf(1.f, fmaxf(-1.f, s));
            mmw_detection o;
            o.frame = (uint32_t)f + p.frame_offset;
            o.range_bin = (uint16_t)r;
            o.doppler_bin = (uint16_t)d;
            o.power = pw;
            o.noise = noise;
            o.angle_bin = (int16_t)kw;
            o.flags = is_peak ? MMW_FLAG_PEAK : 0;
            o.angle_rad = asinf(s);
            dets[(size_t)f * p.max_det + i] = o;
        }
    }
}

// ---------------------------------------------------------------------------
// K5: per-frame lists -> dense ordered list + header {n_written, n_total, n_frames, overflow}
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) compact_kernel(const mmw_detection *__restrict__ dets, const uint32_t *__restrict__ counts,
                                                      mmw_detection *__restrict__ dense, uint32_t *__restrict__ header, int n_frames,
                                                      int max_det, int dense_cap)
{
    __shared__ uint32_t red[2][4];
    const int f = blockIdx.x, tid = threadIdx.x;
    // offset of this frame = sum of clipped counts of earlier frames; block 0 also sums everything
    const int upto = (f == 0) ? n_frames : f;
    uint32_t off = 0, tot = 0;
    for (int i = tid; i < upto; i += 128) {
        const uint32_t c = counts[i];
        off += min(c, (uint32_t)max_det);
        tot += c;
    }
    for (int o = 16; o > 0; o >>= 1) {
        off += __shfl_xor_sync(0xffffffffu, off, o);
        tot += __shfl_xor_sync(0xffffffffu, tot, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = off; red[1][tid >> 5] = tot; }
    __syncthreads();
    off = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    tot = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    if (f == 0) {
        if (tid == 0) {
            header[0] = min(off, (uint32_t)dense_cap);
            header[1] = tot;
            header[2] = (uint32_t)n_frames;
            header[3] = (tot != off || off > (uint32_t)dense_cap) ? 1u : 0u;
        }
        off = 0;
    }
    const uint32_t n = min(counts[f], (uint32_t)max_det);
    // 24-byte records as three 8-byte words
    const uint2 *src = reinterpret_cast<const uint2 *>(dets + (size_t)f * max_det);
    uint2 *dst = reinterpret_cast<uint2 *>(dense);
    for (uint32_t i = tid; i < 3 * n; i += 128) {
        const uint32_t rec = off + i / 3;
        if (rec < (uint32_t)dense_cap) dst[(size_t)off * 3 + i] = src[i];
    }
}

// ---------------------------------------------------------------------------
// export kernels (not on the hot path)
// ---------------------------------------------------------------------------
__global__ void export_cube_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, int A, int Sp, int Cp)
{   // in [A][Cp][Sp] -> out [A][Sp][Cp]
    const size_t n = (size_t)A * Sp * Cp;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int d = i % Cp;
        const int r = (i / Cp) % Sp;
        const int a = i / ((size_t)Cp * Sp);
        out[i] = in[((size_t)a * Cp + d) * Sp + r];
    }
}
__global__ void export_pmap_kernel(const float *__restrict__ in, float *__restrict__ out, int Sp, int Cp)
{
    const int n = Sp * Cp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int d = i % Cp, r = i / Cp;
        out[i] = in[(size_t)d * Sp + r];
    }
}
__global__ void export_mask_kernel(const uint32_t *__restrict__ in, uint8_t *__restrict__ out, int Sp, int Cp)
{
    const int n = Sp * Cp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int d = i % Cp, r = i / Cp;
        out[i] = (uint8_t)((in[(size_t)(d >> 5) * Sp + r] >> (d & 31)) & 1u);
    }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
template <typename K>
static cudaError_t set_smem(K kernel, int bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

template <int N, int R1, int R2, int BT, int NW, bool PAIR>
static cudaError_t run_range(const PlanDev &p, const int16_t *adc, float2 *rs, int n_frames, cudaStream_t st)
{
    auto k = range_fft_kernel<N, R1, R2, BT, NW, PAIR>;
    constexpr int bytes = RangeSmem<N, BT>::kBytes;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = set_smem(k, bytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int nct = (p.C + BT - 1) / BT;
    const long long grid = (long long)n_frames * p.A * nct;
    k<<<(unsigned)grid, NW * 32, bytes, st>>>(p, adc, rs);
    return cudaGetLastError();
}

template <int N, int R1, int R2, int BT, int NW>
static cudaError_t run_doppler(const PlanDev &p, const float2 *rs, float2 *cube, float *pmap, int n_frames, cudaStream_t st)
{
    auto k = doppler_fft_kernel<N, R1, R2, BT, NW>;
    constexpr int bytes = DopplerSmem<N, BT>::kBytes;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = set_smem(k, bytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const long long grid = (long long)n_frames * (p.Sp / BT);
    k<<<(unsigned)grid, NW * 32, bytes, st>>>(p, rs, cube, pmap);
    return cudaGetLastError();
}

bool plan_supported(int Sp, int Cp, const char **why)
{
    auto ok = [](int n) { return n == 64 || n == 128 || n == 256 || n == 512 || n == 1024; };
    if (!ok(Sp)) { if (why) *why = "range FFT length (nextPow2(n_samples)) must be 64..1024"; return false; }
    if (!ok(Cp)) { if (why) *why = "Doppler FFT length (nextPow2(n_chirps)) must be 64..1024"; return false; }
    return true;
}

cudaError_t launch_range_fft(const PlanDev &p, const int16_t *adc, float2 *rs, int n_frames, cudaStream_t st)
{
    switch (p.Sp) {
    case 64:   return run_range<64, 8, 8, 16, 4, true>(p, adc, rs, n_frames, st);
    case 128:  return run_range<128, 8, 16, 16, 4, true>(p, adc, rs, n_frames, st);
    case 256:  return run_range<256, 16, 16, 16, 4, true>(p, adc, rs, n_frames, st);
    case 512:  return run_range<512, 16, 32, 16, 8, true>(p, adc, rs, n_frames, st);
    case 1024: return run_range<1024, 32, 32, 16, 8, false>(p, adc, rs, n_frames, st);
    default:   return cudaErrorInvalidValue;
    }
}

cudaError_t launch_doppler_fft(const PlanDev &p, const float2 *rs, float2 *cube, float *pmap, int n_frames, cudaStream_t st)
{
    switch (p.Cp) {
    case 64:   return run_doppler<64, 8, 8, 16, 4>(p, rs, cube, pmap, n_frames, st);
    case 128:  return run_doppler<128, 8, 16, 16, 4>(p, rs, cube, pmap, n_frames, st);
    case 256:  return run_doppler<256, 16, 16, 16, 8>(p, rs, cube, pmap, n_frames, st);
    case 512:  return run_doppler<512, 16, 32, 16, 8>(p, rs, cube, pmap, n_frames, st);
    case 1024: return run_doppler<1024, 32, 32, 8, 8>(p, rs, cube, pmap, n_frames, st);
    default:   return cudaErrorInvalidValue;
    }
}

static int cfar_smem_bytes(const PlanDev &p)
{
    const int tw_ = kCfarRT + 2 * p.win_r_half, th_ = 32 + 2 * p.win_d_half;
    return (th_ * tw_ + 2 * th_ * (kCfarRT + 1)) * 4 + kCfarRT * 4;
}

cudaError_t launch_cfar(const PlanDev &p, const float *pmap, uint32_t *mask, int n_frames, cudaStream_t st)
{
    const int bytes = cfar_smem_bytes(p);
    static int configured = 0;
    if (bytes > configured) {
        cudaError_t e = set_smem(cfar_kernel, bytes);
        if (e != cudaSuccess) return e;
        configured = bytes;
    }
    dim3 grid((p.Sp + kCfarRT - 1) / kCfarRT, p.Cp / 32, n_frames);
    cfar_kernel<<<grid, kCfarNT, bytes, st>>>(p, pmap, mask);
    return cudaGetLastError();
}

cudaError_t launch_detect(const PlanDev &p, const float2 *rs, const float2 *cube, const float *pmap, const uint32_t *mask,
                          mmw_detection *dets, uint32_t *counts, int n_frames, cudaStream_t st)
{
    const int bytes = ((p.max_det + 3) & ~3) * 4 + (kDetNT / 32) * p.A * 8 + p.n_theta * 8;
    static int configured = 0;
    if (bytes > configured) {
        cudaError_t e = set_smem(detect_kernel, bytes);
        if (e != cudaSuccess) return e;
        configured = bytes;
    }
    detect_kernel<<<n_frames, kDetNT, bytes, st>>>(p, rs, p.keep_cube ? cube : nullptr, pmap, mask, dets, counts);
    return cudaGetLastError();
}

cudaError_t launch_compact(const PlanDev &p, const mmw_detection *dets, const uint32_t *counts, mmw_detection *dense,
                           uint32_t *header, int n_frames, int dense_cap, cudaStream_t st)
{
    compact_kernel<<<n_frames, 128, 0, st>>>(dets, counts, dense, header, n_frames, p.max_det, dense_cap);
    return cudaGetLastError();
}

cudaError_t launch_export_cube(const PlanDev &p, const float2 *cube_frame, float2 *out, cudaStream_t st)
{
    export_cube_kernel<<<592, 256, 0, st>>>(cube_frame, out, p.A, p.Sp, p.Cp);
    return cudaGetLastError();
}
cudaError_t launch_export_pmap(const PlanDev &p, const float *pmap_frame, float *out, cudaStream_t st)
{
    export_pmap_kernel<<<296, 256, 0, st>>>(pmap_frame, out, p.Sp, p.Cp);
    return cudaGetLastError();
}
cudaError_t launch_export_mask(const PlanDev &p, const uint32_t *mask_frame, uint8_t *out, cudaStream_t st)
{
    export_mask_kernel<<<296, 256, 0, st>>>(mask_frame, out, p.Sp, p.Cp);
    return cudaGetLastError();
}

}  // namespace mmw
