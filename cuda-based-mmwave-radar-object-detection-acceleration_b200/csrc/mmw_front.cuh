// mmw_front.cuh — stages 1 and 2 as the two ROLES of one persistent kernel, so that the range spectrum is read back
// out of the L2 instead of out of HBM.  (Included by mmw_pipeline.cu after RangeSmem / DopplerWarp.)
//
// The two-kernel chain writes the whole corner-turned range spectrum of a batch (cfg3: 805 MB) and only then reads it
// back: 16 of the 20 bytes per sample the FFT stages move are that round trip, and the batch is 6x the 126 MB L2.  Here a
// producer role (the range FFT of range_fft_kernel, same code, same roundings) and a consumer role (the warp-private
// Doppler FFT of doppler_fft_warp_kernel) run side by side, one CTA of each per SM, and the consumer trails the producer
// by a few slabs — a slab is one (frame, antenna) plane [Sp][C] of the range spectrum, 1 MB on cfg3 — so its bulk copies
// hit lines that are still in the L2.  HBM then sees the int16 capture once and the write-back of the range spectrum
// once (the detection stage still reads hit rows from it), 12 bytes per sample instead of 20.
//
// Order.  A consumer warp owns one tile (ROWS range bins of one frame) for all A antennas and keeps the |X|^2 sums in
// registers, so the consumer role as a whole works on FG = (consumer warps) / (tiles per frame) frames at a time, walking
// their antennas roughly in step.  The producer therefore emits slabs group by group: for frames [g FG, (g+1) FG): for
// antenna a: for frame f of the group: slab (f, a).  What is live between the roles is then a few slabs, not a few
// frames.
//
// Hand-over.  produced[f A + a] counts the producer tiles of slab (f, a) that are complete in global memory (CTA barrier,
// then one thread: fence + atomic add); a consumer lane polls it with ld.acquire before it issues the bulk copy of its
// rows (plus fence.proxy.async: the copy reads through the async proxy what other SMs wrote through the generic one).
// consumed[f A + a] counts the consumer warps whose copy of that slab has landed; a producer CTA does not store into
// slab sigma before slab sigma - window is consumed, which keeps the producer from running away from the L2 (a stalled
// producer CTA sleeps, and its issue slots go to the consumer CTA on the same SM).  Producer waits only on OLDER slabs and
// a consumer warp only on slabs of its own frame, in order, so with every CTA resident (cooperative launch) there is no
// cycle.  A poll that lasts two seconds traps instead of hanging the GPU.
#pragma once

namespace mmw {

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *p >= want; traps after ~2 s (a hand-over that never comes must not hang the device)
__device__ __forceinline__ unsigned long long wait_counter(const unsigned int *p, unsigned int want)
{
    if (ld_acquire_u32(p) >= want) return 0ull;
    const unsigned long long t0 = global_timer_ns();
    unsigned int spins = 0;
    while (ld_acquire_u32(p) < want) {
        __nanosleep(100);
        if ((++spins & 1023u) == 0u && global_timer_ns() - t0 > 2000000000ull) __trap();
    }
    return global_timer_ns() - t0;                                   // time spent waiting (instrumented runs only read it)
}
__device__ __forceinline__ void cp_async_16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// L2 eviction-priority hint for streaming reads (the capture is read once): keeps it from displacing the range spectrum
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

// slab sequence number -> (frame, antenna): groups of FG frames, antenna-major inside a group
struct SlabOrder {
    int F, A, FG;
    __device__ __forceinline__ void at(int sigma, int &f, int &a) const
    {
        const int per_group = FG * A;
        const int g = sigma / per_group, rem = sigma - g * per_group;
        const int nf = min(FG, F - g * FG);
        a = rem / nf;
        f = g * FG + (rem - a * nf);
    }
};

struct FrontArgs {
    const int16_t *adc;
    float2 *rs;
    float *pmap;
    unsigned int *produced;     // [F * A], zero at launch
    unsigned int *consumed;     // [F * A], zero at launch
    int n_frames;
    int FG;                     // frames per consumer group
    int n_consumer_ctas;        // blockIdx < this: Doppler role
    int window;                 // producer may be this many slabs ahead of the consumer
    unsigned long long *stats;  // optional (profiles/front_probe.py): per CTA {smid | role << 32, start ns, end ns, ns waited, units done, -, -, -}
};

template <int SN, int BT, int DN, int DR1, int DR2, int NW>
struct FrontSmem {
    static constexpr int kDStage = 2;
    static constexpr int kRange = RangeSmem<SN, BT, 1, false>::kBytes;
    static constexpr int kDoppler = DopplerWarp<DN, DR1, DR2>::bytes(NW, kDStage);
    static constexpr int kBytes = kRange > kDoppler ? kRange : kDoppler;
};

// SN/SR1/SR2: range FFT plan (pair form: SN <= 512); BT: chirps per producer tile; PADS: n_samples < SN.
// DN/DR1/DR2: Doppler FFT plan; PADC: n_chirps < DN.  NW warps per CTA in both roles.
template <int SN, int SR1, int SR2, int BT, bool PADS, int DN, int DR1, int DR2, bool PADC, int NW>
__global__ void __launch_bounds__(NW * 32, 512 / (NW * 32)) front_fused_kernel(PlanDev p, FrontArgs g)
{
    static_assert(SR1 * SR2 == SN && DR1 * DR2 == DN, "plan");
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int A = p.A;
    const int C = PADC ? p.C : DN;
    const SlabOrder order{g.n_frames, A, g.FG};

    if ((int)blockIdx.x >= g.n_consumer_ctas) {
        // =================================================================================================
        // producer role: range FFT (range_fft_kernel's pair form, single staging buffer)
        // =================================================================================================
        using L = RangeSmem<SN, BT, 1, false>;
        constexpr int NT = NW * 32;
        constexpr int SUBS = 32 / BT;
        constexpr int NSLOT = NW * SUBS;
        constexpr int LR1 = ilog2(SR1), LR2 = ilog2(SR2);
        uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
        float2 *tw = reinterpret_cast<float2 *>(smem + L::kOffTw);
        float *win = reinterpret_cast<float *>(smem + L::kOffWin);
        unsigned char *stage = smem + L::kOffStage;
        float2 *work = reinterpret_cast<float2 *>(smem + L::kOffWork);

        const int row = lane % BT, sub = lane / BT, slot = warp * SUBS + sub;
        const int S = PADS ? p.S : SN;
        const int nct = (C + BT - 1) / BT;
        const int n_tiles = g.n_frames * A * nct;
        const int rank = (int)blockIdx.x - g.n_consumer_ctas, nprod = (int)gridDim.x - g.n_consumer_ctas;
        const int n_cons_per_slab = p.Sp / DopplerWarp<DN, DR1, DR2>::kRows;
        const uint64_t pol = l2_policy_evict_first();

        auto issue = [&](int tile) {                                  // warp 0: stage the BT int16 rows of `tile`
            const int ct = tile % nct;
            int f, a;
            order.at(tile / nct, f, a);
            const int c0 = ct * BT;
            const int nrows = min(BT, C - c0);
            if (lane == 0) {
                fence_proxy_async();
                mbar_arrive_expect_tx(&bar[0], (uint32_t)(nrows * S * 4));
            }
            __syncwarp();
            if (lane < nrows) {
                const int16_t *src = g.adc + (((size_t)f * C + c0 + lane) * A + a) * (size_t)(2 * S);
                bulk_g2s_hint(stage + lane * L::kStageStride, src, (uint32_t)(S * 4), &bar[0], pol);
            }
        };

        if (tid == 0) {
            mbar_init(&bar[0], 1);
            fence_mbar_init();
        }
        __syncthreads();
        int tile = rank;
        if (warp == 0 && tile < n_tiles) issue(tile);
        for (int i = tid; i < SN; i += NT) tw[i] = p.tw1_r[i];
        for (int i = tid; i < SN; i += NT) win[i] = i < S ? p.win_r[i] : 0.f;
        __syncthreads();

        float2 *wrow = work + row * (SN + 1);
        int it = 0;
        int publish_fa = -1;                                          // slab of the tile finished last, not yet published
        unsigned long long waited = 0ull;
        const unsigned long long t_start = g.stats ? global_timer_ns() : 0ull;
#pragma unroll 1
        for (; tile < n_tiles; tile += nprod, ++it) {
            const int sigma = tile / nct, ct = tile - sigma * nct;
            int f, a;
            order.at(sigma, f, a);
            const int fa = f * A + a;
            const int c0 = ct * BT;
            const bool row_ok = c0 + row < C;
            const float wdop = row_ok ? p.win_d[c0 + row] : 0.f;
            // the throttle's counter (see below) is fetched now and looked at after pass 1: its latency is off the critical path
            unsigned int seen = 0xffffffffu;
            const unsigned int *throttle = nullptr;
            if (tid == 32 && sigma >= g.window) {
                int fo, ao;
                order.at(sigma - g.window, fo, ao);
                throttle = &g.consumed[fo * A + ao];
                seen = *reinterpret_cast<const volatile unsigned int *>(throttle);
            }
            mbar_wait(&bar[0], (uint32_t)(it & 1));
            const unsigned char *srow = stage + row * L::kStageStride;

            // ---- pass 1 ----
#pragma unroll 1
            for (int u = slot; u < SR2 / 2; u += NSLOT) {
                const int n2 = 2 * u;
                float2 xa[SR1], xb[SR1];
#pragma unroll
                for (int m = 0; m < SR1; ++m) {
                    const int n = n2 + m * SR2;
                    if (!PADS || n < S) {
                        const uint2 raw = *reinterpret_cast<const uint2 *>(srow + 4 * n);   // [I(n) I(n+1)] [Q(n) Q(n+1)]
                        const float2 w = *reinterpret_cast<const float2 *>(win + n);
                        const int i0 = (short)(raw.x & 0xffffu), q0 = (short)(raw.y & 0xffffu), i1 = (int)raw.x >> 16, q1 = (int)raw.y >> 16;
                        xa[m] = make_float2((float)i0 * w.x, (float)q0 * w.x);
                        xb[m] = make_float2((float)i1 * w.y, (float)q1 * w.y);
                    } else {
                        xa[m] = xb[m] = make_float2(0.f, 0.f);
                    }
                }
                dft_regs<SR1>(xa);
                dft_regs<SR1>(xb);
                const float4 *twu = reinterpret_cast<const float4 *>(tw) + u * SR1;
                float2 *wo = wrow + n2;
#pragma unroll
                for (int k1 = 0; k1 < SR1; ++k1) {
                    float2 va = xa[bitrev(k1, LR1)], vb = xb[bitrev(k1, LR1)];
                    if (k1 > 0) {
                        const float4 t = twu[k1];
                        va = cmul(va, make_float2(t.x, t.y));
                        vb = cmul(vb, make_float2(t.z, t.w));
                    }
                    wo[k1 * SR2] = va;
                    wo[k1 * SR2 + 1] = vb;
                }
            }
            // not more than `window` slabs ahead of the consumer (waits here, before the stores, cost nothing when the
            // roles are in balance: the barrier below is there anyway)
            // publish the PREVIOUS tile here, half a tile after its stores were issued (right after them the fence held this warp,
            // and with it the CTA's next barrier, for microseconds: profiles/r2/ncu_front_fused.md)
            if (tid == 0 && publish_fa >= 0) {
                __threadfence();
                atomicAdd(&g.produced[publish_fa], 1u);
            }
            // not more than `window` slabs ahead of the consumer role (a hint that keeps the range spectrum in the L2, not a
            // correctness condition: nothing is overwritten, so a relaxed read is enough); thread 32, so that it does not
            // queue behind thread 0's fence
            if (seen < (unsigned int)n_cons_per_slab) waited += wait_counter(throttle, (unsigned int)n_cons_per_slab);
            __syncthreads();
            if (warp == 0 && tile + nprod < n_tiles) issue(tile + nprod);   // staging buffer consumed: refill behind pass 2

            // ---- pass 2: outputs go straight to the corner-turned slab ----
            float2 *out = g.rs + (size_t)fa * (size_t)SN * C + c0 + row;
#pragma unroll 1
            for (int u = slot; u < SR1; u += NSLOT) {
                const int k1 = u;
                float2 y[SR2];
                const float2 *wi = wrow + k1 * SR2;
#pragma unroll
                for (int n2 = 0; n2 < SR2; ++n2) y[n2] = wi[n2];
                dft_regs<SR2>(y);
                if (row_ok) {
                    float2 *o = out + (size_t)k1 * C;
#pragma unroll
                    for (int k2 = 0; k2 < SR2; ++k2) {
                        const float2 v = y[bitrev(k2, LR2)];
                        st_global_f2(o + (size_t)(SR1 * k2) * C, cscale(v, wdop));
                    }
                }
            }
            __syncthreads();                                          // the tile's stores are issued (and ordered before tid 0's fence)
            publish_fa = fa;
        }
        if (tid == 0 && publish_fa >= 0) {
            __threadfence();
            atomicAdd(&g.produced[publish_fa], 1u);
        }
        if (g.stats && tid == 32) g.stats[(size_t)blockIdx.x * 8 + 3] = waited;
        if (g.stats && tid == 0) {
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            unsigned long long *o = g.stats + (size_t)blockIdx.x * 8;
            o[0] = smid | (1ull << 32); o[1] = t_start; o[2] = global_timer_ns(); o[4] = (unsigned long long)it;
        }
        return;
    }

    // =====================================================================================================
    // consumer role: Doppler FFT with warp-private tiles + |X|^2 over antennas (doppler_fft_warp_kernel)
    // =====================================================================================================
    using L = DopplerWarp<DN, DR1, DR2>;
    constexpr int NSTAGE = FrontSmem<SN, BT, DN, DR1, DR2, NW>::kDStage;
    constexpr int SUBS = L::kSubs, ROWS = L::kRows, U1 = L::kU1, U2 = L::kU2;
    constexpr int LR1 = ilog2(DR1), LR2 = ilog2(DR2);
    constexpr bool TWREG = U1 * (DR1 - 1) <= 16;

    float2 *tw = reinterpret_cast<float2 *>(smem + L::kOffTw);
    float2 *ring = reinterpret_cast<float2 *>(smem + L::kOffTw + 8 * DN) + (size_t)warp * NSTAGE * L::kStage;

    const int row = lane / SUBS, sub = lane % SUBS;
    const int Sp = p.Sp;
    const int nrt = Sp / ROWS;
    const int nct = (C + BT - 1) / BT;                               // producer tiles per slab
    const int gw = (int)blockIdx.x * NW + warp;
    const int fi = gw / nrt, rt = gw - fi * nrt;                      // this warp's frame within a group and its range tile
    const int n_groups = (g.n_frames + g.FG - 1) / g.FG;
    const bool active = fi < g.FG;

    // The rows come in by 16-byte cp.async (LDGSTS, L2 only), one commit group per step — NOT by bulk copies: a bulk copy reads
    // through the async proxy, and ordering it after the generic-proxy stores of the producer SMs takes a fence.proxy.async
    // (MEMBAR.ALL.GPU + FENCE.VIEW.ASYNC) in every step of every warp, a fifth of this role's time in the first version
    // (profiles/r2/ncu_front_fused_v1.md).  cp.async stays in the generic proxy: lane 0's ld.acquire + __syncwarp orders it.
    unsigned long long waited = 0ull;
    const unsigned long long t_start = g.stats ? global_timer_ns() : 0ull;
    auto issue_step = [&](int q) {                                    // the q-th step of this warp: group q / A, antenna q % A
        const int grp = q / A, a = q - grp * A;
        const int f = grp * g.FG + fi;
        if (grp < n_groups && f < g.n_frames) {
            if (lane == 0) waited += wait_counter(&g.produced[f * A + a], (unsigned int)nct);
            __syncwarp();
            // the tile's ROWS rows are contiguous in the slab: ROWS * C * 8 bytes, 16 per copy
            const float2 *src = g.rs + (((size_t)f * A + a) * Sp + rt * ROWS) * (size_t)C;
            float2 *dst = ring + (size_t)(q % NSTAGE) * L::kStage;
            const int half_c = C >> 1;
#pragma unroll
            for (int j0 = 0; j0 < ROWS * DN / 2; j0 += 32) {
                const int j = j0 + lane;
                const int r = j / half_c, col = j - r * half_c;
                if (!PADC || r < ROWS) cp_async_16(dst + r * L::kRowStride + 2 * col, src + 2 * j);
            }
        }
        cp_async_commit();                                            // (an empty group past the last step keeps the count in step)
    };

    for (int i = tid; i < DN; i += NW * 32) tw[i] = p.tw1_d[i];
    __syncthreads();                                                 // the only CTA-wide barrier of this role
    if (!active) return;
#pragma unroll
    for (int i = 0; i < NSTAGE - 1; ++i) issue_step(i);

    float2 twr[TWREG ? U1 : 1][TWREG ? DR1 - 1 : 1];
    if constexpr (TWREG) {
#pragma unroll
        for (int u = 0; u < U1; ++u)
#pragma unroll
            for (int k1 = 1; k1 < DR1; ++k1) twr[u][k1 - 1] = tw[tw1_index(sub + u * SUBS, k1, DR1)];
    }

    int q = 0;
#pragma unroll 1
    for (int grp = 0; grp < n_groups; ++grp) {
        const int f = grp * g.FG + fi;
        if (f >= g.n_frames) break;                                   // (only in the last, shorter group)
        float acc[U2][DR2];
#pragma unroll
        for (int u = 0; u < U2; ++u)
#pragma unroll
            for (int j = 0; j < DR2; ++j) acc[u][j] = 0.f;

#pragma unroll 1
        for (int a = 0; a < A; ++a, ++q) {
            issue_step(q + NSTAGE - 1);                              // into the buffer step q - 1 has just left
            cp_async_wait<NSTAGE - 1>();                             // this lane's copies of step q have landed ...
            __syncwarp();                                            // ... and so have the other lanes'
            if (lane == 0) atomicAdd(&g.consumed[f * A + a], 1u);    // this warp's rows of slab (f, a) are on chip (relaxed is enough)
            float2 *r = ring + (size_t)(q % NSTAGE) * L::kStage + row * L::kRowStride;

            float2 x[U1][DR1];
#pragma unroll
            for (int u = 0; u < U1; ++u) {
                const int n2 = sub + u * SUBS;
#pragma unroll
                for (int m = 0; m < DR1; ++m) {
                    const int n = n2 + m * DR2;
                    x[u][m] = (!PADC || n < C) ? r[n] : make_float2(0.f, 0.f);
                }
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < U1; ++u) {
                const int n2 = sub + u * SUBS;
                dft_regs<DR1>(x[u]);
                r[n2] = x[u][0];
#pragma unroll
                for (int k1 = 1; k1 < DR1; ++k1) {
                    const float2 w = TWREG ? twr[u][k1 - 1] : tw[tw1_index(n2, k1, DR1)];
                    r[k1 * (DR2 + 1) + n2] = cmul(x[u][bitrev(k1, LR1)], w);
                }
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < U2; ++u) {
                const int k1 = sub + u * SUBS;
                float2 y[DR2];
                const float2 *wi = r + k1 * (DR2 + 1);
#pragma unroll
                for (int n2 = 0; n2 < DR2; ++n2) y[n2] = wi[n2];
                dft_regs<DR2>(y);
#pragma unroll
                for (int k2 = 0; k2 < DR2; ++k2) {
                    const float2 v = y[bitrev(k2, LR2)];
                    acc[u][k2] = accumulate_power(acc[u][k2], v);
                }
            }
            __syncwarp();
        }

        float *po = g.pmap + (size_t)f * DN * Sp + rt * ROWS + row;
#pragma unroll
        for (int u = 0; u < U2; ++u) {
            const int k1 = sub + u * SUBS;
#pragma unroll
            for (int k2 = 0; k2 < DR2; ++k2) po[(size_t)(k1 + DR1 * k2) * Sp] = acc[u][k2];
        }
    }
    if (g.stats && tid == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long *o = g.stats + (size_t)blockIdx.x * 8;
        o[0] = smid; o[1] = t_start; o[2] = global_timer_ns(); o[3] = waited; o[4] = (unsigned long long)q;
    }
}

}  // namespace mmw
