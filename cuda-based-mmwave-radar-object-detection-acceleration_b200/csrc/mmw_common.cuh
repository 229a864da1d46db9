// mmw_common.cuh — shared declarations of the B200 radar pipeline (device side).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mmw_radar.h"

namespace mmw {

// ---------------------------------------------------------------------------
// device-side view of one plan (all pointers are device pointers)
// ---------------------------------------------------------------------------
struct PlanDev {
    int S, C, A;            // samples / chirps / antennas as captured
    int Sp, Cp;             // FFT lengths (nextPow2)
    int n_theta;            // angle FFT length
    int guard_r, guard_d, win_r_half, win_d_half;   // CFAR: guard and guard+train half widths
    float alpha;
    float lambda_over_d;
    int max_det;            // per-frame detection capacity
    int keep_cube;          // materialise the Doppler cube
    uint32_t frame_offset;  // added to mmw_detection.frame
    const float *win_r;     // [S]   range window
    const float *win_d;     // [C]   Doppler window
    const float2 *tw_d;     // [Cp]  exp(-2 pi i k / Cp), natural order (fused-mode Doppler bin evaluation)
    const float2 *tw1_r;    // [Sp]  pass-1 twiddles of the range FFT in consumption order (tw1_index)
    const float2 *tw1_d;    // [Cp]  pass-1 twiddles of the Doppler FFT in consumption order
    const float2 *tw_a;     // [n_theta]
    unsigned int *sched;    // [8] work counters of the persistent FFT kernels: {next item, CTAs retired} for K1 at [0], K2 at [2]; zero between launches
    const int16_t *base_adc; // one frame [C][A][S] IIQQ subtracted before the range window, or nullptr (static-clutter removal)
    // host-side only: kernel-shape overrides for the sweeps under profiles/ and the kernel-form parity tests.  Read ONCE, in
    // mmw_create (MMW_K1_VARIANT / MMW_K2_VARIANT / MMW_K3_VARIANT / MMW_K4_VARIANT / MMW_CTAS_PER_SM); 0 = pick by shape.
    int k1_variant, k2_variant, k3_variant, k4_variant, ctas_per_sm_cap;
    int sched_dynamic;      // bit 0: K1, bit 1: K2 take their tiles from PlanDev.sched instead of a fixed stride (MMW_SCHED; default 1)
    int reserve_ctas;       // CTA slots the persistent FFT kernels leave free for kernels of other streams (mmw_reserve_ctas; MMW_RESERVE_CTAS)
    // MMW_FRONT: 0 = pick (fused front where supported), 1 = K1 and K2 as two kernels, 2 = fused front wherever supported;
    // MMW_FRONT_WINDOW: slabs the producer role may run ahead of the consumer role (0 = derived from the grid)
    int front_variant, front_window;
    unsigned long long *front_stats;   // MMW_FRONT_STATS=1: per-CTA timing record of the fused front kernel (mmw_front_stats), else nullptr
};

// internal HBM layouts
//   adc   [F][C][A][S]  int16 IIQQ   (the reference's capture format)
//   rs    [F][A][Sp][C] float2       range spectrum, already multiplied by win_d[c]
//   cube  [F][A][Cp][Sp] float2      Doppler cube (only if keep_cube)
//   pmap  [F][Cp][Sp]   float        integrated power, range fastest
//   mask  [F][Cp/32][Sp] uint32      CFAR hits, bit j of word (w, r) <-> doppler 32 w + j
//   keys  [F][max_det]  uint32 (range<<16 | doppler), counts [F], offsets [F+1]
//   dense [<= F*max_det] mmw_detection, ordered by (frame, range, doppler); header[4]

// ---------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk tensor-memory-accelerator copies (UBLKCP)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy; completion is signalled on `bar` as transaction bytes.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void st_global_f2(float2 *p, float2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// ---------------------------------------------------------------------------
// host-side launchers (mmw_pipeline.cu) — return cudaError_t
// ---------------------------------------------------------------------------
cudaError_t launch_range_fft(const PlanDev &p, const int16_t *adc, float2 *rs, int n_frames, cudaStream_t st);
cudaError_t launch_doppler_fft(const PlanDev &p, const float2 *rs, float2 *cube, float *pmap, int n_frames, cudaStream_t st);
// K1 + K2 as one cooperative kernel whose Doppler role reads the range spectrum back out of the L2 (mmw_front.cuh);
// sync: 2 * n_frames * A counters (zeroed by the launcher)
bool front_fused_supported(const PlanDev &p, int n_frames);
cudaError_t launch_front_fused(const PlanDev &p, const int16_t *adc, float2 *rs, float *pmap, int n_frames, unsigned int *sync, cudaStream_t st);
// small batches: K2 over F*A single-antenna frames into per-antenna maps, then the ordered sum (mmw_pipeline.cu)
bool doppler_prefers_split(const PlanDev &p, int n_frames);
cudaError_t launch_power_sum(const PlanDev &p, const float *per_antenna, float *pmap, int n_frames, cudaStream_t st);
// everything stages 3/4 read and write (device pointers)
struct DetectBuffers {
    const float2 *rs;
    const float2 *cube;
    const float *pmap;
    float *noise_map;       // [F][Cp][Sp], valid at hit cells only
    uint32_t *mask;         // [F][Cp/32][Sp]
    uint32_t *keys;         // [F][max_det]  (range << 16 | doppler), ordered
    uint32_t *counts;       // [F]    true hit count per frame
    uint32_t *offsets;      // [F+1]  exclusive scan of min(count, max_det)
    uint32_t *header;       // {n_written, n_total, n_frames, overflow}
    unsigned int *ticket;   // [0] last-CTA-done counter of list_kernel, [1] work cursor of measure_kernel, [2] hit-row count (all self-resetting)
    mmw_detection *dense;   // ordered detection list of the batch
    // wide arrays in fused mode (A >= 32, no Doppler cube): the selective Doppler re-FFT path (nullptr: measure_wide_kernel instead)
    uint4 *rows;            // [F * Sp] {frame, range bin, dense index of the row's first hit, -}: every (frame, range bin) with hits
    float2 *snap;           // [dense capacity][A] antenna snapshots of the detected cells
};
cudaError_t launch_cfar(const PlanDev &p, const float *pmap, uint32_t *mask, float *noise_map, int n_frames, int sm_count, cudaStream_t st);
cudaError_t launch_detect(const PlanDev &p, const DetectBuffers &b, int n_frames, int dense_cap, int sm_count, cudaStream_t st);
// the kernels that write mmw_detection records (the only readers of PlanDev.frame_offset, their argument 0): graph mode patches
// that argument in place instead of re-capturing when the frame offset changes (mmw_api.cu).  Returns the kernel's argument
// count, 0 if `func` is not one of them.
constexpr int kRecordKernelMaxArgs = 12;
int record_kernel_args(const void *func);
cudaError_t launch_doppler_extract(const PlanDev &p, const float2 *rs, const uint32_t *keys, const uint32_t *offsets, const uint4 *rows,
                                   const unsigned int *n_rows, float2 *snap, int dense_cap, int max_rows, cudaStream_t st);
cudaError_t launch_merge(const unsigned char *gathered, int n_ranks, size_t stride_bytes, unsigned char *merged, int merged_cap,
                         cudaStream_t st);
// export helpers (not on the hot path): internal layout -> the canonical layouts of mmw_radar.h
cudaError_t launch_export_cube(const PlanDev &p, const float2 *cube_frame, float2 *out, cudaStream_t st);
cudaError_t launch_export_pmap(const PlanDev &p, const float *pmap_frame, float *out, cudaStream_t st);
cudaError_t launch_export_mask(const PlanDev &p, const uint32_t *mask_frame, uint8_t *out, cudaStream_t st);
bool plan_supported(int Sp, int Cp, const char **why);
constexpr int kMaxDevices = 64;
int current_device();       // ordinal of the calling thread's current device, clamped to [0, kMaxDevices)
void plan_radices(int n, int *r1, int *r2);

// legacy single-frame path (mmw_legacy.cu)
struct LegacyDev {
    const float2 *tw;       // [16384]
    const double2 *base;    // [12800] base frame rx0 (as handed in by the caller)
    float2 *spectrum;       // [16384] optional export
    unsigned long long *best;  // packed (|X|^2 bits, ~index)
};

}  // namespace mmw
