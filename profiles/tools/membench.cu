// membench.cu — what HBM delivers on this B200 for the access patterns the radar kernels use:
// read-only, write-only and copy streams with 128-bit LDG/STG, and a read-only stream staged through
// shared memory by 1-D TMA bulk copies (the way K1/K2 load).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void k_read(const int4 *__restrict__ in, size_t n, int4 *sink)
{
    int4 acc = make_int4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + i));
        acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678) sink[0] = acc;
}
__global__ void k_write(int4 *__restrict__ out, size_t n)
{
    const int4 v = make_int4(1, 2, 3, 4);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}
__global__ void k_copy(const int4 *__restrict__ in, int4 *__restrict__ out, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
// K1's traffic mix: one byte read for two written.  Contiguous form: 16 B in, 2 x 16 B out (two coalesced streams) per thread and step.
__global__ void k_mix12(const int4 *__restrict__ in, int4 *__restrict__ out, size_t n_in)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_in; i += (size_t)gridDim.x * blockDim.x) {
        int4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + i));
        out[i] = v;                      // two fully coalesced output streams
        out[i + n_in] = v;
    }
}
// K1's store pattern: the corner-turned range spectrum [slab][range 512][chirp 256] float2 is written in pieces of 16 chirps
// (128 B) per range row, 2 KB apart; the 16 pieces of a row come from 16 different CTAs (consecutive tiles) at about the same
// time.  A CTA = one tile (slab, chirp block): reads its 32 KB of "ADC" contiguously, then writes 512 x 128 B at stride 2 KB
// with 8-byte stores (lanes 0-15 one row, lanes 16-31 the next), as the kernel does.
__global__ void k_mix12_k1pattern(const int4 *__restrict__ in, float2 *__restrict__ out, size_t n_tiles)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const size_t slab = t / 16, cb = t % 16;
        int4 acc = make_int4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) {            // 32 KB in
            int4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + t * 2048 + i));
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
        const float2 val = make_float2(__int_as_float(acc.x ^ acc.z), __int_as_float(acc.y ^ acc.w));
        float2 *o = out + slab * (512 * 256) + cb * 16 + (lane & 15);
        for (int r = 2 * warp + (lane >> 4); r < 512; r += 2 * nw)          // 64 KB out
            asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(o + (size_t)r * 256), "f"(val.x), "f"(val.y) : "memory");
    }
}
// the same traffic with wider pieces per range row: PIECE chirps (PIECE * 8 bytes) per row, i.e. what a tile of PIECE chirps would store
template <int PIECE>
__global__ void k_mix12_rows(const int4 *__restrict__ in, float2 *__restrict__ out, size_t n_tiles)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    constexpr int BLOCKS = 256 / PIECE;                                   // tiles per slab
    constexpr int IN16 = PIECE * 512 * 4 / 16;                            // int4 per tile (PIECE chirps x 512 samples x 4 B)
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const size_t slab = t / BLOCKS, cb = t % BLOCKS;
        int4 acc = make_int4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < IN16; i += blockDim.x) {
            int4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + t * IN16 + i));
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
        const float2 val = make_float2(__int_as_float(acc.x ^ acc.z), __int_as_float(acc.y ^ acc.w));
        constexpr int LPR = PIECE < 32 ? PIECE : 32;                      // lanes per row piece
        constexpr int RPW = 32 / LPR;                                     // rows per warp store
        float2 *o = out + slab * (512 * 256) + cb * PIECE + (lane % LPR);
        for (int r = RPW * warp + lane / LPR; r < 512; r += RPW * nw)
#pragma unroll
            for (int c = 0; c < PIECE; c += 32)
                asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(o + (size_t)r * 256 + c), "f"(val.x), "f"(val.y) : "memory");
    }
}
// 16-chirp tiles storing into a range spectrum tiled by 32 chirps ([slab][chirp/32][range 512][32]): 128-byte halves of adjacent
// 256-byte pieces, the other half coming from the neighbouring tile (another CTA)
__global__ void k_mix12_tiled32(const int4 *__restrict__ in, float2 *__restrict__ out, size_t n_tiles)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const size_t slab = t / 16, cb = t % 16;
        int4 acc = make_int4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
            int4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + t * 2048 + i));
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
        const float2 val = make_float2(__int_as_float(acc.x ^ acc.z), __int_as_float(acc.y ^ acc.w));
        float2 *o = out + slab * (512 * 256) + (cb >> 1) * (512 * 32) + (cb & 1) * 16 + (lane & 15);
        for (int r = 2 * warp + (lane >> 4); r < 512; r += 2 * nw)
            asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(o + (size_t)r * 32), "f"(val.x), "f"(val.y) : "memory");
    }
}
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// each CTA streams CHUNK-byte pieces through a 2-deep ring of shared-memory buffers with cp.async.bulk
template <int CHUNK>
__global__ void k_tma_read(const char *__restrict__ in, size_t nchunks, int *sink)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bar[2];
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[b])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    int acc = 0;
    size_t c = blockIdx.x;
    int s = 0;
    auto issue = [&](size_t chunk, int st) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[st & 1])), "r"(CHUNK));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + (st & 1) * CHUNK)),
                         "l"(in + chunk * CHUNK), "r"(CHUNK), "r"(s32(&bar[st & 1])));
        }
    };
    if (c < nchunks) issue(c, 0);
    for (; c < nchunks; c += gridDim.x, ++s) {
        if (c + gridDim.x < nchunks) issue(c + gridDim.x, s + 1);
        uint32_t done;
        do {
            asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(s32(&bar[s & 1])), "r"((s >> 1) & 1));
        } while (!done);
        acc ^= reinterpret_cast<const int *>(sm + (s & 1) * CHUNK)[threadIdx.x];
        __syncthreads();
    }
    if (acc == 0x12345678) sink[0] = acc;
}

// K2's loads if the range spectrum were stored tile-contiguously ([slab][chirp block of 16][range 512][16 chirps]): a range row of
// 256 chirps = 16 pieces of 128 B, 64 KB apart.  Warp-private staging as in doppler_fft_warp_kernel: a warp stages two rows
// (32 pieces, one bulk copy per lane) per step through a 2-deep ring.  PIECES = 1: today's layout (one 2 KB copy per row).
template <int PIECES>
__global__ void k_tma_rows(const char *__restrict__ in, size_t n_rows, int *sink)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bar[8][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) {
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[warp][b])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncwarp();
    unsigned char *mine = sm + warp * 2 * 4096;
    constexpr int PB = 2048 / PIECES;                                  // bytes per piece
    auto issue = [&](size_t pair, int st) {                              // rows 2 pair, 2 pair + 1
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[warp][st & 1])), "r"(4096));
        __syncwarp();
        for (int i = lane; i < 2 * PIECES; i += 32) {
            const size_t row = 2 * pair + i / PIECES, slab = row / 512, r = row % 512;
            const int piece = i % PIECES;
            // tiled layout: [slab][piece][r][PB bytes]; PIECES = 1 degenerates to [slab][r][2048]
            const char *src = in + slab * (512 * 2048) + (size_t)piece * (512 * PB) + r * PB;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(mine + (st & 1) * 4096 + i * PB)),
                         "l"(src), "r"(PB), "r"(s32(&bar[warp][st & 1])));
        }
    };
    const size_t n_pairs = n_rows / 2, stride = (size_t)gridDim.x * nw;
    size_t pair = (size_t)blockIdx.x * nw + warp;
    int acc = 0, s = 0;
    if (pair < n_pairs) issue(pair, 0);
    for (; pair < n_pairs; pair += stride, ++s) {
        if (pair + stride < n_pairs) issue(pair + stride, s + 1);
        uint32_t done;
        do {
            asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(s32(&bar[warp][s & 1])), "r"((s >> 1) & 1));
        } while (!done);
        acc ^= reinterpret_cast<const int *>(mine + (s & 1) * 4096)[lane * 32];
        __syncwarp();
    }
    if (acc == 0x12345678) sink[0] = acc;
}

int main()
{
    const size_t bytes = 2048ull << 20;
    char *a, *b;
    cudaMalloc(&a, bytes);
    cudaMalloc(&b, bytes);
    cudaMemset(a, 1, bytes);
    cudaMemset(b, 2, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const size_t n = bytes / 16;
    auto time = [&](const char *name, double moved, auto launch) {
        float best = 1e9;
        for (int it = 0; it < 6; ++it) {
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        printf("%-28s %8.3f ms  %8.1f GB/s   (%s)\n", name, best, moved / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
        printf("grid %d x 512\n", blocks);
        time("read  LDG.128", (double)bytes, [&] { k_read<<<blocks, 512>>>((const int4 *)a, n, (int4 *)b); });
        time("write STG.128", (double)bytes, [&] { k_write<<<blocks, 512>>>((int4 *)b, n); });
        time("copy  (read+write bytes)", 2.0 * bytes, [&] { k_copy<<<blocks, 512>>>((const int4 *)a, (int4 *)b, n); });
    }
    for (int blocks : {148 * 4, 148 * 8}) {
        char nm[64];
        snprintf(nm, sizeof nm, "1 read : 2 write, contiguous, %d", blocks);
        time(nm, 1.5 * bytes, [&] { k_mix12<<<blocks, 512>>>((const int4 *)a, (int4 *)b, n / 2); });
    }
    for (int per_sm : {2, 4, 8}) {
        char nm[64];
        snprintf(nm, sizeof nm, "1 : 2, K1 store pattern, %d/SM", per_sm);
        time(nm, 1.5 * bytes, [&] { k_mix12_k1pattern<<<148 * per_sm, 256>>>((const int4 *)a, (float2 *)b, bytes / 65536); });
    }
    for (int per_sm : {4, 8}) {
        char nm[64];
        snprintf(nm, sizeof nm, "1 : 2, 64 B pieces, %d/SM", per_sm);
        time(nm, 1.5 * bytes, [&] { k_mix12_rows<8><<<148 * per_sm, 256>>>((const int4 *)a, (float2 *)b, bytes / 32768); });
        snprintf(nm, sizeof nm, "1 : 2, 128 B pieces, %d/SM", per_sm);
        time(nm, 1.5 * bytes, [&] { k_mix12_rows<16><<<148 * per_sm, 256>>>((const int4 *)a, (float2 *)b, bytes / 65536); });
        snprintf(nm, sizeof nm, "1 : 2, 256 B pieces, %d/SM", per_sm);
        time(nm, 1.5 * bytes, [&] { k_mix12_rows<32><<<148 * per_sm, 256>>>((const int4 *)a, (float2 *)b, bytes / 131072); });
        snprintf(nm, sizeof nm, "1 : 2, 512 B pieces, %d/SM", per_sm);
        time(nm, 1.5 * bytes, [&] { k_mix12_rows<64><<<148 * per_sm, 256>>>((const int4 *)a, (float2 *)b, bytes / 262144); });
        snprintf(nm, sizeof nm, "1 : 2, whole 2 KB rows, %d/SM", per_sm);
        time(nm, 1.5 * bytes, [&] { k_mix12_rows<256><<<148 * per_sm, 256>>>((const int4 *)a, (float2 *)b, bytes / 1048576); });
    }
    for (int per_sm : {2, 4, 8}) {
        char nm[64];
        snprintf(nm, sizeof nm, "1 : 2, 128 B into 32-tiled, %d/SM", per_sm);
        time(nm, 1.5 * bytes, [&] { k_mix12_tiled32<<<148 * per_sm, 256>>>((const int4 *)a, (float2 *)b, bytes / 65536); });
    }
    cudaFuncSetAttribute(k_tma_rows<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_tma_rows<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_tma_rows<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int per_sm : {2, 3}) {
        char nm[64];
        snprintf(nm, sizeof nm, "read rows, 1 x 2 KB, %d/SM", per_sm);
        time(nm, (double)bytes, [&] { k_tma_rows<1><<<148 * per_sm, 256, 65536>>>(a, bytes / 2048, (int *)b); });
        snprintf(nm, sizeof nm, "read rows, 8 x 256 B, %d/SM", per_sm);
        time(nm, (double)bytes, [&] { k_tma_rows<8><<<148 * per_sm, 256, 65536>>>(a, bytes / 2048, (int *)b); });
        snprintf(nm, sizeof nm, "read rows, 16 x 128 B, %d/SM", per_sm);
        time(nm, (double)bytes, [&] { k_tma_rows<16><<<148 * per_sm, 256, 65536>>>(a, bytes / 2048, (int *)b); });
    }
    cudaFuncSetAttribute(k_tma_read<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_tma_read<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    for (int per_sm : {1, 2, 3}) {
        char nm[64];
        snprintf(nm, sizeof nm, "read  TMA bulk 32K x2, %d/SM", per_sm);
        time(nm, (double)bytes, [&] { k_tma_read<32768><<<148 * per_sm, 256, 65536>>>(a, bytes / 32768, (int *)b); });
    }
    for (int per_sm : {2, 4, 6}) {
        char nm[64];
        snprintf(nm, sizeof nm, "read  TMA bulk 16K x2, %d/SM", per_sm);
        time(nm, (double)bytes, [&] { k_tma_read<16384><<<148 * per_sm, 256, 32768>>>(a, bytes / 16384, (int *)b); });
    }
    return 0;
}
