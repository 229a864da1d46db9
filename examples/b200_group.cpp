// b200_group.cpp — a C++ host sharding one batch of frames over every GPU of the box through include/mmw_radar.h only
// (no torch, no NCCL header, no MPI): mmw_group_create builds one context per device and one NCCL communicator inside the
// library; mmw_group_process_host cuts the batch into contiguous frame blocks, runs the chain on every GPU and gathers the
// detection lists to GPU 0.  The program then runs the same batch on GPU 0 alone and requires the two lists to be identical,
// byte for byte — the acceptance test of frame sharding.  The reference itself is single-GPU, one frame per
// call (cudaBenchMarking.cpp:374-378).
//
//   b200_group <samples> <chirps> <antennas> <frames> [n_gpus (default: all)]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime_api.h>

#include "mmw_radar.h"

int main(int argc, char **argv)
{
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s samples chirps antennas frames [n_gpus]\n", argv[0]);
        return 2;
    }
    const int S = std::atoi(argv[1]), C = std::atoi(argv[2]), A = std::atoi(argv[3]), F = std::atoi(argv[4]);
    int n_gpus = 0;
    if (cudaGetDeviceCount(&n_gpus) != cudaSuccess || n_gpus < 1) {
        std::printf("no CUDA device\n");
        return 1;
    }
    if (argc > 5 && std::atoi(argv[5]) > 0 && std::atoi(argv[5]) < n_gpus) n_gpus = std::atoi(argv[5]);

    // a deterministic synthetic capture in the reference's format ([chirp][antenna][sample], IIQQ int16): two tones + LCG noise
    const size_t frame_shorts = (size_t)2 * S * C * A;
    int16_t *adc = nullptr;
    if (cudaMallocHost((void **)&adc, F * frame_shorts * sizeof(int16_t)) != cudaSuccess) return 1;   // pinned: async uploads
    unsigned lcg = 12345u;
    for (int f = 0; f < F; ++f)
        for (int c = 0; c < C; ++c)
            for (int a = 0; a < A; ++a)
                for (int s = 0; s < S; ++s) {
                    lcg = lcg * 1664525u + 1013904223u;
                    const int noise_i = (int)((lcg >> 16) & 63) - 32;
                    lcg = lcg * 1664525u + 1013904223u;
                    const int noise_q = (int)((lcg >> 16) & 63) - 32;
                    // tone 1 sits on range bin S/8, Doppler bin C/4 + f (moves with the frame); quarter-cycle tables keep this libm-free
                    const int ph = (4 * s * (S / 8) / S + 4 * c * ((C / 4 + f) % C) / C + a) & 3;
                    static const int cs[4] = {1, 0, -1, 0}, sn[4] = {0, 1, 0, -1};
                    const int i_v = 900 * cs[ph] + noise_i, q_v = 900 * sn[ph] + noise_q;
                    int16_t *grp = adc + (((size_t)f * C + c) * A + a) * 2 * S + 4 * (s >> 1);
                    grp[s & 1] = (int16_t)i_v;
                    grp[2 + (s & 1)] = (int16_t)q_v;
                }

    mmw_config cfg;
    const int per_gpu = (F + n_gpus - 1) / n_gpus;
    mmw_default_config(&cfg, S, C, A, per_gpu);
    std::vector<int> devices(n_gpus);
    for (int i = 0; i < n_gpus; ++i) devices[i] = i;
    mmw_group *group = nullptr;
    if (mmw_group_create(&cfg, devices.data(), n_gpus, &group) != MMW_OK) {
        std::printf("mmw_group_create: %s\n", mmw_last_error());
        return 1;
    }
    std::vector<mmw_detection> sharded((size_t)F * cfg.max_det_per_frame), single(sharded.size());
    int n_sharded = 0, n_single = 0;
    int rc = mmw_group_process_host(group, adc, F, sharded.data(), (int)sharded.size(), &n_sharded);
    if (rc != MMW_OK) {
        std::printf("mmw_group_process_host: %s\n", mmw_last_error());
        return 1;
    }
    for (int i = 0; i < n_gpus; ++i) {
        int first = 0, count = 0;
        mmw_shard_frames(F, n_gpus, i, &first, &count);
        std::printf("gpu %d: frames [%d, %d)\n", devices[i], first, first + count);
    }
    mmw_group_destroy(group);

    mmw_default_config(&cfg, S, C, A, F);
    cfg.device = 0;
    mmw_ctx *ctx = nullptr;
    if (mmw_create(&cfg, &ctx) != MMW_OK || mmw_process_host(ctx, adc, F, single.data(), (int)single.size(), &n_single) != MMW_OK) {
        std::printf("single-GPU run: %s\n", mmw_last_error());
        return 1;
    }
    mmw_destroy(ctx);
    const bool same = n_sharded == n_single && std::memcmp(sharded.data(), single.data(), (size_t)n_single * sizeof(mmw_detection)) == 0;
    std::printf("%d GPUs: %d detections gathered to GPU 0, single GPU: %d, identical=%s\n", n_gpus, n_sharded, n_single, same ? "yes" : "NO");
    cudaFreeHost(adc);
    return same && n_single > 0 ? 0 : 1;
}
