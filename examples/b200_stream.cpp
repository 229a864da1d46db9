// b200_stream.cpp — the per-frame loop of cudaTiming() (cudaBenchMarking.cpp:374-378: fread one frame, process it, next)
// kept as a per-frame loop, but without its blocking: a ring of contexts, mmw_submit_host on the frame just read,
// mmw_wait on the frame submitted `depth` reads ago.  The upload of one frame runs under the kernels of the previous
// ones; every frame still gets its own call and its own detection list (a sensor's frame is handled when it arrives).
// Plain C++ against the C ABI; pinned frame buffers come from cudaHostAlloc.
//
//   b200_stream <capture.bin> <samples> <chirps> <antennas> [depth]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime_api.h>

#include "mmw_radar.h"

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv)
{
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s capture.bin samples chirps antennas [depth]\n", argv[0]);
        return 2;
    }
    const int S = std::atoi(argv[2]), C = std::atoi(argv[3]), A = std::atoi(argv[4]);
    const int depth = argc > 5 ? std::atoi(argv[5]) : 4;
    if (depth < 1 || depth > 16) { std::fprintf(stderr, "depth must be 1..16\n"); return 2; }
    FILE *fp = std::fopen(argv[1], "rb");
    if (!fp) { std::printf("unable to read the specified file: %s\n", argv[1]); return 1; }

    mmw_config cfg;
    mmw_default_config(&cfg, S, C, A, 1);                   // one frame per call
    const size_t frame_shorts = (size_t)2 * S * C * A;
    std::vector<mmw_ctx *> ring(depth, nullptr);
    std::vector<int16_t *> buf(depth, nullptr);
    std::vector<int> busy(depth, 0);
    for (int i = 0; i < depth; ++i) {
        if (mmw_create(&cfg, &ring[i]) != MMW_OK) { std::printf("%s\n", mmw_last_error()); return 1; }
        mmw_set_graph_mode(ring[i], 1);                     // one graph replay per frame instead of six launches
        if (cudaHostAlloc((void **)&buf[i], frame_shorts * sizeof(int16_t), cudaHostAllocDefault) != cudaSuccess) {
            std::printf("cudaHostAlloc failed\n");
            return 1;
        }
    }
    std::vector<mmw_detection> dets(cfg.max_det_per_frame);
    long long total_det = 0;
    int frames = 0, collected = 0, overflowed = 0;
    auto collect = [&](int slot) {
        int n = 0;
        const int rc = mmw_wait(ring[slot], dets.data(), (int)dets.size(), &n);
        if (rc != MMW_OK && rc != MMW_ERR_OVERFLOW) { std::printf("%s\n", mmw_last_error()); std::exit(1); }
        overflowed += rc == MMW_ERR_OVERFLOW;
        total_det += n;
        ++collected;
        busy[slot] = 0;
    };
    const double t0 = now_s();
    for (;; ++frames) {
        const int slot = frames % depth;
        if (busy[slot]) collect(slot);                      // the frame submitted `depth` reads ago
        if (std::fread(buf[slot], sizeof(int16_t), frame_shorts, fp) != frame_shorts) break;
        mmw_set_frame_offset(ring[slot], (uint32_t)frames); // mmw_detection.frame = position in the file
        if (mmw_submit_host(ring[slot], buf[slot], 1) != MMW_OK) { std::printf("%s\n", mmw_last_error()); return 1; }
        busy[slot] = 1;
    }
    for (int i = 0; i < depth; ++i) {                       // drain in submission order
        const int slot = (frames + i) % depth;
        if (busy[slot]) collect(slot);
    }
    const double t = now_s() - t0;
    std::printf("b200 stream totalTime %.5f ms average %.5f ms/frame %.1f FPS (%d frames, %d in flight, file read included)\n", 1000.0 * t,
                frames ? 1000.0 * t / frames : 0.0, t > 0 ? frames / t : 0.0, frames, depth);
    std::printf("b200 stream detections %lld in %d frames%s\n", total_det, collected, overflowed ? " (some lists truncated)" : "");
    for (int i = 0; i < depth; ++i) {
        mmw_destroy(ring[i]);
        cudaFreeHost(buf[i]);
    }
    std::fclose(fp);
    return 0;
}
