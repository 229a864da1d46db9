// b200_timing.cpp — the C++ host program a maintainer of the reference would write next to cudaTiming()
// (cudaBenchMarking.cpp:334-395): same capture file, same Timer-style report, but the batched chain of
// include/mmw_radar.h instead of one cudaProcessing() call per frame.  Plain C++ against the C ABI; no torch.
//
//   b200_timing <capture.bin> <samples> <chirps> <antennas> [batch_frames] [first_frame_is_base(0|1)]
//
// Prints one line per stage of the report and the first detections in physical units.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mmw_radar.h"

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv)
{
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s capture.bin samples chirps antennas [batch_frames] [first_frame_is_base]\n", argv[0]);
        return 2;
    }
    const char *path = argv[1];
    const int S = std::atoi(argv[2]), C = std::atoi(argv[3]), A = std::atoi(argv[4]);
    const int batch = argc > 5 ? std::atoi(argv[5]) : 64;
    const int use_base = argc > 6 ? std::atoi(argv[6]) : 0;

    mmw_config cfg;
    mmw_default_config(&cfg, S, C, A, batch);            // Hann windows, CFAR guard 2x2 / train 8x4 / alpha 15
    mmw_ctx *ctx = nullptr;
    if (mmw_create(&cfg, &ctx) != MMW_OK) {
        std::printf("%s\n", mmw_last_error());
        return 1;
    }
    mmw_info info;
    mmw_get_info(ctx, &info);

    std::vector<mmw_detection> dets(1 << 20);
    int n_det = 0, n_frames = 0;
    const double t0 = now_s();
    const int rc = mmw_process_capture_file(ctx, path, 0, 0, use_base, dets.data(), (int)dets.size(), &n_det, &n_frames);
    const double t = now_s() - t0;
    if (rc != MMW_OK && rc != MMW_ERR_OVERFLOW) {
        std::printf("%s\n", mmw_last_error());               // e.g. "unable to read the specified file"
        mmw_destroy(ctx);
        return 1;
    }
    std::printf("b200 totalTime %.5f ms average %.5f ms/frame b200FPS %.5f FPS (%d frames, file read included)\n", 1000.0 * t,
                n_frames ? 1000.0 * t / n_frames : 0.0, t > 0 ? n_frames / t : 0.0, n_frames);
    std::printf("b200 detections %d%s, range FFT %d, Doppler FFT %d, angle FFT %d, %.1f MB workspace\n", n_det,
                rc == MMW_ERR_OVERFLOW ? " (list truncated)" : "", info.Sp, info.Cp, info.n_theta, info.workspace_bytes / 1e6);

    mmw_radar_params rp;
    mmw_default_radar_params(&rp);                        // F0, mu, Fs, Tr of cudaBenchMarking.cpp:10-19
    const int show = n_det < 8 ? n_det : 8;
    std::vector<mmw_target> tg(show > 0 ? show : 1);
    if (show > 0 && mmw_to_physical(&rp, info.Sp, info.Cp, dets.data(), show, tg.data()) == MMW_OK)
        for (int i = 0; i < show; ++i)
            std::printf("  frame %u  range %.3f m  velocity %+.3f m/s  angle %+.1f deg  snr %.1f dB%s\n", tg[i].frame, tg[i].range_m,
                        tg[i].velocity_mps, tg[i].angle_deg, tg[i].snr_db, (tg[i].flags & MMW_FLAG_PEAK) ? "  (peak)" : "");
    mmw_destroy(ctx);
    return 0;
}
