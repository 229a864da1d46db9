/*
 * mmw_legacy.h — drop-in for the reference's only entry point.
 *
 * The reference declares (acceleration.h:27-32, C++ linkage, no extern "C"):
 *
 *     struct Complex_t { double real, imag; };
 *     double cudaProcessing(short *deviceIn, Complex_t *host_baseFrame, int size,
 *                           double *fftTime, double *preProcessTime,
 *                           double *findMaxTime, double *totalTime);
 *
 * which mangles to _Z14cudaProcessingPsP9Complex_tiPdS2_S2_S2_.  libmmw_radar_b200.so exports
 * exactly that symbol, so the reference's UNMODIFIED cudaBenchMarking.cpp (which keeps
 * including its own acceleration.h) links against this library instead of acceleration.o:
 *
 *     g++ -O3 -o acceleration objs/cudaBenchMarking.o -L<dir> -lmmw_radar_b200
 *
 * Behaviour kept from acceleration.cu:417-572: `deviceIn` is a HOST pointer to `size` int16 of one
 * frame (100 samples x 128 chirps x 4 rx, IIQQ); `host_baseFrame` is 12 800 fp64 complex values
 * (rx0 of the base frame, [chirp][sample]); the return value is the distance in metres of the
 * strongest bin below 0.4 * 16384 of the 16 384-point FFT of (rx0 - base); the four timers are
 * accumulated (+=) in seconds; one "Inner CUDA Timing" line is printed per call (set
 * MMW_LEGACY_QUIET=1 to silence); a CUDA failure terminates the process with exit(code) — after
 * printing the error to stderr, which the reference omits (acceleration.cu:19-31).
 * Deviations, on purpose: element 12 800 of the padded buffer is zero as in the reference CPU path
 * (cudaBenchMarking.cpp:281), not uninitialised as in acceleration.cu:156; no racy butterfly.
 *
 * This header deliberately does not redeclare the reference's Timer class: it is not part of the
 * link contract.
 */
#ifndef MMW_LEGACY_H
#define MMW_LEGACY_H

#ifdef __cplusplus

#ifndef ACCELERATION_H           /* the reference header already declares these two */
struct Complex_t {
    double real, imag;
};
double cudaProcessing(short *deviceIn, Complex_t *host_baseFrame, int size, double *fftTime, double *preProcessTime,
                      double *findMaxTime, double *totalTime);
#endif

extern "C" {
#endif

/* C-ABI twin of cudaProcessing for FFI callers (ctypes/cgo/JNI): base_frame is 12 800 (re, im) fp64 pairs.
 * raw_index (optional) receives the raw arg-max bin before the reference's index rescale.
 * Returns the distance in metres, or a negative MMW_ERR_* code cast to double on failure
 * (this variant never calls exit()). */
double mmw_legacy_process_frame(const short *frame_host, const double *base_frame_host, int size, int *raw_index);

/* Same chain for `n_frames` frames already on the host, pipelined through pinned staging buffers;
 * distances[n_frames] receives the per-frame results. Returns MMW_OK or a negative error. */
int mmw_legacy_process_frames(const short *frames_host, int n_frames, const double *base_frame_host, int size,
                              double *distances, int *raw_indices);

/* Same chain for frames already resident in HBM (frames_dev: n_frames * 102 400 int16, device pointer); results stay
 * on the device (raw_dev: n_frames ints) — asynchronous on the legacy stream, for device-timed throughput. */
int mmw_legacy_process_device(const short *frames_dev, int n_frames, const double *base_frame_host, int *raw_dev);
/* Blocks until everything queued by mmw_legacy_process_device has finished. */
int mmw_legacy_sync(void);
/* metres from a raw arg-max bin, operation for operation the reference's formula (acceleration.cu:521-523) */
double mmw_legacy_distance_from_raw(int raw_index);

/* The reference's cudaTiming() loop (cudaBenchMarking.cpp:334-395) in one call: opens the capture at `path`, takes
 * rx0 of frame 0 as the base frame (:357-365), runs every later frame and stores its distance.  distances[] /
 * raw_indices[] (optional) hold up to `capacity` results; *n_frames receives the number of frames processed. */
int mmw_legacy_process_file(const char *path, double *distances, int *raw_indices, int capacity, int *n_frames);

/* Copies the 16 384-point spectrum (complex64, natural order) of the last frame processed through
 * either entry point to `out` (16384 * 2 floats). For parity tests. */
int mmw_legacy_copy_spectrum(float *out);

/* Releases the lazily created device state (also done at process exit). */
void mmw_legacy_shutdown(void);

/* Optional configuration of the legacy path (the reference has none: acceleration.cu:7-15 are compile-time defines).
 * kernel_variant: 0 = pick by batch size (default), 1 = always one CTA per frame, 2 = always the 8-CTA cluster kernel;
 * quiet != 0 silences cudaProcessing's per-call "Inner CUDA Timing" line (acceleration.cu:533).  A negative value keeps
 * the current setting.  Without this call the environment (MMW_LEGACY_VARIANT, MMW_LEGACY_QUIET) is read once, at first use.
 * The device state binds to the CUDA device that is current at the first processing call; later calls make that
 * device current again.  All legacy entry points serialise on one mutex. */
int mmw_legacy_configure(int kernel_variant, int quiet);

#ifdef __cplusplus
}
#endif
#endif
