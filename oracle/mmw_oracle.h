/*
 * mmw_oracle.h — CPU oracle for the mmWave radar hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the reported CPU baseline.  The CUDA library never links or calls it.
 *
 * Parity status
 *   - legacy stages (unpack, reshape, base-frame subtract, zero-pad, FFT,
 *     arg-max, distance): PINNED.  Restated from the reference CPU path
 *     (cudaBenchMarking.cpp:61-105, :149-206, :273-303) and checked
 *     bit-for-bit against the reference's own functions compiled from
 *     /root/reference into oracle/_ref (tests/test_oracle_pin.py) and against
 *     the golden fixtures generated from them (tests/golden/).
 *   - range-FFT-per-chirp / Doppler FFT / |X|^2 integration / CA-CFAR /
 *     angle FFT / peak grouping: PARITY UNPINNED by the reference — the
 *     reference has no code for these stages.  The
 *     definitions below ARE the specification; they reuse every convention
 *     the reference does fix (IIQQ int16 packing, [chirp][ant][sample] frame
 *     order, forward unnormalised FFT, zero-pad to nextPow2, strict->
 *     first-wins arg-max) and are cross-checked against numpy in the tests.
 *
 * All arithmetic is fp64.  Window tables are passed in as fp32 so that the
 * oracle and the GPU multiply by exactly the same numbers.
 */
#ifndef MMW_ORACLE_H
#define MMW_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } orc_cx;   /* same layout as the reference Complex_t (acceleration.h:27-30) */

/* detection record, byte-identical to mmw_detection in include/mmw_radar.h */
typedef struct {
    uint32_t frame;
    uint16_t range_bin;
    uint16_t doppler_bin;
    float    power;
    float    noise;
    int16_t  angle_bin;      /* wrapped to [-Ntheta/2, Ntheta/2) */
    uint16_t flags;          /* bit0: local peak after 3x3 grouping */
    float    angle_rad;
} orc_detection;

typedef struct {
    int guard_r, guard_d;    /* guard half-widths  */
    int train_r, train_d;    /* training half-width beyond the guard */
    double alpha;            /* threshold multiplier */
} orc_cfar_params;

/* ---- legacy path (reference cfg: 100 samples x 128 chirps x 4 rx) ---- */
int    orc_next_pow2(int n);                                   /* cudaBenchMarking.cpp:108-118 */
int    orc_reverse_bits(int num, int bits);                    /* :61-72 */
void   orc_fft(int size, orc_cx *x);                           /* :73-105, in place */
void   orc_reshape(const int16_t *shorts, orc_cx *out, int size,
                   int S, int C, int A);                       /* :149-188 */
int    orc_find_abs_max(const orc_cx *x, int size);            /* :191-206 */
double orc_distance_from_raw(int raw, int n_valid, int n_ext); /* :301-303 */
/* one frame through reshape -> rx0 - base -> pad -> FFT -> argmax -> metres.
 * spectrum (optional) receives the n_ext-point FFT, raw_out the raw arg-max. */
double orc_legacy_frame(const int16_t *frame, const orc_cx *base_rx0, int size,
                        int S, int C, int A, orc_cx *spectrum, int *raw_out);

/* ---- north-star chain (per frame) ---- */
/* Hann, periodic: w[n] = 0.5 - 0.5 cos(2 pi n / L), computed in fp64, rounded to fp32 */
void orc_hann_periodic(int L, float *w);

/* adc: one frame [C][A][S] complex int16 in IIQQ packing (2*S*A*C shorts).
 * rs : [A][Sp][C] complex (Sp = nextPow2(S)), range bin major, chirp minor.
 * The Doppler window is NOT applied here. */
void orc_range_fft(const int16_t *adc, int S, int C, int A,
                   const float *win_r, orc_cx *rs);
void orc_range_fft_base(const int16_t *adc, const int16_t *base, int S, int C, int A,
                        const float *win_r, orc_cx *rs);
/* dc: [A][Sp][Cp] complex (Cp = nextPow2(C)); Doppler window applied over c<C, zero pad to Cp */
void orc_doppler_fft(const orc_cx *rs, int Sp, int C, int A,
                     const float *win_d, orc_cx *dc);
/* P[r][d] = sum_a |dc[a][r][d]|^2, ascending a */
void orc_power(const orc_cx *dc, int Sp, int Cp, int A, double *P);
/* 2-D CA-CFAR: range axis clamps (n_train recounted), Doppler axis wraps.
 * mask[r*Cp+d] = 1 iff P > alpha*noise ; noise[r*Cp+d] = training mean. */
void orc_cfar(const double *P, int Sp, int Cp, const orc_cfar_params *p,
              uint8_t *mask, double *noise);
/* angle spectrum arg-max for one cell. x: A antenna samples. Returns raw bin k in [0,Ntheta);
 * second_ratio (optional) = 2nd-largest/largest |Y|^2 to let tests skip near-ties. */
int  orc_angle_argmax(const orc_cx *x, int A, int n_theta, double *second_ratio);
int  orc_angle_fft_size(int A);                   /* 64 for A<=64 else nextPow2(A) */
double orc_angle_rad(int k_wrapped, int n_theta, double lambda_over_d);
/* 3x3 grouping among detected cells: 1 iff strict local maximum (ties -> lowest (r,d)) */
int  orc_is_group_peak(const double *P, const uint8_t *mask, int Sp, int Cp, int r, int d);

/* whole chain for n_frames frames; detections sorted by (frame, r, d).
 * Any of rs_out/dc_out/P_out/mask_out/noise_out may be NULL.
 * Returns the number of detections written (<= det_cap; the true total goes to *n_total). */
long orc_process_frames(const int16_t *adc, int n_frames, int S, int C, int A,
                        const float *win_r, const float *win_d,
                        const orc_cfar_params *p, double lambda_over_d,
                        orc_detection *dets, long det_cap, long *n_total,
                        orc_cx *rs_out, orc_cx *dc_out, double *P_out,
                        uint8_t *mask_out, double *noise_out,
                        int n_threads);

/* same with static-clutter removal: base = one frame in capture format subtracted before the window, or NULL */
long orc_process_frames_base(const int16_t *adc, const int16_t *base, int n_frames, int S, int C, int A,
                             const float *win_r, const float *win_d,
                             const orc_cfar_params *p, double lambda_over_d,
                             orc_detection *dets, long det_cap, long *n_total,
                             orc_cx *rs_out, orc_cx *dc_out, double *P_out,
                             uint8_t *mask_out, double *noise_out,
                             int n_threads);

/* same, also reporting per detection written the ratio 2nd-largest / largest bin power of its angle spectrum
 * (angle_ratio[det_cap], aligned with dets[]); lets a checker skip near-tie arg-maxes without the whole Doppler cube */
long orc_process_frames_ratio(const int16_t *adc, const int16_t *base, int n_frames, int S, int C, int A,
                              const float *win_r, const float *win_d,
                              const orc_cfar_params *p, double lambda_over_d,
                              orc_detection *dets, long det_cap, long *n_total,
                              orc_cx *rs_out, orc_cx *dc_out, double *P_out,
                              uint8_t *mask_out, double *noise_out,
                              int n_threads, double *angle_ratio);

#ifdef __cplusplus
}
#endif
#endif
