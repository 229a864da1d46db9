/*
 * mmw_oracle.c — CPU oracle (fp64) for the mmWave radar hot path.
 * TEST INFRASTRUCTURE ONLY — see mmw_oracle.h for the rules and parity status.
 *
 * Compile with -ffp-contract=off: the legacy functions are meant to reproduce
 * the reference CPU path bit for bit (the reference is built without FMA).
 */
#include "mmw_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* legacy path: restatement of the reference CPU functions                    */
/* ------------------------------------------------------------------------- */

/* reference: cudaBenchMarking.cpp:108-118 (bit-smearing round-up) */
int orc_next_pow2(int n)
{
    unsigned v = (unsigned)(n - 1);
    for (int sh = 1; sh <= 16; sh <<= 1) v |= v >> sh;
    return (int)(v + 1);
}

/* reference: cudaBenchMarking.cpp:61-72 */
int orc_reverse_bits(int num, int bits)
{
    int out = 0;
    for (int b = 0; b < bits; ++b)
        if ((num >> b) & 1) out |= 1 << (bits - 1 - b);
    return out;
}

static inline orc_cx cx_mul(orc_cx a, orc_cx b)   /* :48-54, same operand order */
{
    orc_cx t;
    t.re = a.re * b.re - a.im * b.im;
    t.im = a.re * b.im + a.im * b.re;
    return t;
}

/* reference: cudaBenchMarking.cpp:73-105 — bit-reversal permutation followed by
 * log2(size) radix-2 decimation-in-time stages; the per-stage twiddle is
 * advanced by the recurrence omega <- omega * w (not recomputed), which we keep
 * so that results are bit-identical.  Forward transform, unnormalised. */
void orc_fft(int size, orc_cx *x)
{
    int bits = (int)log2((double)size);
    for (int i = 0; i < size; ++i) {
        int j = orc_reverse_bits(i, bits);
        if (j > i) { orc_cx t = x[i]; x[i] = x[j]; x[j] = t; }
    }
    for (int span = 2; span <= size; span <<= 1) {
        const int half = span / 2;
        const double theta = -2 * M_PI / span;
        const orc_cx w = { cos(theta), sin(theta) };
        for (int blk = 0; blk < size; blk += span) {
            orc_cx omega = { 1.0, 0.0 };
            for (int j = 0; j < half; ++j) {
                orc_cx *lo = &x[blk + j], *hi = &x[blk + j + half];
                orc_cx p = cx_mul(omega, *hi);
                hi->re = lo->re - p.re;  hi->im = lo->im - p.im;
                lo->re = lo->re + p.re;  lo->im = lo->im + p.im;
                omega = cx_mul(omega, w);
            }
        }
    }
}

/* reference: cudaBenchMarking.cpp:149-188.  `size` shorts in IIQQ groups
 * [I(2m) I(2m+1) Q(2m) Q(2m+1)] ordered [chirp][rx][sample]; the output is
 * [rx][chirp][sample].  Elements beyond size/2 are left as zero (the reference
 * reads uninitialised heap there; callers never pass a short frame). */
void orc_reshape(const int16_t *shorts, orc_cx *out, int size, int S, int C, int A)
{
    const long n = (long)S * C * A;
    orc_cx *flat = (orc_cx *)calloc((size_t)n, sizeof(orc_cx));
    for (long g = 0; 4 * g + 3 < size && 2 * g + 1 < n; ++g) {
        const int16_t *q = shorts + 4 * g;
        flat[2 * g].re     = (double)q[0];
        flat[2 * g].im     = (double)q[2];
        flat[2 * g + 1].re = (double)q[1];
        flat[2 * g + 1].im = (double)q[3];
    }
    for (int a = 0; a < A; ++a)
        for (int c = 0; c < C; ++c)
            memcpy(out + ((long)a * C + c) * S, flat + ((long)c * A + a) * S,
                   (size_t)S * sizeof(orc_cx));
    free(flat);
}

/* reference: cudaBenchMarking.cpp:191-206 (strict >, starts from 0 => first max wins) */
int orc_find_abs_max(const orc_cx *x, int size)
{
    int best = 0;
    double best_v = 0;
    for (int i = 0; i < size; ++i) {
        double v = sqrt(x[i].re * x[i].re + x[i].im * x[i].im);
        if (v > best_v) { best_v = v; best = i; }
    }
    return best;
}

/* reference: cudaBenchMarking.cpp:301-303, constants :7,:11,:13 */
double orc_distance_from_raw(int raw, int n_valid, int n_ext)
{
    const double light = 3.0e8, Fs = 2.0e6, mu = 5.987e12;
    double Fs_extend = Fs * n_ext / n_valid;
    int idx = raw * n_valid / n_ext;
    return light * (((double)idx / n_ext) * Fs_extend) / (2 * mu);
}

/* reference: loop body of cpuTiming(), cudaBenchMarking.cpp:273-303 */
double orc_legacy_frame(const int16_t *frame, const orc_cx *base_rx0, int size,
                        int S, int C, int A, orc_cx *spectrum, int *raw_out)
{
    const int n_valid = S * C;
    const int n_ext = orc_next_pow2(n_valid);
    orc_cx *all = (orc_cx *)malloc((size_t)S * C * A * sizeof(orc_cx));
    orc_cx *buf = (orc_cx *)malloc((size_t)n_ext * sizeof(orc_cx));
    orc_reshape(frame, all, size, S, C, A);
    for (int i = 0; i < n_valid; ++i) {            /* rx0 is the first S*C elements */
        buf[i].re = all[i].re - base_rx0[i].re;
        buf[i].im = all[i].im - base_rx0[i].im;
    }
    for (int i = n_valid; i < n_ext; ++i) buf[i].re = buf[i].im = 0;
    orc_fft(n_ext, buf);
    int raw = orc_find_abs_max(buf, (int)floor(0.4 * n_ext));
    if (spectrum) memcpy(spectrum, buf, (size_t)n_ext * sizeof(orc_cx));
    if (raw_out) *raw_out = raw;
    free(all);
    free(buf);
    return orc_distance_from_raw(raw, n_valid, n_ext);
}

/* ------------------------------------------------------------------------- */
/* north-star chain — definitions (no reference counterpart) */
/* ------------------------------------------------------------------------- */

void orc_hann_periodic(int L, float *w)
{
    for (int n = 0; n < L; ++n)
        w[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)n / (double)L));
}

/* n1+n2: int16 -> double, * win_r[s], zero-pad S -> Sp, forward FFT per (chirp, antenna).
 * Packing and frame order follow the reference (cudaBenchMarking.cpp:156-165, :168-180). */
void orc_range_fft(const int16_t *adc, int S, int C, int A, const float *win_r, orc_cx *rs)
{
    orc_range_fft_base(adc, NULL, S, C, A, win_r, rs);
}

/* same with static-clutter removal: `base` (one frame, same packing) is subtracted sample by sample before the
 * window — the reference's base-frame subtraction (cudaBenchMarking.cpp:277-280) for every antenna */
void orc_range_fft_base(const int16_t *adc, const int16_t *base, int S, int C, int A, const float *win_r, orc_cx *rs)
{
    const int Sp = orc_next_pow2(S);
    orc_cx *row = (orc_cx *)malloc((size_t)Sp * sizeof(orc_cx));
    for (int c = 0; c < C; ++c)
        for (int a = 0; a < A; ++a) {
            const int16_t *src = adc + ((long)c * A + a) * S * 2;
            const int16_t *bsrc = base ? base + ((long)c * A + a) * S * 2 : NULL;
            for (int s = 0; s < S; ++s) {
                const long o = 4 * (s >> 1) + (s & 1);
                double w = (double)win_r[s];
                row[s].re = ((double)src[o] - (bsrc ? (double)bsrc[o] : 0.0)) * w;
                row[s].im = ((double)src[o + 2] - (bsrc ? (double)bsrc[o + 2] : 0.0)) * w;
            }
            for (int s = S; s < Sp; ++s) row[s].re = row[s].im = 0;
            orc_fft(Sp, row);
            for (int r = 0; r < Sp; ++r) rs[((long)a * Sp + r) * C + c] = row[r];
        }
    free(row);
}

/* n3: corner turn is implicit in the [a][r][c] layout; window over chirps, pad C -> Cp, FFT */
void orc_doppler_fft(const orc_cx *rs, int Sp, int C, int A, const float *win_d, orc_cx *dc)
{
    const int Cp = orc_next_pow2(C);
    for (long ar = 0; ar < (long)A * Sp; ++ar) {
        const orc_cx *src = rs + ar * C;
        orc_cx *dst = dc + ar * Cp;
        for (int c = 0; c < C; ++c) {
            double w = (double)win_d[c];
            dst[c].re = src[c].re * w;
            dst[c].im = src[c].im * w;
        }
        for (int c = C; c < Cp; ++c) dst[c].re = dst[c].im = 0;
        orc_fft(Cp, dst);
    }
}

/* n4 */
void orc_power(const orc_cx *dc, int Sp, int Cp, int A, double *P)
{
    const long M = (long)Sp * Cp;
    for (long m = 0; m < M; ++m) P[m] = 0;
    for (int a = 0; a < A; ++a) {
        const orc_cx *x = dc + (long)a * M;
        for (long m = 0; m < M; ++m) P[m] += x[m].re * x[m].re + x[m].im * x[m].im;
    }
}

/* n5: training cells of CUT (r,d) = { (r+i, (d+j) mod Cp) : |i|<=Wr, |j|<=Wd,
 *      not(|i|<=Gr and |j|<=Gd), 0 <= r+i < Sp }.  noise = mean over them.
 * The sum runs over the training cells themselves (not outer-box minus
 * inner-box) so that a strong CUT never cancels against itself. */
void orc_cfar(const double *P, int Sp, int Cp, const orc_cfar_params *p,
              uint8_t *mask, double *noise)
{
    const int Gr = p->guard_r, Gd = p->guard_d;
    const int Wr = Gr + p->train_r, Wd = Gd + p->train_d;
    for (int r = 0; r < Sp; ++r)
        for (int d = 0; d < Cp; ++d) {
            double sum = 0;
            int cnt = 0;
            for (int j = -Wd; j <= Wd; ++j) {
                int dd = ((d + j) % Cp + Cp) % Cp;
                int in_guard_d = (j >= -Gd && j <= Gd);
                for (int i = -Wr; i <= Wr; ++i) {
                    int rr = r + i;
                    if (rr < 0 || rr >= Sp) continue;
                    if (in_guard_d && i >= -Gr && i <= Gr) continue;
                    sum += P[(long)rr * Cp + dd];
                    ++cnt;
                }
            }
            double nz = cnt > 0 ? sum / cnt : 0.0;
            long m = (long)r * Cp + d;
            if (noise) noise[m] = nz;
            mask[m] = (uint8_t)(cnt > 0 && P[m] > p->alpha * nz);
        }
}

int orc_angle_fft_size(int A)
{
    return A <= 64 ? 64 : orc_next_pow2(A);
}

/* n6: zero-pad A -> Ntheta, forward DFT over antennas, arg-max |Y|^2 (strict >, first wins) */
int orc_angle_argmax(const orc_cx *x, int A, int n_theta, double *second_ratio)
{
    orc_cx *y = (orc_cx *)calloc((size_t)n_theta, sizeof(orc_cx));
    memcpy(y, x, (size_t)A * sizeof(orc_cx));
    orc_fft(n_theta, y);
    int best = 0;
    double best_v = 0, second = 0;
    for (int k = 0; k < n_theta; ++k) {
        double v = y[k].re * y[k].re + y[k].im * y[k].im;
        if (v > best_v) { second = best_v; best_v = v; best = k; }
        else if (v > second) second = v;
    }
    if (second_ratio) *second_ratio = best_v > 0 ? second / best_v : 1.0;
    free(y);
    return best;
}

double orc_angle_rad(int k_wrapped, int n_theta, double lambda_over_d)
{
    double s = (double)k_wrapped * lambda_over_d / (double)n_theta;
    if (s > 1) s = 1;
    if (s < -1) s = -1;
    return asin(s);
}

/* n7 */
int orc_is_group_peak(const double *P, const uint8_t *mask, int Sp, int Cp, int r, int d)
{
    const double v = P[(long)r * Cp + d];
    const long key = (long)r * Cp + d;
    for (int i = -1; i <= 1; ++i)
        for (int j = -1; j <= 1; ++j) {
            if (!i && !j) continue;
            int rr = r + i;
            if (rr < 0 || rr >= Sp) continue;
            int dd = ((d + j) % Cp + Cp) % Cp;
            long m = (long)rr * Cp + dd;
            if (!mask[m]) continue;
            if (P[m] > v) return 0;
            if (P[m] == v && m < key) return 0;
        }
    return 1;
}

typedef struct {
    orc_cx *rs, *dc, *x;
    double *P, *nz;
    uint8_t *mask;
} orc_scratch;

static void scratch_alloc(orc_scratch *w, int Sp, int C, int Cp, int A)
{
    const long M = (long)Sp * Cp;
    w->rs = (orc_cx *)malloc((size_t)A * Sp * C * sizeof(orc_cx));
    w->dc = (orc_cx *)malloc((size_t)A * M * sizeof(orc_cx));
    w->P = (double *)malloc((size_t)M * sizeof(double));
    w->mask = (uint8_t *)malloc((size_t)M);
    w->nz = (double *)malloc((size_t)M * sizeof(double));
    w->x = (orc_cx *)malloc((size_t)A * sizeof(orc_cx));
}

static void scratch_free(orc_scratch *w)
{
    free(w->rs); free(w->dc); free(w->P); free(w->mask); free(w->nz); free(w->x);
}

/* one frame; per-thread scratch is reused across frames (the caller's optional output buffers win) */
static const int16_t *g_base_frame = NULL;   /* set by orc_process_frames_base for the duration of the call */
static double *g_angle_ratio = NULL;         /* set by orc_process_frames_ratio: per detection, 2nd-largest / largest angle-bin power */

static const orc_detection *g_dets0 = NULL;  /* start of the caller's detection buffer (index base of g_angle_ratio) */

static long process_one(const int16_t *adc, int f, int S, int C, int A,
                        const float *win_r, const float *win_d,
                        const orc_cfar_params *p, double lambda_over_d,
                        orc_detection *dets, long cap, orc_scratch *w,
                        orc_cx *rs_out, orc_cx *dc_out, double *P_out,
                        uint8_t *mask_out, double *noise_out, long *total)
{
    const int Sp = orc_next_pow2(S), Cp = orc_next_pow2(C);
    const long M = (long)Sp * Cp;
    const int n_theta = orc_angle_fft_size(A);
    orc_cx *rs = rs_out ? rs_out : w->rs;
    orc_cx *dc = dc_out ? dc_out : w->dc;
    double *P = P_out ? P_out : w->P;
    uint8_t *mask = mask_out ? mask_out : w->mask;
    double *nz = noise_out ? noise_out : w->nz;
    orc_cx *x = w->x;

    orc_range_fft_base(adc, g_base_frame, S, C, A, win_r, rs);
    orc_doppler_fft(rs, Sp, C, A, win_d, dc);
    orc_power(dc, Sp, Cp, A, P);
    orc_cfar(P, Sp, Cp, p, mask, nz);

    long n = 0, tot = 0;
    for (int r = 0; r < Sp; ++r)
        for (int d = 0; d < Cp; ++d) {
            long m = (long)r * Cp + d;
            if (!mask[m]) continue;
            ++tot;
            if (n >= cap) continue;
            for (int a = 0; a < A; ++a) x[a] = dc[(long)a * M + m];
            double ratio = 0;
            int k = orc_angle_argmax(x, A, n_theta, &ratio);
            int kw = k < n_theta / 2 ? k : k - n_theta;
            if (g_angle_ratio) g_angle_ratio[(dets - g_dets0) + n] = ratio;
            orc_detection *o = &dets[n++];
            o->frame = (uint32_t)f;
            o->range_bin = (uint16_t)r;
            o->doppler_bin = (uint16_t)d;
            o->power = (float)P[m];
            o->noise = (float)nz[m];
            o->angle_bin = (int16_t)kw;
            o->flags = (uint16_t)orc_is_group_peak(P, mask, Sp, Cp, r, d);
            o->angle_rad = (float)orc_angle_rad(kw, n_theta, lambda_over_d);
        }
    *total = tot;
    return n;
}

long orc_process_frames(const int16_t *adc, int n_frames, int S, int C, int A,
                        const float *win_r, const float *win_d,
                        const orc_cfar_params *p, double lambda_over_d,
                        orc_detection *dets, long det_cap, long *n_total,
                        orc_cx *rs_out, orc_cx *dc_out, double *P_out,
                        uint8_t *mask_out, double *noise_out, int n_threads)
{
    const int Sp = orc_next_pow2(S), Cp = orc_next_pow2(C);
    const long M = (long)Sp * Cp;
    const long frame_shorts = 2L * S * C * A;
    /* every frame gets an equal slice of the caller's buffer so that frames can
     * run in parallel; slices are compacted afterwards, preserving frame order */
    const long per = n_frames > 0 ? det_cap / n_frames : 0;
    g_dets0 = dets;
    long *cnt = (long *)calloc((size_t)(n_frames > 0 ? n_frames : 1), sizeof(long));
    long *tot = (long *)calloc((size_t)(n_frames > 0 ? n_frames : 1), sizeof(long));
    if (n_threads < 1) n_threads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
    {
        orc_scratch w;
        scratch_alloc(&w, Sp, C, Cp, A);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int f = 0; f < n_frames; ++f)
            cnt[f] = process_one(adc + (long)f * frame_shorts, f, S, C, A, win_r, win_d, p,
                                 lambda_over_d, dets + (long)f * per, per, &w,
                                 rs_out ? rs_out + (long)f * A * Sp * C : NULL,
                                 dc_out ? dc_out + (long)f * A * M : NULL,
                                 P_out ? P_out + (long)f * M : NULL,
                                 mask_out ? mask_out + (long)f * M : NULL,
                                 noise_out ? noise_out + (long)f * M : NULL, &tot[f]);
        scratch_free(&w);
    }
    long n = 0, t = 0;
    for (int f = 0; f < n_frames; ++f) {
        if (n != (long)f * per && cnt[f] > 0) {
            memmove(dets + n, dets + (long)f * per, (size_t)cnt[f] * sizeof(orc_detection));
            if (g_angle_ratio) memmove(g_angle_ratio + n, g_angle_ratio + (long)f * per, (size_t)cnt[f] * sizeof(double));
        }
        n += cnt[f];
        t += tot[f];
    }
    if (n_total) *n_total = t;
    free(cnt);
    free(tot);
    (void)n_threads;
    return n;
}

/* orc_process_frames with static-clutter removal (base = one frame or NULL); not re-entrant */
long orc_process_frames_base(const int16_t *adc, const int16_t *base, int n_frames, int S, int C, int A,
                             const float *win_r, const float *win_d,
                             const orc_cfar_params *p, double lambda_over_d,
                             orc_detection *dets, long det_cap, long *n_total,
                             orc_cx *rs_out, orc_cx *dc_out, double *P_out,
                             uint8_t *mask_out, double *noise_out, int n_threads)
{
    g_base_frame = base;
    long n = orc_process_frames(adc, n_frames, S, C, A, win_r, win_d, p, lambda_over_d, dets, det_cap, n_total,
                                rs_out, dc_out, P_out, mask_out, noise_out, n_threads);
    g_base_frame = NULL;
    return n;
}

/* orc_process_frames_base that also reports, per detection written, the ratio of the second-largest to the largest bin of
 * its angle spectrum (angle_ratio[det_cap]; lets a checker skip near-tie arg-maxes without keeping the whole Doppler
 * cube); not re-entrant */
long orc_process_frames_ratio(const int16_t *adc, const int16_t *base, int n_frames, int S, int C, int A,
                              const float *win_r, const float *win_d,
                              const orc_cfar_params *p, double lambda_over_d,
                              orc_detection *dets, long det_cap, long *n_total,
                              orc_cx *rs_out, orc_cx *dc_out, double *P_out,
                              uint8_t *mask_out, double *noise_out, int n_threads, double *angle_ratio)
{
    g_angle_ratio = angle_ratio;
    long n = orc_process_frames_base(adc, base, n_frames, S, C, A, win_r, win_d, p, lambda_over_d, dets, det_cap, n_total,
                                     rs_out, dc_out, P_out, mask_out, noise_out, n_threads);
    g_angle_ratio = NULL;
    return n;
}
