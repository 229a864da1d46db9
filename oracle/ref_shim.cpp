// ref_shim.cpp — exposes the REFERENCE's own CPU functions through a C ABI so
// that the oracle restatement can be pinned against them.
//
// TEST INFRASTRUCTURE ONLY.  This file contains no reference code: it
// #includes the reference translation unit from where it lies
// (/root/reference/cudaBenchMarking.cpp, passed with -I by oracle/Makefile)
// with its main() renamed, and adds thin extern "C" trampolines.  The output
// (oracle/_ref/libref_cpu.so) is git-ignored; it only exists where
// /root/reference was present at build time.
#define main ref_main_unused
#include "cudaBenchMarking.cpp"
#undef main
#undef c
#undef pi

// The reference's cudaTiming() references the GPU entry point; the CPU-only
// shim never calls it.
double cudaProcessing(short *, Complex_t *, int, double *, double *, double *, double *)
{
    return -1.0;
}

extern "C" {

int ref_sample_size(void) { return SampleSize; }
int ref_chirp_size(void) { return ChirpSize; }
int ref_rx_size(void) { return RxSize; }

int ref_next_pow2(int n) { return nextPow2(n); }
int ref_reverse_bits(int num, int bits) { return reverseBits(num, bits); }
void ref_butterfly_fft(int size, Complex_t *x) { butterfly_fft(size, x); }
void ref_reshape(short *in, Complex_t *out, int size) { ReshapeComplex_t(in, out, size); }
int ref_find_abs_max(Complex_t *x, int size) { return FindAbsMax(x, size); }

// One iteration of the reference's cpuTiming() frame loop
// (cudaBenchMarking.cpp:273-303), calling the reference's own functions in the
// reference's order.  spectrum (optional) receives the 16384-point FFT.
double ref_cpu_frame(short *frame, Complex_t *baseFrameRx0, int size, Complex_t *spectrum, int *raw_out)
{
    const int nValid = ChirpSize * SampleSize;
    const int extendSize = nextPow2(nValid);
    Complex_t *reshaped = (Complex_t *)malloc(sizeof(Complex_t) * SampleSize * ChirpSize * RxSize);
    Complex_t *fftBuf = (Complex_t *)malloc(sizeof(Complex_t) * extendSize);
    ReshapeComplex_t(frame, reshaped, size);
    for (int i = 0; i < nValid; i++)
        fftBuf[i] = Complex_t_SUB(reshaped[i], baseFrameRx0[i]);
    for (int i = nValid; i < extendSize; i++)
    {
        fftBuf[i].real = 0;
        fftBuf[i].imag = 0;
    }
    butterfly_fft(extendSize, fftBuf);
    double Fs_extend = Fs * extendSize / (ChirpSize * SampleSize);
    int raw = FindAbsMax(fftBuf, floor(0.4 * extendSize));
    int maxDisIdx = raw * (ChirpSize * SampleSize) / extendSize;
    double maxDis = 3.0e8 * (((double)maxDisIdx / extendSize) * Fs_extend) / (2 * mu);
    if (spectrum)
        memcpy(spectrum, fftBuf, sizeof(Complex_t) * extendSize);
    if (raw_out)
        *raw_out = raw;
    free(reshaped);
    free(fftBuf);
    return maxDis;
}

// Times n_frames iterations of the reference frame loop body on one thread
// (the reference has no threading); returns seconds.
double ref_cpu_time_frames(short *frames, int n_frames, Complex_t *baseFrameRx0, int size, double *checksum)
{
    Timer t;
    double acc = 0;
    double t0 = t.elapsed();
    for (int f = 0; f < n_frames; f++)
        acc += ref_cpu_frame(frames + (long)f * size, baseFrameRx0, size, NULL, NULL);
    double t1 = t.elapsed();
    if (checksum)
        *checksum = acc;
    return t1 - t0;
}

}
