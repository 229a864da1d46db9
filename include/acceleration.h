/*
 * acceleration.h — interface-compatible stand-in for the reference's header of the same name.
 *
 * WHY THIS FILE EXISTS.  The reference's caller (cudaBenchMarking.cpp) does `#include "acceleration.h"` and needs three
 * things from it: the `Complex_t` layout, the `Timer` helper it times its loops with, and the prototype of
 * `cudaProcessing`.  A build tree that has this library but not the reference checkout can compile that caller, unmodified,
 * against THIS header (-I<repo>/include) and link it with libmmw_radar_b200.so.  Nothing here is copied from the
 * reference: the declarations are restated from its ABI (citations below) and the Timer is written afresh.  Where the
 * reference checkout is present, its own header works just as well — both define the include guard ACCELERATION_H, and
 * mmw_legacy.h yields to whichever came first.
 *
 * ABI facts this header must (and does) reproduce — checked by tests/test_abi.py and tests/test_gpu_dropin.py:
 *   - struct Complex_t: two doubles, `real` then `imag`, 16 bytes, no padding            (reference acceleration.h:27-30)
 *   - double cudaProcessing(short*, Complex_t*, int, double*, double*, double*, double*) with C++ linkage, i.e. the
 *     symbol _Z14cudaProcessingPsP9Complex_tiPdS2_S2_S2_                                   (reference acceleration.h:32)
 *   - class Timer: default-constructible, `void reset()`, `double elapsed() const` in SECONDS since construction or the
 *     last reset(); the caller uses exactly these (cudaBenchMarking.cpp:214-331, 335-394)  (reference acceleration.h:10-24)
 *   - the C headers the reference's header pulls in for its includer (printf, malloc, memmove, floor, log2 are used by
 *     cudaBenchMarking.cpp without further includes)
 */
#ifndef ACCELERATION_H
#define ACCELERATION_H

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

/* Wall-clock stopwatch, seconds as double.  A monotonic clock is used (an adjustment of the system time cannot make an
 * interval negative); resolution is that of std::chrono::steady_clock (nanoseconds on Linux). */
class Timer {
    using tick = std::chrono::steady_clock;
    tick::time_point start_;

public:
    Timer() { reset(); }
    void reset() { start_ = tick::now(); }
    double elapsed() const
    {
        const std::chrono::duration<double> dt = tick::now() - start_;
        return dt.count();
    }
};

/* fp64 complex sample, the element type of the base frame handed to cudaProcessing */
struct Complex_t {
    double real;
    double imag;
};

/*
 * One frame (100 samples x 128 chirps x 4 receivers, int16 IIQQ, `size` shorts at `deviceIn` — a HOST pointer despite
 * its name) through: receiver 0 minus `host_baseFrame` (12 800 values), zero-pad to 16 384, forward FFT, arg-max of |X|
 * over the first 40 % of the bins, converted to metres.  The four timers are accumulated with += in seconds.
 * Implemented by libmmw_radar_b200.so (csrc/mmw_legacy.cu); full contract in mmw_legacy.h.
 */
double cudaProcessing(short *deviceIn, Complex_t *host_baseFrame, int size,
                      double *fftTime, double *preProcessTime, double *findMaxTime, double *totalTime);

#endif /* ACCELERATION_H */
