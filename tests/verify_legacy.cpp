// verify_legacy.cpp — the verification step the reference left commented out (cudaBenchMarking.cpp:405-419):
// every frame of a capture goes through the CPU path and through cudaProcessing(), and the two distances must agree
// to 1e-5.  TEST PROGRAM: the CPU side is the oracle (oracle/mmw_oracle.h), the GPU side is the drop-in symbol of
// libmmw_radar_b200.so called exactly as cudaTiming() calls it (cudaBenchMarking.cpp:339-378).
//
//   verify_legacy <capture.bin>      exit 0 = every frame verified
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mmw_legacy.h"
#include "mmw_oracle.h"

int main(int argc, char **argv)
{
    const int SampleSize = 100, ChirpSize = 128, RxSize = 4;                    // cudaBenchMarking.cpp:3-6
    const int NumDataPerFrame = ChirpSize * SampleSize * RxSize * 2;
    if (argc < 2) { fprintf(stderr, "usage: %s capture.bin\n", argv[0]); return 2; }
    FILE *fp = fopen(argv[1], "rb");
    if (fp == NULL) { printf("unable to read the specified file\n"); return 2; }
    short *inputData = (short *)malloc(NumDataPerFrame * sizeof(short));
    int size = (int)fread(inputData, sizeof(short), NumDataPerFrame, fp);
    // base frame: rx0 of frame 0 in [chirp][sample] order (ReshapeComplex_t + memmove, :357-365)
    orc_cx *reshaped = (orc_cx *)calloc((size_t)NumDataPerFrame / 2, sizeof(orc_cx));
    orc_reshape(inputData, reshaped, size, SampleSize, ChirpSize, RxSize);
    Complex_t *baseFrameRx0 = (Complex_t *)malloc(ChirpSize * SampleSize * sizeof(Complex_t));
    memcpy(baseFrameRx0, reshaped, ChirpSize * SampleSize * sizeof(Complex_t));

    double fftTime = 0, preProcessTime = 0, findMaxTime = 0, totalTime = 0;
    int numFrameRead = 0, bad = 0;
    while ((size = (int)fread(inputData, sizeof(short), NumDataPerFrame, fp)) > 0) {
        numFrameRead++;
        const double cpuRes = orc_legacy_frame(inputData, (const orc_cx *)baseFrameRx0, size, SampleSize, ChirpSize, RxSize, NULL, NULL);
        const double cudaRes = cudaProcessing(inputData, baseFrameRx0, size, &fftTime, &preProcessTime, &findMaxTime, &totalTime);
        if (fabs(cudaRes - cpuRes) >= 1e-5) {                                    // the reference's tolerance (:412)
            printf("CUDA result verification failed at frame %d\n", numFrameRead);
            printf("Ref Res %.6f CUDA res %.6f\n", cpuRes, cudaRes);
            bad++;
        }
    }
    fclose(fp);
    printf("verified %d frames, %d mismatches, cuda inner time %.5f ms average %.5f ms/frame\n", numFrameRead, bad,
           1000.0 * totalTime, numFrameRead ? 1000.0 * totalTime / numFrameRead : 0.0);
    free(reshaped); free(baseFrameRx0); free(inputData);
    return bad ? 1 : (numFrameRead > 0 ? 0 : 2);
}
