/*
 * algoGPU.h -- the reference's GPU boundary, re-declared for libswb200.so.
 *
 * Same three extern "C" prototypes as the reference header (algoGPU.h:5, :7, :9), which
 * TestFileWithGPU.cpp:82,87,92 calls, plus SmithDiagonalGPU (SmithDiagonalGPUrefactored.cu:174),
 * which the reference exports but never declares.  All four compute the same thing in the
 * reference -- the Gotoh local-alignment score with MATCH=1, MISMATCH=-1, G_INIT=1, G_EXT=1
 * (simpleGPU.cu:20-23, cudaLazy.cu:11-14, cudaSmithM.cu:77-80, SmithDiagonalGPU.cu:14-17) -- by
 * four copies of a one-launch-per-anti-diagonal kernel; here all four are the one sm_100a wavefront
 * engine.  Arguments: HOST pointers to raw bytes (not NUL-terminated), explicit lengths; the
 * buffers are only read.  Return: the score (>= 0).  There is no error channel in this signature,
 * so on a CUDA failure these print to stderr and abort() rather than return a wrong number.
 */
#ifndef SWB200_ALGOGPU_H
#define SWB200_ALGOGPU_H

#if defined(__GNUC__)
#define SWB200_API __attribute__((visibility("default")))
#else
#define SWB200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* replaces simpleGPU.cu:109-163 */
SWB200_API int SequentialSmithWatermanScoreGPU(unsigned char* seq1, unsigned char* seq2, int len1, int len2);

/* replaces cudaLazy.cu:58-99 */
SWB200_API int SmithWatermanLazyGPU(const unsigned char* seq1, const unsigned char* seq2, int n, int m);

/* replaces cudaSmithM.cu:128-189 */
SWB200_API int SmithWatermanScoreCUDA(const unsigned char* seq1, const unsigned char* seq2, int n, int m);

/* replaces SmithDiagonalGPUrefactored.cu:174-230 (linear-gap kernel; equal to the affine score
 * because G_INIT == G_EXT there) */
SWB200_API int SmithDiagonalGPU(unsigned char* seq1, unsigned char* seq2, int n, int m);

#ifdef __cplusplus
}
#endif
#endif
