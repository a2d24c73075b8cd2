/* b200_common.h — status codes and shared helpers of the C-ABI boundary.
 *
 * The reference (kyle-wedgwood/ArmadilloCUDALinearInterpolation) has no C layer:
 * Armadillo types and CUDA live in one nvcc translation unit (EventDrivenMap.cu:5-9)
 * and every CUDA/cuRAND failure ends in fprintf(stderr)+exit(-1)
 * (EventDrivenMap.cu:18-54).  This boundary replaces that convention: every entry
 * point returns an int status (0 = success, <0 = error), never exits, and the text
 * of the last error of the calling thread is available from b200_last_error().
 *
 * All pointers are plain host pointers unless the function name ends in `_dev`,
 * in which case data pointers are device pointers valid on the current CUDA device
 * and `stream` is a cudaStream_t passed as void* (NULL = default stream).
 */
#ifndef B200_COMMON_H
#define B200_COMMON_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK                 0
#define B200_ERR_INVALID_ARG   -1   /* NULL pointer, size 0 where >0 is required, bad enum   */
#define B200_ERR_CUDA          -2   /* a CUDA runtime call failed; see b200_last_error()      */
#define B200_ERR_NOT_SORTED    -3   /* grid knots are not strictly ascending                   */
#define B200_ERR_TOO_SMALL     -4   /* fewer than two knots (arma: "at least two unique")      */
#define B200_ERR_NO_DEVICE     -5   /* no CUDA device / wrong architecture: there is NO CPU path */
#define B200_ERR_UNSUPPORTED   -6   /* shape outside what the kernels are built for            */
#define B200_ERR_NONFINITE     -7   /* NaN in grid knots                                       */

/* Text of the most recent error raised on the calling thread ("" if none). */
const char* b200_last_error(void);

/* Library/ABI version: (major<<16)|(minor<<8)|patch. */
int b200_version(void);

/* Number of kernels this library has launched so far in this process (every launch site counts itself):
 * the bench reports the difference over its timed region as `gpu_launches`. */
unsigned long long b200_launch_count(void);

/* Number of visible CUDA devices (<=0: none — every compute entry point then fails
 * with B200_ERR_NO_DEVICE; nothing in this library computes on the host). */
int b200_device_count(void);

/* Select the CUDA device used by subsequent calls of this thread. */
int b200_set_device(int device);

/* Pinned (page-locked) host buffers for the host-pointer entry points: using them for
 * queries/outputs lets the library overlap H2D, kernel and D2H in chunks. */
int b200_host_alloc(void** ptr, size_t bytes);
int b200_host_free(void* ptr);

/* Block until all work queued by this library on the current device is complete. */
int b200_synchronize(void);

/* ---- bench introspection: machine ceilings measured live (not part of the hot path) ----
 * Independent 32-byte record gathers from a table of `table_bytes` (choose >> L2): time of the
 * best of 3 timed passes and gathers per second.  On B200 each miss fills a 128-byte line. */
int b200_bench_random_gather(size_t table_bytes, size_t n_gathers, double* ms_out, double* gathers_per_s);
/* Dense FP64 FMA issue ceiling of the current device, in TFLOP/s (2 flops per FMA). */
int b200_bench_fp64_fma(double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* B200_COMMON_H */
