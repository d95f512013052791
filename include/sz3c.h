/* sz3c.h -- C interface of libSZ3c (sz3_b200 build): same symbols, argument meaning and memory ownership as the
 * reference shim (tools/sz3c/include/sz3c.h:52-59, tools/sz3c/src/sz3c.cpp:11-101), backed by the CUDA library.
 * r1 is the fastest-varying dimension; unused dimensions are 0; five dimensions fold r5*r4.
 * Buffers returned by SZ_compress_args / SZ_decompress are malloc()ed; release them with free_buf(). */
#ifndef SZ3_SZ3C_H
#define SZ3_SZ3C_H
#include <stddef.h>
#include <stdio.h>

/* error bound modes (SZ2 numbering, not SZ3::EB) */
#define ABS 0
#define REL 1
#define VR_REL 1
#define ABS_AND_REL 2
#define ABS_OR_REL 3
#define PSNR 4
#define NORM 5
#define PW_REL 10
#define ABS_AND_PW_REL 11
#define ABS_OR_PW_REL 12
#define REL_AND_PW_REL 13
#define REL_OR_PW_REL 14

/* data types */
#ifndef SZ_FLOAT
#define SZ_FLOAT 0
#define SZ_DOUBLE 1
#define SZ_UINT8 2
#define SZ_INT8 3
#define SZ_UINT16 4
#define SZ_INT16 5
#define SZ_UINT32 6
#define SZ_INT32 7
#define SZ_UINT64 8
#define SZ_INT64 9
#endif

#define SZ3C_API

#ifdef __cplusplus
extern "C" {
#endif
SZ3C_API unsigned char *SZ_compress_args(int dataType, void *data, size_t *outSize, int errBoundMode, double absErrBound,
                                         double relBoundRatio, double pwrBoundRatio, size_t r5, size_t r4, size_t r3,
                                         size_t r2, size_t r1);
SZ3C_API void *SZ_decompress(int dataType, unsigned char *bytes, size_t byteLength, size_t r5, size_t r4, size_t r3,
                             size_t r2, size_t r1);
SZ3C_API void free_buf(void *p);
#ifdef __cplusplus
}
#endif
#endif
