/* sz3_b200/csrc/zstd_decl.h -- prototypes of the libzstd entry points the host tail calls.
 *
 * The image ships the zstd runtime (libzstd.so.1, v1.5.5) without its development header; the stream format only
 * needs these functions (reference include/SZ3/lossless/Lossless_zstd.hpp:32,35,44).  Link with -l:libzstd.so.1.
 */
#ifndef SZ3B_ZSTD_DECL_H
#define SZ3B_ZSTD_DECL_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
size_t ZSTD_compressBound(size_t srcSize);
size_t ZSTD_compress(void *dst, size_t dstCapacity, const void *src, size_t srcSize, int compressionLevel);
size_t ZSTD_decompress(void *dst, size_t dstCapacity, const void *src, size_t compressedSize);
unsigned ZSTD_isError(size_t code);
size_t ZSTD_findFrameCompressedSize(const void *src, size_t srcSize);
unsigned long long ZSTD_getFrameContentSize(const void *src, size_t srcSize);
typedef struct ZSTD_CCtx_s ZSTD_CCtx;
ZSTD_CCtx *ZSTD_createCCtx(void);
size_t ZSTD_freeCCtx(ZSTD_CCtx *cctx);
size_t ZSTD_compressCCtx(ZSTD_CCtx *cctx, void *dst, size_t dstCapacity, const void *src, size_t srcSize,
                         int compressionLevel);
/* streaming decompression (stable API): output becomes available block by block */
typedef struct ZSTD_DCtx_s ZSTD_DCtx;
typedef struct ZSTD_inBuffer_s {
    const void *src;
    size_t size;
    size_t pos;
} ZSTD_inBuffer;
typedef struct ZSTD_outBuffer_s {
    void *dst;
    size_t size;
    size_t pos;
} ZSTD_outBuffer;
ZSTD_DCtx *ZSTD_createDCtx(void);
size_t ZSTD_freeDCtx(ZSTD_DCtx *dctx);
size_t ZSTD_decompressStream(ZSTD_DCtx *zds, ZSTD_outBuffer *output, ZSTD_inBuffer *input);
size_t ZSTD_DCtx_reset(ZSTD_DCtx *dctx, int reset); /* 1 = ZSTD_reset_session_only */
const char *ZSTD_getErrorName(size_t code);
unsigned ZSTD_versionNumber(void);
#ifdef __cplusplus
}
#endif
#endif
