/* Minimal declaration shim for the system libzstd runtime (libzstd.so.1, v1.5.5).
 *
 * The image ships the zstd shared object but not its development header.  The SZ3
 * hot path only ever calls the four functions below (reference:
 * include/SZ3/lossless/Lossless_zstd.hpp:32,35,44; include/SZ3/api/impl/SZImpl.hpp:42),
 * so declaring them is enough to build both the reference oracle and the product.
 * Link with  -l:libzstd.so.1 .
 *
 * TEST/BUILD INFRASTRUCTURE: used by oracle/ builds only; the product carries its own
 * copy of these prototypes in sz3_b200/csrc/zstd_decl.h.
 */
#ifndef ORACLE_ZSTD_SHIM_H
#define ORACLE_ZSTD_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
size_t ZSTD_compressBound(size_t srcSize);
size_t ZSTD_compress(void *dst, size_t dstCapacity, const void *src, size_t srcSize, int compressionLevel);
size_t ZSTD_decompress(void *dst, size_t dstCapacity, const void *src, size_t compressedSize);
unsigned ZSTD_isError(size_t code);
const char *ZSTD_getErrorName(size_t code);
unsigned ZSTD_versionNumber(void);
#ifdef __cplusplus
}
#endif
#endif
