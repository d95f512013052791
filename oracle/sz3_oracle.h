/* oracle/sz3_oracle.h -- C ABI of the plain-C CPU restatement of SZ3's predict->quantize->encode path.
 *
 * TEST INFRASTRUCTURE ONLY (see header of sz3_oracle.c).  The entry points deliberately mirror the ones
 * oracle/ref_driver.cpp exports for the real reference so tests can diff them one to one.
 */
#ifndef SZ3_ORACLE_H
#define SZ3_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* POD mirror of SZ3::Config (reference include/SZ3/utils/Config.hpp:441-478). */
typedef struct orc_config {
    int32_t N;
    uint64_t dims[4];
    int32_t cmprAlgo;        /* ALGO enum, Config.hpp:68 */
    int32_t errorBoundMode;  /* EB enum, Config.hpp:54 */
    double absErrorBound;
    double relErrorBound;
    double psnrErrorBound;
    double l2normErrorBound;
    int32_t openmp;          /* for the oracle: >0 means "emit the OMP container with this many slabs" */
    int32_t quantbinCnt;
    int32_t blockSize;       /* <=0 -> default for N (Config.hpp:175) */
    int32_t lorenzo;
    int32_t lorenzo2;
    int32_t regression;
    int32_t regression2;
    int32_t interpAlgo;
    int32_t interpDirection;
    int32_t interpAnchorStride;
    double interpAlpha;
    double interpBeta;
} orc_config;

enum { ORC_EB_ABS, ORC_EB_REL, ORC_EB_PSNR, ORC_EB_L2NORM, ORC_EB_ABS_AND_REL, ORC_EB_ABS_OR_REL };
enum { ORC_ALGO_LORENZO_REG, ORC_ALGO_INTERP_LORENZO, ORC_ALGO_INTERP, ORC_ALGO_NOPRED, ORC_ALGO_LOSSLESS };

void orc_config_default(orc_config *c, int N, const uint64_t *dims);

size_t orc_size_bound(int dtype, const orc_config *c);
/* dtype 0 = float32, 1 = float64.  Return bytes written, or -1 on error. */
long long orc_compress(int dtype, const orc_config *c, const void *data, char *out, size_t cap);
int orc_decompress(int dtype, const char *cmp, size_t n, void *out, orc_config *conf_out);

long long orc_interp_decompose(int dtype, const orc_config *c, double eb, void *data, int *quant, unsigned char *blob,
                               size_t *blob_len);
long long orc_blockwise_decompose(int dtype, const orc_config *c, double eb, void *data, int *quant,
                                  unsigned char *blob, size_t *blob_len);
long long orc_huffman_encode(const int *q, size_t n, unsigned char *out, size_t *tree_len);
long long orc_huffman_decode(const unsigned char *in, size_t in_len, size_t n, int *out);
int orc_tune(int dtype, orc_config *c, const void *data);
double orc_abs_eb(int dtype, const orc_config *c, const void *data);
size_t orc_config_save(const orc_config *c, unsigned char *out);

#ifdef __cplusplus
}
#endif
#endif
