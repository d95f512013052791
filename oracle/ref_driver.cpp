// oracle/ref_driver.cpp -- thin extern "C" driver around the UNMODIFIED reference headers.
//
// TEST INFRASTRUCTURE ONLY.  This file contains no SZ3 algorithm code: it only calls the
// reference's own templates (included from /root/reference/include at build time, see
// oracle/Makefile) and exposes them through a C ABI so that tests/ and bench.py can load
// oracle/_ref/libsz3ref.so with ctypes.  It is how the C restatement in oracle/sz3_oracle.c
// and the CUDA product are pinned to the real reference behaviour.
//
// Entry points:
//   ref_compress / ref_decompress / ref_size_bound      -> SZ_compress / SZ_decompress (api/sz.hpp:43,117)
//   ref_interp_decompose                                 -> InterpolationDecomposition::compress + save
//   ref_blockwise_decompose                              -> BlockwiseDecomposition::compress + save
//   ref_huffman_encode / ref_huffman_decode              -> HuffmanEncoder<int>
//   ref_tune                                             -> the tuner inside SZ_compress_Interp_lorenzo
//   ref_abs_eb                                           -> calAbsErrorBound
#include <cstdint>
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cstring>
#include <vector>

#include "SZ3/api/sz.hpp"

using namespace SZ3;

extern "C" {

// POD mirror of SZ3::Config (utils/Config.hpp:441-478); same layout as sz3b_config in include/sz3b.h.
struct ref_config {
    int32_t N;
    uint64_t dims[4];
    int32_t cmprAlgo;
    int32_t errorBoundMode;
    double absErrorBound;
    double relErrorBound;
    double psnrErrorBound;
    double l2normErrorBound;
    int32_t openmp;
    int32_t quantbinCnt;
    int32_t blockSize;
    int32_t lorenzo;
    int32_t lorenzo2;
    int32_t regression;
    int32_t regression2;
    int32_t interpAlgo;
    int32_t interpDirection;
    int32_t interpAnchorStride;
    double interpAlpha;
    double interpBeta;
};
}

static Config to_conf(const ref_config *c) {
    Config conf;
    std::vector<size_t> dims(c->dims, c->dims + c->N);
    conf.setDims(dims.begin(), dims.end());
    conf.cmprAlgo = static_cast<uint8_t>(c->cmprAlgo);
    conf.errorBoundMode = static_cast<uint8_t>(c->errorBoundMode);
    conf.absErrorBound = c->absErrorBound;
    conf.relErrorBound = c->relErrorBound;
    conf.psnrErrorBound = c->psnrErrorBound;
    conf.l2normErrorBound = c->l2normErrorBound;
    conf.openmp = c->openmp != 0;
    conf.quantbinCnt = c->quantbinCnt;
    if (c->blockSize > 0) conf.blockSize = c->blockSize;
    conf.lorenzo = c->lorenzo != 0;
    conf.lorenzo2 = c->lorenzo2 != 0;
    conf.regression = c->regression != 0;
    conf.regression2 = c->regression2 != 0;
    conf.interpAlgo = static_cast<uint8_t>(c->interpAlgo);
    conf.interpDirection = static_cast<uint8_t>(c->interpDirection);
    conf.interpAnchorStride = c->interpAnchorStride;
    conf.interpAlpha = c->interpAlpha;
    conf.interpBeta = c->interpBeta;
    return conf;
}

static void from_conf(const Config &conf, ref_config *c) {
    c->N = conf.N;
    for (int i = 0; i < 4; i++) c->dims[i] = i < conf.N ? conf.dims[i] : 0;
    c->cmprAlgo = conf.cmprAlgo;
    c->errorBoundMode = conf.errorBoundMode;
    c->absErrorBound = conf.absErrorBound;
    c->relErrorBound = conf.relErrorBound;
    c->psnrErrorBound = conf.psnrErrorBound;
    c->l2normErrorBound = conf.l2normErrorBound;
    c->openmp = conf.openmp;
    c->quantbinCnt = conf.quantbinCnt;
    c->blockSize = conf.blockSize;
    c->lorenzo = conf.lorenzo;
    c->lorenzo2 = conf.lorenzo2;
    c->regression = conf.regression;
    c->regression2 = conf.regression2;
    c->interpAlgo = conf.interpAlgo;
    c->interpDirection = conf.interpDirection;
    c->interpAnchorStride = conf.interpAnchorStride;
    c->interpAlpha = conf.interpAlpha;
    c->interpBeta = conf.interpBeta;
}

template <class T>
static long long compress_t(const ref_config *c, const void *data, char *out, size_t cap) {
    try {
        Config conf = to_conf(c);
        return static_cast<long long>(SZ_compress<T>(conf, static_cast<const T *>(data), out, cap));
    } catch (std::exception &e) {
        fprintf(stderr, "[ref_compress] %s\n", e.what());
        return -1;
    }
}

template <class T>
static int decompress_t(const char *cmp, size_t n, void *out, ref_config *cout_) {
    try {
        Config conf;
        T *p = static_cast<T *>(out);
        SZ_decompress<T>(conf, cmp, n, p);
        if (cout_) from_conf(conf, cout_);
        return 0;
    } catch (std::exception &e) {
        fprintf(stderr, "[ref_decompress] %s\n", e.what());
        return -1;
    }
}

template <class T, uint N>
static long long interp_t(const ref_config *c, double eb, void *data, int *quant, unsigned char *blob,
                          size_t *blob_len) {
    Config conf = to_conf(c);
    conf.absErrorBound = eb;
    auto dec = make_decomposition_interpolation<T, N>(conf, LinearQuantizer<T>(eb, conf.quantbinCnt / 2));
    std::vector<int> q = dec.compress(conf, static_cast<T *>(data));
    memcpy(quant, q.data(), q.size() * sizeof(int));
    unsigned char *p = blob;
    dec.save(p);
    *blob_len = p - blob;
    return static_cast<long long>(q.size());
}

template <class T, uint N>
static long long blockwise_t(const ref_config *c, double eb, void *data, int *quant, unsigned char *blob,
                             size_t *blob_len) {
    Config conf = to_conf(c);
    conf.absErrorBound = eb;
    auto quantizer = LinearQuantizer<T>(eb, conf.quantbinCnt / 2);
    std::vector<std::shared_ptr<concepts::PredictorInterface<T, N>>> predictors;
    int methodCnt = conf.lorenzo + conf.lorenzo2 + conf.regression;
    std::vector<int> q;
    unsigned char *p = blob;
    T *d = static_cast<T *>(data);
    // Mirrors the selection logic of make_compressor_lorenzo_regression (SZAlgoLorenzoReg.hpp:22-64) but stops
    // after the decomposition so that quant indices and the saved side streams can be inspected.
    if (methodCnt == 1 && conf.lorenzo) {
        auto dec = make_decomposition_blockwise<T, N>(conf, LorenzoPredictor<T, N, 1>(eb), quantizer);
        q = dec.compress(conf, d);
        dec.save(p);
    } else if (methodCnt == 1 && conf.lorenzo2) {
        auto dec = make_decomposition_blockwise<T, N>(conf, LorenzoPredictor<T, N, 2>(eb), quantizer);
        q = dec.compress(conf, d);
        dec.save(p);
    } else if (methodCnt == 1 && conf.regression) {
        auto dec = make_decomposition_blockwise<T, N>(conf, RegressionPredictor<T, N>(conf.blockSize, eb), quantizer);
        q = dec.compress(conf, d);
        dec.save(p);
    } else {
        if (conf.lorenzo) predictors.push_back(std::make_shared<LorenzoPredictor<T, N, 1>>(eb));
        if (conf.lorenzo2) predictors.push_back(std::make_shared<LorenzoPredictor<T, N, 2>>(eb));
        if (conf.regression) predictors.push_back(std::make_shared<RegressionPredictor<T, N>>(conf.blockSize, eb));
        auto dec = make_decomposition_blockwise<T, N>(conf, ComposedPredictor<T, N>(predictors), quantizer);
        q = dec.compress(conf, d);
        dec.save(p);
    }
    memcpy(quant, q.data(), q.size() * sizeof(int));
    *blob_len = p - blob;
    return static_cast<long long>(q.size());
}

template <class T, uint N>
static int tune_t(ref_config *c, const void *data) {
    Config conf = to_conf(c);
    std::vector<T> copy(static_cast<const T *>(data), static_cast<const T *>(data) + conf.num);
    size_t cap = SZ_compress_size_bound<T>(conf);
    std::vector<uchar> out(cap);
    conf.cmprAlgo = ALGO_INTERP_LORENZO;
    SZ_compress_Interp_lorenzo<T, N>(conf, copy.data(), out.data(), cap);
    from_conf(conf, c);
    return 0;
}

#define DISPATCH_TN(fn, dtype, N, ...)                      \
    do {                                                    \
        if (dtype == 0) {                                   \
            if (N == 1) return fn<float, 1>(__VA_ARGS__);   \
            if (N == 2) return fn<float, 2>(__VA_ARGS__);   \
            if (N == 3) return fn<float, 3>(__VA_ARGS__);   \
            if (N == 4) return fn<float, 4>(__VA_ARGS__);   \
        } else {                                            \
            if (N == 1) return fn<double, 1>(__VA_ARGS__);  \
            if (N == 2) return fn<double, 2>(__VA_ARGS__);  \
            if (N == 3) return fn<double, 3>(__VA_ARGS__);  \
            if (N == 4) return fn<double, 4>(__VA_ARGS__);  \
        }                                                   \
    } while (0)

extern "C" {

const char *ref_version() { return SZ3_VER; }

// OpenMP team size of the following ref_compress / ref_decompress calls with conf.openmp (SZ_compress_OMP takes
// omp_get_num_threads() as its slab count, SZImplOMP.hpp:26-31).  Launchers such as torchrun export OMP_NUM_THREADS=1;
// the benchmark's reference arm and the container parity tests set the count they mean explicitly.
void ref_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int ref_get_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

unsigned ref_zstd_version() { return ZSTD_versionNumber(); }

size_t ref_size_bound(int dtype, const ref_config *c) {
    Config conf = to_conf(c);
    if (dtype == SZ_INT32) return SZ_compress_size_bound<int32_t>(conf);
    if (dtype == SZ_INT64) return SZ_compress_size_bound<int64_t>(conf);
    return dtype == 0 ? SZ_compress_size_bound<float>(conf) : SZ_compress_size_bound<double>(conf);
}

// dtype: 0 = float, 1 = double, 7 = int32, 9 = int64 (SZ_FLOAT / SZ_DOUBLE / SZ_INT32 / SZ_INT64, Config.hpp:27-36;
// the integer types only through the whole-stream entry points, as tools/sz3/sz3.cpp:458-461 uses them).
// Returns bytes written or -1.
long long ref_compress(int dtype, const ref_config *c, const void *data, char *out, size_t cap) {
    if (dtype == SZ_INT32) return compress_t<int32_t>(c, data, out, cap);
    if (dtype == SZ_INT64) return compress_t<int64_t>(c, data, out, cap);
    return dtype == 0 ? compress_t<float>(c, data, out, cap) : compress_t<double>(c, data, out, cap);
}

int ref_decompress(int dtype, const char *cmp, size_t n, void *out, ref_config *conf_out) {
    if (dtype == SZ_INT32) return decompress_t<int32_t>(cmp, n, out, conf_out);
    if (dtype == SZ_INT64) return decompress_t<int64_t>(cmp, n, out, conf_out);
    return dtype == 0 ? decompress_t<float>(cmp, n, out, conf_out) : decompress_t<double>(cmp, n, out, conf_out);
}

// data is overwritten with the reconstruction (as the reference does).  quant must hold num ints, blob num*8+4096.
long long ref_interp_decompose(int dtype, const ref_config *c, double eb, void *data, int *quant, unsigned char *blob,
                               size_t *blob_len) {
    Config conf = to_conf(c);
    int N = conf.N;
    DISPATCH_TN(interp_t, dtype, N, c, eb, data, quant, blob, blob_len);
    return -1;
}

long long ref_blockwise_decompose(int dtype, const ref_config *c, double eb, void *data, int *quant,
                                  unsigned char *blob, size_t *blob_len) {
    Config conf = to_conf(c);
    int N = conf.N;
    DISPATCH_TN(blockwise_t, dtype, N, c, eb, data, quant, blob, blob_len);
    return -1;
}

// Writes HuffmanEncoder<int>::save followed by ::encode (size_t outSize + bits). Returns total bytes.
long long ref_huffman_encode(const int *q, size_t n, unsigned char *out, size_t *tree_len) {
    try {
        HuffmanEncoder<int> enc;
        enc.preprocess_encode(q, n, 0);
        unsigned char *p = out;
        enc.save(p);
        *tree_len = p - out;
        enc.encode(q, n, p);
        enc.postprocess_encode();
        return p - out;
    } catch (std::exception &e) {
        fprintf(stderr, "[ref_huffman_encode] %s\n", e.what());
        return -1;
    }
}

long long ref_huffman_decode(const unsigned char *in, size_t in_len, size_t n, int *out) {
    HuffmanEncoder<int> enc;
    const unsigned char *p = in;
    size_t rem = in_len;
    enc.load(p, rem);
    auto v = enc.decode(p, n);
    enc.postprocess_decode();
    memcpy(out, v.data(), n * sizeof(int));
    return p - in;
}

// Runs the reference auto-tuner + compression and reports the tuned configuration back in *c.
int ref_tune(int dtype, ref_config *c, const void *data) {
    Config conf = to_conf(c);
    int N = conf.N;
    DISPATCH_TN(tune_t, dtype, N, c, data);
    return -1;
}

double ref_abs_eb(int dtype, const ref_config *c, const void *data) {
    Config conf = to_conf(c);
    if (dtype == SZ_INT32)
        calAbsErrorBound<int32_t>(conf, static_cast<const int32_t *>(data));
    else if (dtype == SZ_INT64)
        calAbsErrorBound<int64_t>(conf, static_cast<const int64_t *>(data));
    else if (dtype == 0)
        calAbsErrorBound<float>(conf, static_cast<const float *>(data));
    else
        calAbsErrorBound<double>(conf, static_cast<const double *>(data));
    return conf.absErrorBound;
}

size_t ref_config_save(const ref_config *c, unsigned char *out) {
    Config conf = to_conf(c);
    unsigned char *p = out;
    return conf.save(p);
}

}  // extern "C"
