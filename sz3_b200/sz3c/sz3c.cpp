// sz3_b200/sz3c/sz3c.cpp -- libSZ3c on top of the drop-in SZ3/api/sz.hpp (CUDA library behind it).
// Mirrors the reference shim tools/sz3c/src/sz3c.cpp:11-101: default algorithm (ALGO_INTERP_LORENZO), no config file,
// only ABS / REL / ABS_AND_REL / ABS_OR_REL accepted (anything else prints and exit(0)s, as the reference does),
// pwrBoundRatio ignored, float and double only, results handed out as malloc() memory.
#include "sz3c.h"

#include <cstdlib>
#include <cstring>

#include "SZ3/api/sz.hpp"

namespace {
SZ3::Config make_conf(size_t r5, size_t r4, size_t r3, size_t r2, size_t r1) {
    if (r2 == 0) return SZ3::Config(r1);
    if (r3 == 0) return SZ3::Config(r2, r1);
    if (r4 == 0) return SZ3::Config(r3, r2, r1);
    if (r5 == 0) return SZ3::Config(r4, r3, r2, r1);
    return SZ3::Config(r5 * r4, r3, r2, r1);
}

template <class T>
unsigned char *compress_as(const SZ3::Config &conf, void *data, size_t *outSize) {
    char *buf = SZ_compress<T>(conf, static_cast<const T *>(data), *outSize);
    unsigned char *out = static_cast<unsigned char *>(malloc(*outSize));
    memcpy(out, buf, *outSize);
    delete[] buf;
    return out;
}
}  // namespace

extern "C" unsigned char *SZ_compress_args(int dataType, void *data, size_t *outSize, int errBoundMode, double absErrBound,
                                           double relBoundRatio, double /*pwrBoundRatio*/, size_t r5, size_t r4, size_t r3,
                                           size_t r2, size_t r1) {
    SZ3::Config conf = make_conf(r5, r4, r3, r2, r1);
    conf.absErrorBound = absErrBound;
    conf.relErrorBound = relBoundRatio;
    switch (errBoundMode) {
        case ABS: conf.errorBoundMode = SZ3::EB_ABS; break;
        case REL: conf.errorBoundMode = SZ3::EB_REL; break;
        case ABS_AND_REL: conf.errorBoundMode = SZ3::EB_ABS_AND_REL; break;
        case ABS_OR_REL: conf.errorBoundMode = SZ3::EB_ABS_OR_REL; break;
        default:
            printf("errBoundMode %d not support\n ", errBoundMode);
            exit(0);
    }
    if (dataType == SZ_FLOAT) return compress_as<float>(conf, data, outSize);
    if (dataType == SZ_DOUBLE) return compress_as<double>(conf, data, outSize);
    printf("dataType %d not support\n", dataType);
    exit(0);
}

extern "C" void *SZ_decompress(int dataType, unsigned char *bytes, size_t byteLength, size_t r5, size_t r4, size_t r3,
                               size_t r2, size_t r1) {
    size_t n = r1;
    if (r2) n *= r2;
    if (r2 && r3) n *= r3;
    if (r2 && r3 && r4) n *= r4;
    if (r2 && r3 && r4 && r5) n *= r5;
    SZ3::Config conf;
    if (dataType == SZ_FLOAT) {
        float *dec = static_cast<float *>(malloc(n * sizeof(float)));
        SZ_decompress<float>(conf, reinterpret_cast<char *>(bytes), byteLength, dec);
        return dec;
    }
    if (dataType == SZ_DOUBLE) {
        double *dec = static_cast<double *>(malloc(n * sizeof(double)));
        SZ_decompress<double>(conf, reinterpret_cast<char *>(bytes), byteLength, dec);
        return dec;
    }
    printf("dataType %d not support\n", dataType);
    exit(0);
}

extern "C" void free_buf(void *p) { free(p); }
