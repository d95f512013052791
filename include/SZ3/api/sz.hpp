// SZ3/api/sz.hpp -- drop-in replacement header (sz3_b200): the SZ_compress / SZ_decompress templates of the reference
// (include/SZ3/api/sz.hpp:43-82, 94-102, 117-157, 172-177) forwarding to the CUDA library through its C ABI
// (include/sz3b.h).  Same signatures, same ownership rules (new[] buffers), same exception types:
//   SZ3B_E_INVALID_ARGUMENT -> std::invalid_argument     SZ3B_E_RUNTIME -> std::runtime_error
//   SZ3B_E_CUDA / SZ3B_E_UNSUPPORTED -> std::runtime_error (the reference has no such state: there is no CPU path
//   behind this header, a missing GPU or a configuration outside the GPU path is reported, never silently replaced)
// T is float or double on the GPU path.  `data` may also be a CUDA device pointer: see SZ_compress_device below.
#ifndef SZ3_SZ_HPP
#define SZ3_SZ_HPP

#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "SZ3/utils/Config.hpp"
#include "SZ3/version.hpp"
#include "sz3b.h"

namespace SZ3 {
namespace b200 {
// SZ_FLOAT ... SZ_INT64 (def.hpp) of the element type.  The templates instantiate for every arithmetic type the
// reference's callers use (tools/H5Z-SZ3/src/H5Z_SZ3.cpp:195-225 switches over ten of them); the GPU path is built for
// float, double, int32_t and int64_t, the others are refused at run time (SZ3B_E_UNSUPPORTED -> std::runtime_error).
template <class T>
constexpr int dtype_of() {
    static_assert(std::is_arithmetic<T>::value, "sz3_b200: SZ_compress / SZ_decompress take arithmetic element types");
    return std::is_same<T, float>::value    ? 0
           : std::is_same<T, double>::value ? 1
           : std::is_integral<T>::value
               ? (sizeof(T) == 1 ? (std::is_signed<T>::value ? 3 : 2)
                  : sizeof(T) == 2 ? (std::is_signed<T>::value ? 5 : 4)
                  : sizeof(T) == 4 ? (std::is_signed<T>::value ? 7 : 6)
                                   : (std::is_signed<T>::value ? 9 : 8))
               : -1;
}
inline void raise(int rc) {
    if (rc == SZ3B_OK) return;
    const std::string msg = sz3b_last_error();
    if (rc == SZ3B_E_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
// OpenMP analogue: conf.openmp = true emits the reference's slab container; the slab count plays the role of
// omp_get_max_threads() (api/impl/SZImplOMP.hpp:29-35) and can be set here.  Default: one slab per GPU the call will
// use (sz3b_get_device_fanout(): every visible device, driven from inside the one SZ_compress call; a single slab --
// no loss of ratio -- on a one-GPU machine or under a one-process-per-GPU launcher).
inline int &omp_slabs() {
    static int n = 0;
    return n;
}
inline int omp_slab_count() { return omp_slabs() > 0 ? omp_slabs() : sz3b_get_device_fanout(); }
}  // namespace b200

template <class T>
size_t SZ_compress_size_bound(const Config &conf) {
    const sz3b_config p = conf.to_pod(b200::omp_slab_count());
    return sz3b_compress_bound(b200::dtype_of<T>(), &p);
}
}  // namespace SZ3

// data_loc: SZ3B_HOST or SZ3B_DEVICE (zero-copy: `data` is a device pointer on the current CUDA device)
template <class T>
size_t SZ_compress_located(const SZ3::Config &config, const T *data, int data_loc, char *cmpData, size_t cmpCap) {
    using namespace SZ3;
    if (config.N > 4) throw std::invalid_argument("Data dimension higher than 4 is not supported.");
    const sz3b_config p = config.to_pod(b200::omp_slab_count());
    if (cmpCap < sz3b_compress_bound(b200::dtype_of<T>(), &p)) throw std::invalid_argument(SZ3_ERROR_COMP_BUFFER_NOT_LARGE_ENOUGH);
    size_t cmpSize = 0;
    b200::raise(sz3b_compress(b200::dtype_of<T>(), &p, data, data_loc, cmpData, cmpCap, &cmpSize, nullptr));
    return cmpSize;
}

template <class T>
size_t SZ_compress(const SZ3::Config &config, const T *data, char *cmpData, size_t cmpCap) {
    return SZ_compress_located<T>(config, data, SZ3B_HOST, cmpData, cmpCap);
}

template <class T>
char *SZ_compress(const SZ3::Config &config, const T *data, size_t &cmpSize) {
    const size_t cap = SZ3::SZ_compress_size_bound<T>(config);
    char *buffer = new char[cap];
    try {
        cmpSize = SZ_compress<T>(config, data, buffer, cap);
    } catch (...) {
        delete[] buffer;
        throw;
    }
    return buffer;
}

template <class T>
size_t SZ_compress_device(const SZ3::Config &config, const T *device_data, char *cmpData, size_t cmpCap) {
    return SZ_compress_located<T>(config, device_data, SZ3B_DEVICE, cmpData, cmpCap);
}

template <class T>
void SZ_decompress(SZ3::Config &config, const char *cmpData, size_t cmpSize, T *&decData) {
    using namespace SZ3;
    sz3b_config p;
    b200::raise(sz3b_peek_config(cmpData, cmpSize, &p));   // magic / version / Config blob checks
    std::memcpy(&config.sz3MagicNumber, cmpData, 4);
    std::memcpy(&config.sz3DataVer, cmpData + 4, 4);
    config.from_pod(p);
    const bool own = decData == nullptr;
    if (own) decData = new T[config.num];
    try {
        b200::raise(sz3b_decompress(b200::dtype_of<T>(), cmpData, cmpSize, decData, SZ3B_HOST, &p));
    } catch (...) {
        if (own) {
            delete[] decData;
            decData = nullptr;
        }
        throw;
    }
    config.from_pod(p);
}

template <class T>
T *SZ_decompress(SZ3::Config &config, const char *cmpData, size_t cmpSize) {
    T *decData = nullptr;
    SZ_decompress<T>(config, cmpData, cmpSize, decData);
    return decData;
}

#endif
