// sz3_b200/csrc/misc_kernels.cu -- small data-proportional helpers around the hot kernels:
//   k_minmax          data_range of calAbsErrorBound (reference include/SZ3/utils/Statistic.hpp:12-20)
//   k_profile_blocks  profiling_block of the tuner (include/SZ3/utils/Sample.hpp:9-127)
//   k_gather_cubes    sample_blocks / sampleBlocks copy (Sample.hpp:130-200)
//   k_widen           uint16/uint32 index stream -> int32 for the stage-level test entry points
#include <cuda_runtime.h>

#include "launch.hpp"

namespace sz3b {

template <class T>
__device__ __forceinline__ T dmin(T a, T b) { return b < a ? b : a; }
template <class T>
__device__ __forceinline__ T dmax(T a, T b) { return a < b ? b : a; }

template <class T>
__device__ __forceinline__ void atomic_min_f(T *addr, T v);
template <class T>
__device__ __forceinline__ void atomic_max_f(T *addr, T v);

// ordered-int tricks are unnecessary: a CAS loop on the bit pattern is exact and runs once per warp
template <>
__device__ __forceinline__ void atomic_min_f<float>(float *addr, float v) {
    unsigned *a = reinterpret_cast<unsigned *>(addr);
    unsigned old = *a;
    while (v < __uint_as_float(old)) {
        unsigned assumed = old;
        old = atomicCAS(a, assumed, __float_as_uint(v));
        if (old == assumed) break;
    }
}
template <>
__device__ __forceinline__ void atomic_max_f<float>(float *addr, float v) {
    unsigned *a = reinterpret_cast<unsigned *>(addr);
    unsigned old = *a;
    while (__uint_as_float(old) < v) {
        unsigned assumed = old;
        old = atomicCAS(a, assumed, __float_as_uint(v));
        if (old == assumed) break;
    }
}
template <>
__device__ __forceinline__ void atomic_min_f<double>(double *addr, double v) {
    unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
    unsigned long long old = *a;
    while (v < __longlong_as_double(old)) {
        unsigned long long assumed = old;
        old = atomicCAS(a, assumed, static_cast<unsigned long long>(__double_as_longlong(v)));
        if (old == assumed) break;
    }
}
template <>
__device__ __forceinline__ void atomic_max_f<double>(double *addr, double v) {
    unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double(old) < v) {
        unsigned long long assumed = old;
        old = atomicCAS(a, assumed, static_cast<unsigned long long>(__double_as_longlong(v)));
        if (old == assumed) break;
    }
}

// integer element types: the hardware's own signed atomics
template <>
__device__ __forceinline__ void atomic_min_f<int32_t>(int32_t *addr, int32_t v) { atomicMin(addr, v); }
template <>
__device__ __forceinline__ void atomic_max_f<int32_t>(int32_t *addr, int32_t v) { atomicMax(addr, v); }
template <>
__device__ __forceinline__ void atomic_min_f<int64_t>(int64_t *addr, int64_t v) {
    atomicMin(reinterpret_cast<long long *>(addr), static_cast<long long>(v));
}
template <>
__device__ __forceinline__ void atomic_max_f<int64_t>(int64_t *addr, int64_t v) {
    atomicMax(reinterpret_cast<long long *>(addr), static_cast<long long>(v));
}

// mm[0], mm[1] must be pre-set to data[0] (k_minmax_init)
template <class T>
__global__ void k_minmax_init(const T *__restrict__ data, T *__restrict__ mm) {
    mm[0] = data[0];
    mm[1] = data[0];
}

template <class T>
__global__ void __launch_bounds__(256) k_minmax(const T *__restrict__ data, uint64_t n, T *__restrict__ mm) {
    T lo = data[0], hi = data[0];
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        T v = data[i];
        lo = dmin(lo, v);   // NaN never replaces lo/hi, as in the reference's comparisons
        hi = dmax(hi, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = dmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = dmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomic_min_f<T>(&mm[0], lo);
        atomic_max_f<T>(&mm[1], hi);
    }
}

template <class T>
void launch_minmax(const T *data, uint64_t n, T *mm, cudaStream_t st) {
    k_minmax_init<T><<<1, 1, 0, st>>>(data, mm);
    uint64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks == 0) blocks = 1;
    k_minmax<T><<<static_cast<unsigned>(blocks), 256, 0, st>>>(data, n, mm);
}

struct Dims4 {
    uint32_t d[4];
};

// One warp per sampled block: the lanes share the block's (bs / pstride + 1)^N probe points, min / max by shuffles.
// Every lane starts from the block's first point like the reference's loop does, so a NaN there poisons the result
// the same way; NaNs elsewhere are skipped by both (all comparisons false).
template <class T>
__global__ void __launch_bounds__(128) k_profile_blocks(const T *__restrict__ data, int N, Dims4 dims, uint32_t bs,
                                                        uint32_t pstride, double abs_eb, uint8_t *__restrict__ flags,
                                                        uint64_t nblocks) {
    const uint64_t b = static_cast<uint64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= nblocks) return;
    uint32_t cb[4] = {1, 1, 1, 1}, start[4] = {0, 0, 0, 0};
    uint64_t stride[4] = {0, 0, 0, 0};
    uint64_t acc = 1;
    for (int d = N - 1; d >= 0; d--) {
        stride[d] = acc;
        acc *= dims.d[d];
        cb[d] = (dims.d[d] - bs - 1) / bs + 1;
    }
    uint64_t r = b, base = 0;
    for (int d = N - 1; d >= 0; d--) {
        start[d] = static_cast<uint32_t>(r % cb[d]) * bs;
        r /= cb[d];
        base += start[d] * stride[d];
    }
    T mn = data[base], mx = mn;
    const uint32_t np = bs / pstride + 1;  // ii = 0, pstride, ... <= bs
    uint32_t cnt[4] = {1, 1, 1, 1};
    uint32_t total = 1;
    for (int d = 0; d < N; d++) {
        cnt[d] = np;
        total *= np;
    }
    for (uint32_t it = lane; it < total; it += 32) {
        uint32_t rr = it;
        uint64_t off = base;
        for (int d = N - 1; d >= 0; d--) {
            off += static_cast<uint64_t>(rr % cnt[d]) * pstride * stride[d];
            rr /= cnt[d];
        }
        T v = data[off];
        if (v < mn)
            mn = v;
        else if (v > mx)
            mx = v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const T omn = __shfl_xor_sync(0xffffffffu, mn, o), omx = __shfl_xor_sync(0xffffffffu, mx, o);
        if (omn < mn) mn = omn;
        if (omx > mx) mx = omx;
    }
    if (lane == 0) flags[b] = static_cast<double>(static_cast<T>(mx - mn)) > abs_eb ? 1 : 0;
}

template <class T>
void launch_profile_blocks(const T *data, int N, const uint32_t *dims, uint32_t block, uint32_t pstride, double abs_eb,
                           uint8_t *flags, uint64_t nblocks, cudaStream_t st) {
    if (nblocks == 0) return;
    Dims4 d4;
    for (int i = 0; i < 4; i++) d4.d[i] = i < N ? dims[i] : 1;
    k_profile_blocks<T><<<static_cast<unsigned>((nblocks + 3) / 4), 128, 0, st>>>(data, N, d4, block, pstride, abs_eb, flags,
                                                                                  nblocks);
}

template <class T>
__global__ void __launch_bounds__(256) k_gather_cubes(const T *__restrict__ data, int N, Dims4 dims, uint32_t edge,
                                                      const uint64_t *__restrict__ starts, uint64_t per_cube,
                                                      T *__restrict__ out) {
    const uint32_t k = blockIdx.y;
    const uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= per_cube) return;
    uint64_t r = e, off = starts[k], acc = 1;
    for (int d = N - 1; d >= 0; d--) {
        off += (r % edge) * acc;
        r /= edge;
        acc *= dims.d[d];
    }
    out[static_cast<uint64_t>(k) * per_cube + e] = data[off];
}

template <class T>
void launch_gather_cubes(const T *data, int N, const uint32_t *dims, uint32_t edge, const uint64_t *starts,
                         uint32_t ncubes, T *out, cudaStream_t st) {
    if (ncubes == 0) return;
    Dims4 d4;
    uint64_t per = 1;
    for (int i = 0; i < 4; i++) d4.d[i] = i < N ? dims[i] : 1;
    for (int i = 0; i < N; i++) per *= edge;
    dim3 grid(static_cast<unsigned>((per + 255) / 256), ncubes);
    k_gather_cubes<T><<<grid, 256, 0, st>>>(data, N, d4, edge, starts, per, out);
}

template <class QT>
__global__ void __launch_bounds__(256) k_widen(const QT *__restrict__ q, uint64_t n, int32_t *__restrict__ out) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = static_cast<int32_t>(q[i]);
}

template <class QT>
void launch_widen(const QT *q, uint64_t n, int32_t *out, cudaStream_t st) {
    if (n == 0) return;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    k_widen<QT><<<static_cast<unsigned>(blocks), 256, 0, st>>>(q, n, out);
}

template void launch_minmax<float>(const float *, uint64_t, float *, cudaStream_t);
template void launch_minmax<double>(const double *, uint64_t, double *, cudaStream_t);
template void launch_profile_blocks<float>(const float *, int, const uint32_t *, uint32_t, uint32_t, double, uint8_t *,
                                           uint64_t, cudaStream_t);
template void launch_profile_blocks<double>(const double *, int, const uint32_t *, uint32_t, uint32_t, double,
                                            uint8_t *, uint64_t, cudaStream_t);
template void launch_gather_cubes<float>(const float *, int, const uint32_t *, uint32_t, const uint64_t *, uint32_t,
                                         float *, cudaStream_t);
template void launch_gather_cubes<double>(const double *, int, const uint32_t *, uint32_t, const uint64_t *, uint32_t,
                                          double *, cudaStream_t);
#define SZ3B_INST_MISC_INT(T)                                                                                          \
    template void launch_minmax<T>(const T *, uint64_t, T *, cudaStream_t);                                            \
    template void launch_profile_blocks<T>(const T *, int, const uint32_t *, uint32_t, uint32_t, double, uint8_t *,    \
                                           uint64_t, cudaStream_t);                                                    \
    template void launch_gather_cubes<T>(const T *, int, const uint32_t *, uint32_t, const uint64_t *, uint32_t, T *, \
                                         cudaStream_t);
SZ3B_INST_MISC_INT(int32_t)
SZ3B_INST_MISC_INT(int64_t)
template void launch_widen<uint16_t>(const uint16_t *, uint64_t, int32_t *, cudaStream_t);
template void launch_widen<uint32_t>(const uint32_t *, uint64_t, int32_t *, cudaStream_t);

// Small upload without the copy engine: the source is pinned (device-mapped) host memory read by the SMs.  Used while
// the bulk input copy owns the H2D engine, where a cudaMemcpyAsync would queue behind it.
__global__ void __launch_bounds__(256) k_upload_bytes(unsigned char *__restrict__ dst, const unsigned char *__restrict__ src,
                                                      size_t bytes) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t nvec = bytes / 16;
    if (i < nvec) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
    if (i < bytes - nvec * 16) dst[nvec * 16 + i] = src[nvec * 16 + i];
}
void launch_upload_bytes(void *dst, const void *src_mapped, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return;
    const size_t n = bytes / 16 > 16 ? bytes / 16 : 16;
    k_upload_bytes<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(static_cast<unsigned char *>(dst),
                                                                          static_cast<const unsigned char *>(src_mapped), bytes);
}

}  // namespace sz3b
