// sz3_b200/csrc/interp_kernels.cu -- __global__ wrappers + launchers of the fused interpolation predict+quantize
// kernels (bodies in interp_body.cuh).  Compiled with -fmad=false: the reference arithmetic has no FMA contraction.
#include <cuda_runtime.h>

#include "device_ctx.cuh"
#include "interp_body.cuh"
#include "interp_line.cuh"
#include "interp_lean.cuh"
#include "launch.hpp"

namespace sz3b {

// ---------------------------------------------------------------------------------------------------------------------
// anchors / first element (reference InterpolationDecomposition.hpp:92-98, 215-222)
// ---------------------------------------------------------------------------------------------------------------------
template <class T, class QT>
__global__ void __launch_bounds__(256) k_interp_anchor(InterpArgs<T, QT> A, uint32_t anchor_stride, uint64_t n_anchor) {
    const uint32_t batch = blockIdx.y;
    const uint64_t gid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gid >= n_anchor) return;
    const InterpShape &sh = A.sh;
    const T *dat = A.data + batch * A.data_bstride;
    if (anchor_stride == 0) {
        // no anchor grid: the first element is quantized against a zero prediction (:92-93)
        T rec;
        int qv = quantize<T>(dat[0], static_cast<T>(0), A.qp, rec);
        uint64_t pos = batch * A.q_bstride;
        A.q[pos] = static_cast<QT>(qv);
        if (qv == 0) A.unpred_tmp[pos] = dat[0];
        if (A.work) A.work[batch * A.data_bstride] = rec;
        if (A.recon2) A.recon2[batch * A.recon2_bstride] = rec;
        atomicAdd(&A.hist[qv], 1ull);
        return;
    }
    uint64_t r = gid, off = 0, off2 = 0;
    bool even = true;
    for (int d = sh.N - 1; d >= 0; d--) {
        uint32_t ext = (sh.dims[d] - 1) / anchor_stride + 1;
        uint32_t x = static_cast<uint32_t>(r % ext) * anchor_stride;
        r /= ext;
        off += x * sh.stride[d];
        off2 += (x >> 1) * A.stride2[d];
        even = even && !(x & 1);
    }
    T v = dat[off];
    uint64_t pos = batch * A.q_bstride + gid;
    A.q[pos] = 0;
    A.unpred_tmp[pos] = v;
    if (A.work) A.work[batch * A.data_bstride + off] = v;
    if (A.recon2 && even) A.recon2[batch * A.recon2_bstride + off2] = v;
    if (gid == 0) atomicAdd(&A.hist[0], static_cast<unsigned long long>(n_anchor));
}

// ---------------------------------------------------------------------------------------------------------------------
// tile schedule, N == 3
// ---------------------------------------------------------------------------------------------------------------------
// line-walker tile schedule (interp_line.cuh): two CTAs per SM; three lanes build the pass tables while the
// rest of the CTA already fills shared memory
template <class T, class QT>
__device__ __forceinline__ void ltile_body(const InterpArgs<T, QT> &A);

// Levels with fewer tiles than SMs (the coarsest ones) are pure latency: one CTA of 1024 threads per tile halves it.
template <class T, class QT>
__global__ void __launch_bounds__(1024, 1) k_interp_ltile_wide(InterpArgs<T, QT> A) {
    ltile_body<T, QT>(A);
}

template <class T, class QT>
__global__ void __launch_bounds__(kTileThreads, sizeof(T) == 4 ? 2 : 1) k_interp_ltile(InterpArgs<T, QT> A) {
    ltile_body<T, QT>(A);
}

template <class T, class QT>
__device__ __forceinline__ void ltile_body(const InterpArgs<T, QT> &A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    unsigned *shist = reinterpret_cast<unsigned *>(smem_raw + sizeof(T) * kTileSmemElems);
    __shared__ LineTile lt;
    DevCtx2 ctx(shist, A.hist, A.qp.radius);
    for (int i = threadIdx.x; i < kHistWindow; i += blockDim.x) shist[i] = 0;
    LineGeom lg;
    const uint32_t tile = blockIdx.x + A.tile0;
    line_geom(A, tile, blockIdx.y, lg);
    if ((threadIdx.x & 31u) == 0 && (threadIdx.x >> 5) < 3)
        line_pass_setup(A, tile, blockIdx.y, static_cast<int>(threadIdx.x >> 5), blockDim.x, lt.ps[threadIdx.x >> 5]);
    line_fill(A, ctx, lg, sm);
    __syncthreads();
    line_tile_passes(A, ctx, sm, lg, lt);
    ctx.flush();
}

// ---------------------------------------------------------------------------------------------------------------------
// generic schedule, any N
// ---------------------------------------------------------------------------------------------------------------------
template <class T, class QT>
__global__ void __launch_bounds__(256) k_interp_pass(InterpArgs<T, QT> A, int p, uint64_t total) {
    __shared__ unsigned shist[kHistWindow];
    DevCtx ctx(shist, A.hist, A.qp.radius);
    ctx.clear();
    const uint64_t gid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    pass_point(A, ctx, p, gid, total, blockIdx.y);
    ctx.flush();
}

// row-mapped per-pass schedule (interp_lean.cuh), N >= 3; RECOVER = decompression side
template <class T, class QT, bool RECOVER>
__global__ void __launch_bounds__(kLeanThreads) k_interp_lean(LeanArgs<T, QT> P) {
    __shared__ LeanShared S;
    __shared__ unsigned shist[kHistWindow];
    DevCtx2 ctx(shist, P.A.hist, P.A.qp.radius);
    if (!RECOVER) {
        for (int i = threadIdx.x; i < kHistWindow; i += blockDim.x) shist[i] = 0;
    }
    lean_cta<T, QT, DevCtx2, RECOVER>(P, ctx, S, blockIdx.x, blockIdx.y, blockIdx.z);
    if (!RECOVER) {
        ctx.pass_end();
        ctx.flush();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------------
template <class T, class QT>
bool interp_launch_lean(const InterpArgs<T, QT> &A, int p, uint32_t nbatch, bool write_work, bool recover, const T *unpred_in,
                        cudaStream_t st) {
    const InterpShape &sh = A.sh;
    const int L = sh.N - 1;
    if (A.nb[L] > static_cast<uint32_t>(kLeanMaxBlocks)) return false;
    LeanArgs<T, QT> P;
    P.A = A;
    P.p = p;
    P.write_work = write_work ? 1u : 0u;
    P.unpred_in = unpred_in;
    P.lg_s = 0;
    while ((1u << P.lg_s) < A.s) P.lg_s++;
    const uint32_t row_len = lean_row_len(A, p);
    if (row_len == 0) return true;
    const uint32_t max_rows = lean_max_rows(A, p);
    P.chunks_per_brow = (max_rows + kLeanRows - 1) / kLeanRows;
    uint64_t nbrows = 1;
    for (int d = 0; d < L; d++) nbrows *= A.nb[d];
    const uint64_t gx = nbrows * P.chunks_per_brow;
    const uint32_t gy = (row_len + kLeanThreads - 1) / kLeanThreads;
    P.nchunks_L = gy;
    if (gx > 0x7fffffffull || gy > 65535u || nbatch > 65535u) return false;
    dim3 grid(static_cast<unsigned>(gx), gy, nbatch);
    if (recover)
        k_interp_lean<T, QT, true><<<grid, kLeanThreads, 0, st>>>(P);
    else
        k_interp_lean<T, QT, false><<<grid, kLeanThreads, 0, st>>>(P);
    return true;
}

template <class T, class QT>
void interp_launch_anchors(const InterpArgs<T, QT> &A, uint32_t anchor_stride, uint64_t n_anchor, uint32_t nbatch,
                           cudaStream_t st) {
    dim3 grid(static_cast<unsigned>((n_anchor + 255) / 256), nbatch);
    k_interp_anchor<T, QT><<<grid, 256, 0, st>>>(A, anchor_stride, n_anchor);
}

template <class T, class QT>
void interp_launch_ltiles(const InterpArgs<T, QT> &A, uint64_t ntiles, uint32_t nbatch, cudaStream_t st) {
    static std::atomic<unsigned long long> attr_set{0};
    const size_t smem = sizeof(T) * kTileSmemElems + sizeof(unsigned) * kHistWindow;
    once_per_device(attr_set, [&] {
        cudaFuncSetAttribute(k_interp_ltile<T, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if constexpr (sizeof(T) == 4)
            cudaFuncSetAttribute(k_interp_ltile_wide<T, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    });
    dim3 grid(static_cast<unsigned>(ntiles), nbatch);
    // (the 1024-thread variant for launches that leave SMs idle exists for 4-byte elements only: at 64 registers the
    //  double-precision body would spill)
    if constexpr (sizeof(T) == 4) {
        if (ntiles * nbatch <= 148) {
            k_interp_ltile_wide<T, QT><<<grid, 1024, smem, st>>>(A);
            return;
        }
    }
    k_interp_ltile<T, QT><<<grid, kTileThreads, smem, st>>>(A);
}

template <class T, class QT>
void interp_launch_pass(const InterpArgs<T, QT> &A, int p, uint32_t nbatch, cudaStream_t st) {
    uint64_t total = pass_points(A, p);
    if (total == 0) return;
    dim3 grid(static_cast<unsigned>((total + 255) / 256), nbatch);
    k_interp_pass<T, QT><<<grid, 256, 0, st>>>(A, p, total);
}

#define SZ3B_INST(T, QT)                                                                                         \
    template void interp_launch_anchors<T, QT>(const InterpArgs<T, QT> &, uint32_t, uint64_t, uint32_t,        \
                                               cudaStream_t);                                                   \
    template void interp_launch_ltiles<T, QT>(const InterpArgs<T, QT> &, uint64_t, uint32_t, cudaStream_t);     \
    template void interp_launch_pass<T, QT>(const InterpArgs<T, QT> &, int, uint32_t, cudaStream_t);          \
    template bool interp_launch_lean<T, QT>(const InterpArgs<T, QT> &, int, uint32_t, bool, bool, const T *,  \
                                            cudaStream_t);
SZ3B_INST(float, uint16_t)
SZ3B_INST(float, uint32_t)
SZ3B_INST(double, uint16_t)
SZ3B_INST(double, uint32_t)
// integer element types (tools/sz3/sz3.cpp:458-461): the per-pass kernels only
#define SZ3B_INST_GEN(T, QT)                                                                                     \
    template void interp_launch_anchors<T, QT>(const InterpArgs<T, QT> &, uint32_t, uint64_t, uint32_t,        \
                                               cudaStream_t);                                                   \
    template void interp_launch_pass<T, QT>(const InterpArgs<T, QT> &, int, uint32_t, cudaStream_t);           \
    template bool interp_launch_lean<T, QT>(const InterpArgs<T, QT> &, int, uint32_t, bool, bool, const T *,  \
                                            cudaStream_t);
SZ3B_INST_GEN(int32_t, uint16_t)
SZ3B_INST_GEN(int32_t, uint32_t)
SZ3B_INST_GEN(int64_t, uint16_t)
SZ3B_INST_GEN(int64_t, uint32_t)

}  // namespace sz3b
