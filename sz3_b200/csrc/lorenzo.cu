// sz3_b200/csrc/lorenzo.cu -- kernels of the BlockwiseDecomposition path with a Lorenzo predictor in the stack
// (single Lorenzo, Lorenzo 1+2, Lorenzo + regression; reference include/SZ3/api/impl/SZAlgoLorenzoReg.hpp:22-64).
//
//   k_bw_pad         the reference's zero-padded working copy (BlockwiseIterator.hpp:205-222,246-281)
//   k_bw_front       one front of the block wavefront: one warp per block, body in lorenzo.cuh
//   k_bw_spec_coef   lattice guess of the reconstructed regression coefficients (selection guess pass)
//   k_bw_rank        dense rank of the regression-selected blocks (single-CTA scan) + their count
//   k_bw_gather_fit  fitted coefficients of the selected blocks, dense, in row-major block order (chain input)
//
// Compiled with -fmad=false like every kernel that reproduces reference arithmetic.
#include <cuda_runtime.h>

#include "launch.hpp"
#include "lorenzo.cuh"

namespace sz3b {

template <class T>
__global__ void __launch_bounds__(256) k_bw_pad(const T *__restrict__ data, BlockShape bs, uint64_t ps0, uint64_t ps1,
                                                uint64_t ps2, uint64_t ps3, T *__restrict__ W, uint64_t b_lo, uint64_t b_hi, uint64_t e_lo,
                                                uint64_t e_hi, uint64_t w_bstride) {
    data += static_cast<uint64_t>(blockIdx.y) * bs.num;   // batch member (the tuner's sampled blocks)
    W += static_cast<uint64_t>(blockIdx.y) * w_bstride;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const uint64_t ps[kMaxDim] = {ps0, ps1, ps2, ps3};
    for (uint64_t i = e_lo + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < e_hi; i += stride) {
        uint64_t r = i, w = 0, b = 0, bmul = 1;
        for (int d = bs.N - 1; d >= 0; d--) {
            const uint64_t x = r % bs.dims[d];
            r /= bs.dims[d];
            w += (x + kBwPad) * ps[d];
            b += (x / bs.B) * bmul;
            bmul *= bs.nb[d];
        }
        if (b >= b_lo && b < b_hi) W[w] = data[i];   // a sub-range restores the originals of blocks that are not final
    }
}

// Row-major walk of the blocks [b_lo, nblocks) by one warp, the coefficient chain inline: the exact sequential
// semantics of the reference, used to finish when the selection iteration keeps being invalidated (lorenzo.cuh).
template <class T, class QT>
__global__ void __launch_bounds__(32) k_bw_serial(BwArgs<T, QT> A, uint64_t b_lo, uint64_t b_hi, const T *chain_init,
                                                  unsigned long long nsel0, unsigned long long *__restrict__ nsel_out,
                                                  uint32_t tile_cap) {
    extern __shared__ __align__(16) unsigned char bw_smem[];
    __shared__ BwSerial<T> st;
    T *tile = reinterpret_cast<T *>(bw_smem);
    T *est = tile + tile_cap;
    const int N = A.bs.N;
    if (threadIdx.x == 0) {
        for (int d = 0; d <= N; d++) st.prev[d] = chain_init ? chain_init[d] : static_cast<T>(0);
        st.nsel = nsel0;
    }
    __syncwarp();
    for (uint64_t b = b_lo; b < b_hi; b++) {
        uint32_t bi[kMaxDim] = {0, 0, 0, 0};
        uint64_t r = b;
        for (int d = N - 1; d >= 0; d--) {
            bi[d] = static_cast<uint32_t>(r % A.bs.nb[d]);
            r /= A.bs.nb[d];
        }
        bw_process_block<T, QT>(A, bi, tile, est, threadIdx.x, blockDim.x, &st);
        __threadfence_block();
        __syncwarp();
    }
    if (threadIdx.x == 0) *nsel_out = st.nsel;
}

// One CTA (one warp) per tuple of leading block coordinates (lead_lo + blockIdx.x); the last block coordinate follows
// from the front.  Only blocks whose row-major index lies in [A.b_lo, A.b_hi) are processed (window of the selection
// iteration).  N == 1 has a single CTA that walks all fronts (the 1-D Lorenzo recurrence is serial).
// BATCH: the tuner's sampled blocks (blockIdx.y = batch member).  Kept out of the plain instantiation because offsetting
// the pointers needs a writable copy of the argument block, which then lives in local memory instead of the constant bank.
template <class T, class QT, bool BATCH>
__global__ void __launch_bounds__(32) k_bw_front(const BwArgs<T, QT> A0, uint32_t f0, uint32_t f1, uint32_t lead_lo,
                                                 uint32_t tile_cap) {
    extern __shared__ __align__(16) unsigned char bw_smem[];
    T *tile = reinterpret_cast<T *>(bw_smem);
    T *est = tile + tile_cap;
    BwArgs<T, QT> Ab;
    if (BATCH) {   // same shape, own working array / index range / selection
        Ab = A0;
        Ab.W += static_cast<uint64_t>(blockIdx.y) * Ab.w_bstride;
        Ab.q += static_cast<uint64_t>(blockIdx.y) * Ab.q_bstride;
        Ab.unpred_tmp += static_cast<uint64_t>(blockIdx.y) * Ab.q_bstride;
        if (Ab.sel_out) Ab.sel_out += static_cast<uint64_t>(blockIdx.y) * Ab.sel_bstride;
        if (Ab.sel_in) Ab.sel_in += static_cast<uint64_t>(blockIdx.y) * Ab.sel_bstride;
    }
    const BwArgs<T, QT> &A = BATCH ? Ab : A0;
    const int N = A.bs.N;
    uint32_t bi[kMaxDim] = {0, 0, 0, 0};
    const uint32_t lead = blockIdx.x + lead_lo;
    uint32_t r = lead, s = 0;
    for (int d = N - 2; d >= 0; d--) {
        bi[d] = r % A.bs.nb[d];
        r /= A.bs.nb[d];
        s += bi[d];
    }
    if (f1 == f0 + 1) {   // one front per launch: most CTAs of a launch have no block on it
        const uint64_t b = static_cast<uint64_t>(lead) * A.bs.nb[N - 1] + (f0 - s);
        if (f0 < s || f0 - s >= A.bs.nb[N - 1] || b < A.b_lo || b >= A.b_hi) return;
    }
    const uint64_t row_base = static_cast<uint64_t>(lead) * A.bs.nb[N - 1];
    for (uint32_t f = f0; f < f1; f++) {
        if (f < s) continue;
        const uint32_t last = f - s;
        if (last >= A.bs.nb[N - 1]) break;
        const uint64_t b = row_base + last;
        if (b < A.b_lo || b >= A.b_hi) continue;
        bi[N - 1] = last;
        bw_process_block<T, QT>(A, bi, tile, est, threadIdx.x, blockDim.x);
        __threadfence_block();
        __syncwarp();
    }
}

template <class T>
__global__ void __launch_bounds__(256) k_bw_spec_coef(const T *__restrict__ c_fit, const uint8_t *__restrict__ valid,
                                                      uint64_t nblocks, int N, QuantParams q_liner, QuantParams q_indep,
                                                      T *__restrict__ c_spec) {
    const int nc = N + 1;
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nblocks * nc) return;
    const uint64_t b = i / nc;
    const int d = static_cast<int>(i - b * nc);
    c_spec[i] = valid[b] ? coef_lattice_guess<T>(c_fit[i], d < N ? q_liner : q_indep) : static_cast<T>(0);
}

__global__ void __launch_bounds__(1024) k_bw_rank(const uint8_t *__restrict__ sel, uint64_t b_lo, uint64_t b_hi, int reg_sid,
                                                  uint32_t base, uint32_t *__restrict__ rank,
                                                  unsigned long long *__restrict__ count) {
    __shared__ unsigned warp_sum[32];
    __shared__ unsigned carry_s;
    if (threadIdx.x == 0) carry_s = base;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint64_t start = b_lo; start < b_hi; start += 1024) {
        const uint64_t b = start + threadIdx.x;
        const unsigned v = b < b_hi && sel[b] == reg_sid ? 1u : 0u;
        unsigned x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            unsigned w = warp_sum[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sum[lane] = w;   // inclusive
        }
        __syncthreads();
        const unsigned carry = carry_s;
        const unsigned before = carry + (wid ? warp_sum[wid - 1] : 0u) + x - v;
        if (b < b_hi) rank[b] = before;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry_s - base;   // regression-selected blocks inside the window
}

template <class T>
__global__ void __launch_bounds__(256) k_bw_gather_fit(const T *__restrict__ c_fit, const uint8_t *__restrict__ sel,
                                                       int reg_sid, const uint32_t *__restrict__ rank, uint64_t b_lo,
                                                       uint64_t b_hi, int nc, T *__restrict__ c_dense) {
    const uint64_t i = b_lo * nc + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= b_hi * nc) return;
    const uint64_t b = i / nc;
    if (sel[b] != reg_sid) return;
    c_dense[static_cast<uint64_t>(rank[b]) * nc + (i - b * nc)] = c_fit[i];
}

__global__ void __launch_bounds__(256) k_widen_u8(const uint8_t *__restrict__ in, uint64_t n, int32_t *__restrict__ out) {
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// ---------------------------------------------------------------------------------------------------------------------
template <class T>
void launch_bw_pad(const T *data, const BlockShape &bs, const uint64_t *pstride, T *W, uint64_t b_lo, uint64_t b_hi,
                   cudaStream_t st, uint32_t nbatch, uint64_t w_bstride) {
    if (b_lo >= b_hi) return;
    // elements of the outermost-dimension slabs the block range touches
    const uint64_t per_plane = bs.nblocks / bs.nb[0];
    const uint64_t i_lo = b_lo / per_plane, i_hi = (b_hi - 1) / per_plane;
    const uint64_t x_hi = (i_hi + 1) * bs.B < bs.dims[0] ? (i_hi + 1) * bs.B : bs.dims[0];
    const uint64_t e_lo = i_lo * bs.B * bs.stride[0], e_hi = x_hi * bs.stride[0];
    uint64_t blocks = (e_hi - e_lo + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    k_bw_pad<T><<<dim3(static_cast<unsigned>(blocks), nbatch), 256, 0, st>>>(data, bs, pstride[0], pstride[1], pstride[2],
                                                                             pstride[3], W, b_lo, b_hi, e_lo, e_hi, w_bstride);
}

static size_t bw_tile_cap(const BlockShape &bs) {
    size_t tile_cap = 1;
    for (int d = 0; d < bs.N; d++) tile_cap *= (bs.dims[d] < bs.B ? bs.dims[d] : bs.B) + kBwPad;
    return tile_cap;
}

template <class T, class QT>
const char *launch_bw_serial(const BwArgs<T, QT> &A, uint64_t b_lo, uint64_t b_hi, const T *chain_init,
                             unsigned long long nsel0, unsigned long long *nsel_out, cudaStream_t st) {
    const size_t smem = bw_scratch_elems(A.bs, A.nk) * sizeof(T);
    if (smem > 200 * 1024) return "blockSize too large for the shared-memory tile of the Lorenzo kernel";
    static thread_local size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        if (cudaFuncSetAttribute(k_bw_serial<T, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return "cannot raise the dynamic shared memory limit";
        attr_set = 200 * 1024;
    }
    k_bw_serial<T, QT><<<1, A.bs.N == 1 ? 1 : 32, smem, st>>>(A, b_lo, b_hi, chain_init, nsel0, nsel_out,
                                                                 static_cast<uint32_t>(bw_tile_cap(A.bs)));
    return nullptr;
}

template <class T, class QT>
const char *launch_bw_fronts(const BwArgs<T, QT> &A, cudaStream_t st, int *launches) {
    const BlockShape &bs = A.bs;
    const int N = bs.N;
    uint64_t nlead = 1;
    for (int d = 0; d < N - 1; d++) nlead *= bs.nb[d];
    if (nlead > 0x7fffffffull) return "block grid exceeds the launch grid of the Lorenzo kernel";
    const size_t tile_cap = bw_tile_cap(bs);
    // (the per-diagonal tables of BwArgs stay in global memory: a few hundred bytes shared by every block, L1 hits;
    //  staging them per CTA costs more than it saves, one block per CTA)
    const size_t smem = bw_scratch_elems(bs, A.nk) * sizeof(T);
    if (smem > 200 * 1024) return "blockSize too large for the shared-memory tile of the Lorenzo kernel";
    static thread_local size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        if (cudaFuncSetAttribute(k_bw_front<T, QT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(k_bw_front<T, QT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return "cannot raise the dynamic shared memory limit";
        attr_set = 200 * 1024;
    }
    const uint64_t b_lo = A.b_lo, b_hi = A.b_hi < bs.nblocks ? A.b_hi : bs.nblocks;
    if (b_lo >= b_hi) return nullptr;
    const uint32_t nbl = bs.nb[N - 1];
    const uint64_t lead_lo = b_lo / nbl, lead_hi = (b_hi - 1) / nbl;
    // fronts touched by the window: per row of blocks (one lead tuple) the coordinate sum of the leads + the k range
    uint32_t f_min = ~0u, f_max = 0;
    if (b_lo == 0 && b_hi == bs.nblocks) {
        f_min = 0;
        f_max = bw_num_fronts(bs) - 1;
    } else {
        for (uint64_t lead = lead_lo; lead <= lead_hi; lead++) {
            uint64_t r = lead;
            uint32_t s = 0;
            for (int d = N - 2; d >= 0; d--) {
                s += static_cast<uint32_t>(r % bs.nb[d]);
                r /= bs.nb[d];
            }
            const uint64_t k0 = lead == lead_lo ? b_lo - lead * nbl : 0;
            const uint64_t k1 = lead == lead_hi ? (b_hi - 1) - lead * nbl : nbl - 1;
            f_min = s + k0 < f_min ? static_cast<uint32_t>(s + k0) : f_min;
            f_max = s + k1 > f_max ? static_cast<uint32_t>(s + k1) : f_max;
        }
    }
    const unsigned grid = static_cast<unsigned>(lead_hi - lead_lo + 1);
    const unsigned nbatch = A.nbatch ? A.nbatch : 1u;
    const uint32_t tc = static_cast<uint32_t>(tile_cap), ll = static_cast<uint32_t>(lead_lo);
    if (N == 1) {
        // the 1-D recurrence is serial point by point: one thread walks the blocks (a warp would only add barriers)
        if (nbatch > 1)
            k_bw_front<T, QT, true><<<dim3(1, nbatch), 1, smem, st>>>(A, f_min, f_max + 1, 0, tc);
        else
            k_bw_front<T, QT, false><<<1, 1, smem, st>>>(A, f_min, f_max + 1, 0, tc);
        *launches += 1;
    } else {
        for (uint32_t f = f_min; f <= f_max; f++) {
            if (nbatch > 1)
                k_bw_front<T, QT, true><<<dim3(grid, nbatch), 32, smem, st>>>(A, f, f + 1, ll, tc);
            else
                k_bw_front<T, QT, false><<<grid, 32, smem, st>>>(A, f, f + 1, ll, tc);
        }
        *launches += static_cast<int>(f_max - f_min + 1);
    }
    return nullptr;
}

template <class T>
void launch_bw_spec_coef(const T *c_fit, const uint8_t *valid, uint64_t nblocks, int N, const QuantParams &q_liner,
                         const QuantParams &q_indep, T *c_spec, cudaStream_t st) {
    const uint64_t n = nblocks * (N + 1);
    k_bw_spec_coef<T><<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(c_fit, valid, nblocks, N, q_liner, q_indep, c_spec);
}

void launch_bw_rank(const uint8_t *sel, uint64_t b_lo, uint64_t b_hi, int reg_sid, uint32_t base, uint32_t *rank,
                    unsigned long long *count, cudaStream_t st) {
    k_bw_rank<<<1, 1024, 0, st>>>(sel, b_lo, b_hi, reg_sid, base, rank, count);
}

template <class T>
void launch_bw_gather_fit(const T *c_fit, const uint8_t *sel, int reg_sid, const uint32_t *rank, uint64_t b_lo, uint64_t b_hi,
                          int nc, T *c_dense, cudaStream_t st) {
    const uint64_t n = (b_hi - b_lo) * nc;
    if (n == 0) return;
    k_bw_gather_fit<T><<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(c_fit, sel, reg_sid, rank, b_lo, b_hi, nc, c_dense);
}

void launch_widen_u8(const uint8_t *in, uint64_t n, int32_t *out, cudaStream_t st) {
    if (n == 0) return;
    k_widen_u8<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(in, n, out);
}

#define SZ3B_INST_LZ(T)                                                                                              \
    template void launch_bw_pad<T>(const T *, const BlockShape &, const uint64_t *, T *, uint64_t, uint64_t,         \
                                   cudaStream_t, uint32_t, uint64_t);                                                \
    template const char *launch_bw_serial<T, uint16_t>(const BwArgs<T, uint16_t> &, uint64_t, uint64_t, const T *,   \
                                                       unsigned long long, unsigned long long *, cudaStream_t);      \
    template const char *launch_bw_serial<T, uint32_t>(const BwArgs<T, uint32_t> &, uint64_t, uint64_t, const T *,   \
                                                       unsigned long long, unsigned long long *, cudaStream_t);      \
    template const char *launch_bw_fronts<T, uint16_t>(const BwArgs<T, uint16_t> &, cudaStream_t, int *);            \
    template const char *launch_bw_fronts<T, uint32_t>(const BwArgs<T, uint32_t> &, cudaStream_t, int *);            \
    template void launch_bw_spec_coef<T>(const T *, const uint8_t *, uint64_t, int, const QuantParams &,             \
                                         const QuantParams &, T *, cudaStream_t);                                    \
    template void launch_bw_gather_fit<T>(const T *, const uint8_t *, int, const uint32_t *, uint64_t, uint64_t, int, \
                                          T *, cudaStream_t);
SZ3B_INST_LZ(float)
SZ3B_INST_LZ(double)
SZ3B_INST_LZ(int32_t)
SZ3B_INST_LZ(int64_t)

}  // namespace sz3b
