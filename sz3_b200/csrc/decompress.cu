// sz3_b200/csrc/decompress.cu -- kernels of the SZ_decompress half of the path: the `recover` side of
// InterpolationDecomposition::decompress (reference include/SZ3/decomposition/InterpolationDecomposition.hpp:26-76),
// BlockwiseDecomposition::decompress (decomposition/BlockwiseDecomposition.hpp:48-67) with
// RegressionPredictor::predecompress (predictor/RegressionPredictor.hpp:62-71,157-165), and
// LinearQuantizer::recover (quantizer/LinearQuantizer.hpp:74-86).
//
//   k_zero_count / k_zero_scatter   the quantizer's unpredictable values arrive as a dense list in traversal order;
//                                   rank every zero index (chunked counts + scan + in-chunk scan) and scatter the list
//                                   to a position-indexed array, so that recover kernels never need a running cursor
//   k_interp_recover_anchor/_pass   level -> pass -> all points (valid by the "global-pass schedule" equivalence,
//                                   SURVEY.md Appendix B): each thread predicts one point from already recovered
//                                   neighbours in the output array and applies its index
//   k_reg_chain_recover             coefficient recurrence (one lane per coefficient; a pure dependent-add chain)
//   k_reg_recover                   row-mapped predict + recover
//
// Same arithmetic rules as compression (-fmad=false, T arithmetic left to right).
#include <cuda_runtime.h>

#include "blockwise.cuh"
#include "interp_body.cuh"
#include "launch.hpp"

namespace sz3b {

constexpr int kZThreads = 256;
constexpr int kZPer = 16;
constexpr int kZChunk = kZThreads * kZPer;

template <class QT>
__global__ void __launch_bounds__(kZThreads) k_zero_count(const QT *__restrict__ q, uint64_t n, unsigned *__restrict__ chunk_zeros,
                                                         unsigned *__restrict__ chunk_bits) {
    __shared__ unsigned wz[kZThreads / 32];
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kZChunk + static_cast<uint64_t>(threadIdx.x) * kZPer;
    unsigned z = 0;
#pragma unroll
    for (int k = 0; k < kZPer; k++)
        if (base + k < n) z += q[base + k] == 0;
    for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    if ((threadIdx.x & 31) == 0) wz[threadIdx.x >> 5] = z;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < kZThreads / 32; w++) t += wz[w];
        chunk_zeros[blockIdx.x] = t;
        chunk_bits[blockIdx.x] = 0;
    }
}

template <class QT, class T>
__global__ void __launch_bounds__(kZThreads) k_zero_scatter(const QT *__restrict__ q, uint64_t n,
                                                           const unsigned long long *__restrict__ zero_off,
                                                           const T *__restrict__ unpred, uint64_t n_unpred,
                                                           T *__restrict__ unpred_tmp) {
    __shared__ unsigned ws[kZThreads / 32];
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kZChunk + static_cast<uint64_t>(threadIdx.x) * kZPer;
    unsigned z = 0;
#pragma unroll
    for (int k = 0; k < kZPer; k++)
        if (base + k < n) z += q[base + k] == 0;
    // exclusive scan of z over the CTA
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = z;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= static_cast<unsigned>(o)) inc += t;
    }
    if (lane == 31) ws[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned w = lane < kZThreads / 32 ? ws[lane] : 0, winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= static_cast<unsigned>(o)) winc += t;
        }
        if (lane < kZThreads / 32) ws[lane] = winc - w;
    }
    __syncthreads();
    unsigned long long rank = zero_off[blockIdx.x] + ws[wid] + inc - z;
    if (z) {
#pragma unroll
        for (int k = 0; k < kZPer; k++)
            if (base + k < n && q[base + k] == 0) {
                if (rank < n_unpred) unpred_tmp[base + k] = unpred[rank];
                rank++;
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// interpolation
// ---------------------------------------------------------------------------------------------------------------------
template <class T, class QT>
struct RecoverArgs {
    InterpShape sh;
    T *out;                      // array being reconstructed (also the source of already recovered neighbours)
    const QT *q;                 // indices, traversal order
    const T *unpred_tmp;         // value at pos where q[pos] == 0
    QuantParams qp;
    uint32_t s;
    uint32_t nb[kMaxDim];
    const uint64_t *block_base;
};

template <class T, class QT>
__global__ void __launch_bounds__(256) k_interp_recover_anchor(RecoverArgs<T, QT> A, uint32_t anchor_stride, uint64_t n_anchor) {
    const uint64_t gid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gid >= n_anchor) return;
    const InterpShape &sh = A.sh;
    if (anchor_stride == 0) {   // first element, recovered against a zero prediction (:33-34)
        const int qv = static_cast<int>(A.q[0]);
        A.out[0] = qv ? recover_pred<T>(static_cast<T>(0), qv, A.qp) : A.unpred_tmp[0];
        return;
    }
    uint64_t r = gid, off = 0;
    for (int d = sh.N - 1; d >= 0; d--) {
        const uint32_t ext = (sh.dims[d] - 1) / anchor_stride + 1;
        off += static_cast<uint64_t>(static_cast<uint32_t>(r % ext) * anchor_stride) * sh.stride[d];
        r /= ext;
    }
    const int qv = static_cast<int>(A.q[gid]);   // anchors are written as unpredictable (index 0)
    A.out[off] = qv ? recover_pred<T>(static_cast<T>(0), qv, A.qp) : A.unpred_tmp[gid];
}

// one predicted point of pass p at level stride A.s (mirror of pass_point in interp_body.cuh)
template <class T, class QT>
__global__ void __launch_bounds__(256) k_interp_recover_pass(RecoverArgs<T, QT> A, int p, uint64_t total) {
    const uint64_t gid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const InterpShape &sh = A.sh;
    const uint32_t s = A.s;
    const int D = sh.perm[p];
    uint32_t step[kMaxDim], ext[kMaxDim], x[kMaxDim], bidx[kMaxDim];
    for (int qd = 0; qd < sh.N; qd++) {
        const int d = sh.perm[qd];
        step[d] = qd < p ? s : 2 * s;
        ext[d] = qd == p ? ((sh.dims[d] - 1) / s + 1) / 2 : (sh.dims[d] - 1) / step[d] + 1;
    }
    uint64_t off = 0, blin = 0;
    const uint32_t B = kInterpBlock * s;
    if (total <= 0xffffffffull) {   // 32-bit index arithmetic
        uint32_t r = static_cast<uint32_t>(gid);
        for (int d = sh.N - 1; d >= 0; d--) {
            const uint32_t qd = r / ext[d];
            const uint32_t idx = r - qd * ext[d];
            r = qd;
            x[d] = d == D ? (2 * idx + 1) * s : idx * step[d];
            bidx[d] = d == D ? x[d] / B : (x[d] ? (x[d] - 1) / B : 0);
            off += x[d] * sh.stride[d];
        }
    } else {
        uint64_t r = gid;
        for (int d = sh.N - 1; d >= 0; d--) {
            const uint32_t idx = static_cast<uint32_t>(r % ext[d]);
            r /= ext[d];
            x[d] = d == D ? (2 * idx + 1) * s : idx * step[d];
            bidx[d] = d == D ? x[d] / B : (x[d] ? (x[d] - 1) / B : 0);
            off += x[d] * sh.stride[d];
        }
    }
    for (int d = 0; d < sh.N; d++) blin = blin * A.nb[d] + bidx[d];
    BlockGeom g;
    block_geom(sh, s, bidx, g);
    uint64_t base = A.block_base[blin];
    PassGeom pg;
    for (int pp = 0; pp < p; pp++) {
        pass_geom(sh, s, g, pp, pg);
        base += pg.size;
    }
    pass_geom(sh, s, g, p, pg);
    const uint32_t n = pg.n;
    const uint32_t i = (x[D] - g.begin[D]) / s;
    const int64_t sd = static_cast<int64_t>(s) * static_cast<int64_t>(sh.stride[D]);
    auto v = [&](uint32_t k) -> T {
        return A.out[static_cast<int64_t>(off) + (static_cast<int64_t>(k) - static_cast<int64_t>(i)) * sd];
    };
    T pred;
    uint64_t in_pass;
    if (sh.old_api) {
        pred = predict_line_old<T>(sh.cubic, i, n, v);
        uint64_t line = 0;
        for (int d = 0; d < sh.N; d++)
            if (d != D) line = line * pg.cnt[d] + (x[d] - pg.lo[d]) / pg.step[d];
        in_pass = line * (n / 2) + line_offset_old(sh.cubic, i, n);
    } else {
        T r2 = 0;
        if (!sh.cubic && i + 1 == n && n >= 4) {
            // linear tail: needs the reconstruction of i-2, recovered in this same pass by another thread; rebuild it
            // from its own index instead of racing on the output array
            uint32_t x2[kMaxDim] = {x[0], x[1], x[2], x[3]};
            x2[D] -= 2 * s;
            const uint64_t pos2 = base + pass_offset(sh, pg, x2, i - 2);
            const int q2 = static_cast<int>(A.q[pos2]);
            const T p2 = interp_linear<T>(v(i - 3), v(i - 1));
            r2 = q2 ? recover_pred<T>(p2, q2, A.qp) : A.unpred_tmp[pos2];
        }
        pred = predict_line<T>(sh.cubic, i, n, v, r2);
        in_pass = pass_offset(sh, pg, x, i);
    }
    const uint64_t pos = base + in_pass;
    const int qv = static_cast<int>(A.q[pos]);
    A.out[off] = qv ? recover_pred<T>(pred, qv, A.qp) : A.unpred_tmp[pos];
}

// ---------------------------------------------------------------------------------------------------------------------
// regression
// ---------------------------------------------------------------------------------------------------------------------
// pred_and_recover_coefficients (RegressionPredictor.hpp:157-165): current = recover(current, index); the running
// value is the previous block's coefficient.  coef_unp[pos] holds the stored exact value where coef_q[pos] == 0.
// The recurrence is serial in the blocks (every step rounds), but everything except the dependent add can be taken
// off it: a CTA stages chunks of blocks in shared memory -- the step 2 (q - radius) eb of every coefficient, computed
// by all threads from coalesced loads -- then one lane per coefficient walks the chunk (an add and a select per
// block, operands prefetched from shared memory), and the recovered coefficients leave coalesced.  (One lane reading
// its index from global memory block by block paid a memory latency per block: 37 ms for the 262 144 blocks of C3.)
constexpr int kRcChunk = 256;
constexpr int kRcThreads = 128;
template <class T>
__global__ void __launch_bounds__(kRcThreads) k_reg_chain_recover(const int32_t *__restrict__ coef_q, const T *__restrict__ coef_unp,
                                                                  uint64_t nblocks, int N, QuantParams q_liner, QuantParams q_indep,
                                                                  T *__restrict__ c_rec) {
    __shared__ double sd[kRcChunk * (kMaxDim + 1)];
    __shared__ T sv[kRcChunk * (kMaxDim + 1)];         // the stored exact value where q == 0
    __shared__ T so[kRcChunk * (kMaxDim + 1)];         // the recovered coefficients (an array of their own: the walk's
                                                       // loads must not wait for its stores)
    __shared__ uint8_t sz[kRcChunk * (kMaxDim + 1)];   // q == 0
    const int tid = threadIdx.x, nc = N + 1;
    T cur = 0;
    for (uint64_t base = 0; base < nblocks; base += kRcChunk) {
        const uint32_t cnt = static_cast<uint32_t>(nblocks - base < kRcChunk ? nblocks - base : kRcChunk);
        const uint32_t m = cnt * nc;
        for (uint32_t i = tid; i < m; i += kRcThreads) {
            const int qv = coef_q[base * nc + i];
            const QuantParams &qp = static_cast<int>(i % nc) < N ? q_liner : q_indep;
            sz[i] = qv == 0;
            sd[i] = static_cast<double>(2 * (qv - qp.radius)) * qp.eb;
            if (qv == 0) sv[i] = coef_unp[base * nc + i];
        }
        __syncthreads();
        if (tid < nc) {
            // eight blocks at a time: operands into registers first, then the eight dependent steps
            // (recover_pred, core.cuh: pred + 2 (q - radius) eb in double, stored as T), then the results
            for (uint32_t i0 = 0; i0 < cnt; i0 += 8) {
                double d[8];
                T u[8], r[8];
                bool z[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {   // (reads past `cnt` stay inside the arrays; their results are dropped)
                    const uint32_t k = (i0 + j) * nc + tid;
                    d[j] = sd[k];
                    u[j] = sv[k];
                    z[j] = sz[k] != 0;
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const T step = from_double<T>(static_cast<double>(cur) + d[j]);
                    cur = z[j] ? u[j] : step;
                    r[j] = cur;
                }
                T last = cur;
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (i0 + j < cnt) {
                        so[(i0 + j) * nc + tid] = r[j];
                        last = r[j];
                    }
                cur = last;   // (differs from r[7] only in a ragged last group)
            }
        }
        __syncthreads();
        for (uint32_t i = tid; i < m; i += kRcThreads) c_rec[base * nc + i] = so[i];
        __syncthreads();
    }
}

constexpr int kRecThreads = 128;
constexpr int kRecChunk = 512;

template <class T, class QT>
__global__ void __launch_bounds__(kRecThreads) k_reg_recover(T *__restrict__ out, BlockShape bs, uint32_t nchunks, uint32_t mgB,
                                                            const T *__restrict__ c_rec, QuantParams qp,
                                                            const QT *__restrict__ q, const T *__restrict__ unpred_tmp) {
    const int N = bs.N;
    uint32_t xr[kMaxDim] = {0, 0, 0, 0};
    uint32_t chunk = blockIdx.x;
    if (N >= 2) {
        xr[N - 2] = blockIdx.x / nchunks;
        chunk = blockIdx.x - xr[N - 2] * nchunks;
    }
    if (N >= 3) xr[N - 3] = blockIdx.y;
    if (N >= 4) xr[N - 4] = blockIdx.z;
    RegRow rr;
    reg_row_setup(bs, xr, rr);
    uint64_t row_off = 0;
    for (int d = 0; d < N - 1; d++) row_off += xr[d] * bs.stride[d];
    const uint32_t len = bs.dims[N - 1];
    const int nc = N + 1;
#pragma unroll
    for (int k = 0; k < kRecChunk / kRecThreads; k++) {
        const uint32_t x = chunk * kRecChunk + k * kRecThreads + threadIdx.x;
        if (x < len) {
            uint64_t blin, pos;
            uint32_t li[kMaxDim] = {rr.li[0], rr.li[1], rr.li[2], rr.li[3]};
            reg_row_locate(bs, rr, x, mgB, &blin, &li[N - 1], &pos);
            const T pred = reg_predict<T>(N, c_rec + blin * nc, li);
            const int qv = static_cast<int>(q[pos]);
            out[row_off + x] = qv ? recover_pred<T>(pred, qv, qp) : unpred_tmp[pos];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------------
uint64_t zero_num_chunks(uint64_t n) { return (n + kZChunk - 1) / kZChunk; }

template <class QT>
void launch_zero_count(const QT *q, uint64_t n, unsigned *chunk_zeros, unsigned *chunk_bits, cudaStream_t st) {
    const uint64_t nch = zero_num_chunks(n);
    if (nch) k_zero_count<QT><<<static_cast<unsigned>(nch), kZThreads, 0, st>>>(q, n, chunk_zeros, chunk_bits);
}
template <class QT, class T>
void launch_zero_scatter(const QT *q, uint64_t n, const unsigned long long *zero_off, const T *unpred, uint64_t n_unpred,
                         T *unpred_tmp, cudaStream_t st) {
    const uint64_t nch = zero_num_chunks(n);
    if (nch) k_zero_scatter<QT, T><<<static_cast<unsigned>(nch), kZThreads, 0, st>>>(q, n, zero_off, unpred, n_unpred, unpred_tmp);
}

template <class T, class QT>
void launch_interp_recover(const InterpShape &sh, T *out, const QT *q, const T *unpred_tmp, const QuantParams &qp, uint32_t s,
                           const uint32_t nb[kMaxDim], const uint64_t *block_base, int pass, uint32_t anchor_stride,
                           uint64_t n_anchor, cudaStream_t st) {
    RecoverArgs<T, QT> A;
    A.sh = sh;
    A.out = out;
    A.q = q;
    A.unpred_tmp = unpred_tmp;
    A.qp = qp;
    A.s = s;
    for (int d = 0; d < kMaxDim; d++) A.nb[d] = nb ? nb[d] : 1;
    A.block_base = block_base;
    if (pass < 0) {
        k_interp_recover_anchor<T, QT><<<static_cast<unsigned>((n_anchor + 255) / 256), 256, 0, st>>>(A, anchor_stride, n_anchor);
        return;
    }
    // points of this pass (same count as pass_points in interp_body.cuh)
    uint64_t total = 1;
    for (int qd = 0; qd < sh.N; qd++) {
        const int d = sh.perm[qd];
        total *= qd == pass ? ((sh.dims[d] - 1) / s + 1) / 2 : (sh.dims[d] - 1) / (qd < pass ? s : 2 * s) + 1;
    }
    if (total == 0) return;
    k_interp_recover_pass<T, QT><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(A, pass, total);
}

template <class T>
void launch_reg_chain_recover(const int32_t *coef_q, const T *coef_unp, uint64_t nblocks, int N, const QuantParams &q_liner,
                              const QuantParams &q_indep, T *c_rec, cudaStream_t st) {
    k_reg_chain_recover<T><<<1, kRcThreads, 0, st>>>(coef_q, coef_unp, nblocks, N, q_liner, q_indep, c_rec);
}

template <class T, class QT>
const char *launch_reg_recover(T *out, const BlockShape &bs, const T *c_rec, const QuantParams &qp, const QT *q,
                               const T *unpred_tmp, cudaStream_t st) {
    const int N = bs.N;
    const uint32_t len = bs.dims[N - 1];
    const uint32_t nchunks = (len + kRecChunk - 1) / kRecChunk;
    const uint64_t gx = static_cast<uint64_t>(nchunks) * (N >= 2 ? bs.dims[N - 2] : 1);
    const uint32_t gy = N >= 3 ? bs.dims[N - 3] : 1, gz = N >= 4 ? bs.dims[N - 4] : 1;
    if (gx > 0x7fffffffull || gy > 65535u || gz > 65535u) return "array shape exceeds the launch grid of the regression kernel";
    const uint32_t mgB = (bs.B > 1 && static_cast<uint64_t>(len) * bs.B < (1ull << 32)) ? 0xffffffffu / bs.B + 1u : 0u;
    dim3 grid(static_cast<unsigned>(gx), gy, gz);
    k_reg_recover<T, QT><<<grid, kRecThreads, 0, st>>>(out, bs, nchunks, mgB, c_rec, qp, q, unpred_tmp);
    return nullptr;
}

#define SZ3B_INST_DEC(T, QT)                                                                                             \
    template void launch_zero_scatter<QT, T>(const QT *, uint64_t, const unsigned long long *, const T *, uint64_t, T *, \
                                             cudaStream_t);                                                              \
    template void launch_interp_recover<T, QT>(const InterpShape &, T *, const QT *, const T *, const QuantParams &,     \
                                               uint32_t, const uint32_t *, const uint64_t *, int, uint32_t, uint64_t,    \
                                               cudaStream_t);                                                            \
    template const char *launch_reg_recover<T, QT>(T *, const BlockShape &, const T *, const QuantParams &, const QT *,  \
                                                   const T *, cudaStream_t);
SZ3B_INST_DEC(float, uint16_t)
SZ3B_INST_DEC(float, uint32_t)
SZ3B_INST_DEC(double, uint16_t)
SZ3B_INST_DEC(double, uint32_t)
SZ3B_INST_DEC(int32_t, uint16_t)
SZ3B_INST_DEC(int32_t, uint32_t)
SZ3B_INST_DEC(int64_t, uint16_t)
SZ3B_INST_DEC(int64_t, uint32_t)
template void launch_zero_count<uint16_t>(const uint16_t *, uint64_t, unsigned *, unsigned *, cudaStream_t);
template void launch_zero_count<uint32_t>(const uint32_t *, uint64_t, unsigned *, unsigned *, cudaStream_t);
template void launch_reg_chain_recover<float>(const int32_t *, const float *, uint64_t, int, const QuantParams &,
                                              const QuantParams &, float *, cudaStream_t);
template void launch_reg_chain_recover<double>(const int32_t *, const double *, uint64_t, int, const QuantParams &,
                                               const QuantParams &, double *, cudaStream_t);
template void launch_reg_chain_recover<int32_t>(const int32_t *, const int32_t *, uint64_t, int, const QuantParams &,
                                                const QuantParams &, int32_t *, cudaStream_t);
template void launch_reg_chain_recover<int64_t>(const int32_t *, const int64_t *, uint64_t, int, const QuantParams &,
                                                const QuantParams &, int64_t *, cudaStream_t);

}  // namespace sz3b
