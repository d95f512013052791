// sz3_b200/csrc/blockwise.cu -- kernels of the BlockwiseDecomposition path (reference
// include/SZ3/decomposition/BlockwiseDecomposition.hpp:28-46) for the linear-regression predictor
// (include/SZ3/predictor/RegressionPredictor.hpp:28-60,77-91,148-155):
//
//   k_reg_fit      one thread per block: the N+1 sums in the reference's row-major order (double accumulators,
//                  products in T), then the coefficient formulas in double with the reference's store-to-T points.
//                  Sequential per block on purpose: for T = double the accumulation order decides the coefficient bits.
//   k_reg_chain    the coefficient delta-quantization chain (prev_coeffs = previous block's RECONSTRUCTED coefficients):
//                  a true serial recurrence over blocks, N+1 independent scalar chains -> N+1 lanes of one warp;
//                  the other lanes stage the fitted coefficients through shared memory.
//   k_reg_predict  fused predict + LinearQuantizer, one thread per element in memory order (coalesced reads), index
//                  written at its block-major traversal position, histogram fused (device_ctx.cuh).
//
// Compiled with -fmad=false (no FMA contraction anywhere in the reference arithmetic).
#include <cuda_runtime.h>

#include "blockwise.cuh"
#include "device_ctx.cuh"

namespace sz3b {

// ---------------------------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(128) k_reg_fit(const T *__restrict__ data, BlockShape bs, T *__restrict__ c_fit,
                                                 uint8_t *__restrict__ valid) {
    const uint64_t b = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (b >= bs.nblocks) return;
    T coef[kMaxDim + 1];
    const bool ok = reg_fit_block<T>(data, bs, b, coef);
    valid[b] = ok ? 1 : 0;
    if (!ok) return;
    for (int d = 0; d <= bs.N; d++) c_fit[b * (bs.N + 1) + d] = coef[d];
}

// ---------------------------------------------------------------------------------------------------------------------
// Serial chain.  One warp; lane d < N+1 owns coefficient d.  Blocks are staged through shared memory in chunks.
constexpr int kChainChunk = 256;

template <class T>
__global__ void __launch_bounds__(32) k_reg_chain(const T *__restrict__ c_fit, const uint8_t *__restrict__ sel,
                                                  uint64_t nblocks, int N, QuantParams q_liner, QuantParams q_indep,
                                                  int32_t *__restrict__ coef_q, T *__restrict__ c_rec,
                                                  unsigned long long *__restrict__ n_sel_out,
                                                  unsigned long long *__restrict__ n_unpred,
                                                  unsigned long long *__restrict__ unpred_pos, T *__restrict__ unpred_val,
                                                  const T *__restrict__ init, unsigned long long pos_base) {
    __shared__ T s_fit[kChainChunk * (kMaxDim + 1)];
    __shared__ T s_rec[kChainChunk * (kMaxDim + 1)];
    __shared__ int32_t s_q[kChainChunk * (kMaxDim + 1)];
    __shared__ uint8_t s_sel[kChainChunk];
    const int lane = threadIdx.x;
    const int nc = N + 1;
    const QuantParams qp = lane < N ? q_liner : q_indep;
    T prev = init && lane < nc ? init[lane] : static_cast<T>(0);   // chain continued from an earlier launch
    uint64_t nsel = 0;   // selected blocks so far (same on every lane)
    for (uint64_t base = 0; base < nblocks; base += kChainChunk) {
        const uint32_t cnt = static_cast<uint32_t>(nblocks - base < kChainChunk ? nblocks - base : kChainChunk);
        for (uint32_t i = lane; i < cnt * nc; i += 32) s_fit[i] = c_fit[base * nc + i];
        for (uint32_t i = lane; i < cnt; i += 32) s_sel[i] = sel ? sel[base + i] : 1;   // no selection array: every block
        __syncwarp();
        uint32_t local_sel = 0;
        if (lane < nc) {
            for (uint32_t i = 0; i < cnt; i++) {
                if (!s_sel[i]) continue;
                const T c = s_fit[i * nc + lane];
                T rec;
                const int qv = quantize<T>(c, prev, qp, rec);
                if (qv == 0) {
                    const unsigned long long slot = atomicAdd(n_unpred, 1ull);
                    unpred_pos[slot] = pos_base + (nsel + local_sel) * nc + lane;
                    unpred_val[slot] = c;
                }
                s_q[local_sel * nc + lane] = qv;
                s_rec[i * nc + lane] = rec;
                prev = rec;
                local_sel++;
            }
        }
        local_sel = __shfl_sync(0xffffffffu, local_sel, 0);
        __syncwarp();
        for (uint32_t i = lane; i < local_sel * nc; i += 32) coef_q[nsel * nc + i] = s_q[i];
        for (uint32_t i = lane; i < cnt * nc; i += 32)
            if (s_sel[i / nc]) c_rec[base * nc + i] = s_rec[i];
        nsel += local_sel;
        __syncwarp();
    }
    if (lane == 0) *n_sel_out = nsel;
}

// ---------------------------------------------------------------------------------------------------------------------
// Speculative form of the chain for the dense case (every block selected): one warp per coefficient walks windows of
// 32 blocks.
//   1. every lane guesses the lattice index K of its fitted coefficient relative to the last exact value r_cur
//      (the quantizer snaps to the lattice r_cur + 2*eb*Z, so K does not depend on the blocks in between);
//   2. the guessed steps d = 2*(K_b - K_{b-1})*eb are broadcast and the rounding-exact running value
//      r_b = T(r_{b-1} + d_b) is rebuilt by every lane (one dependent FP64 add per block: the irreducible serial part);
//   3. every lane runs the real quantizer on (c_b, r_{b-1}); lanes up to and including the first one whose index
//      differs from the guess are exact by induction (their predecessor value was exact) and are committed with the
//      REAL quantizer results; the window restarts after them.
// A wrong guess costs one extra window, never a wrong result.
template <class T>
__global__ void __launch_bounds__(32 * (kMaxDim + 1)) k_reg_chain_spec(const T *__restrict__ c_fit, uint64_t nblocks, int N,
                                                                       QuantParams q_liner, QuantParams q_indep,
                                                                       int32_t *__restrict__ coef_q, T *__restrict__ c_rec,
                                                                       unsigned long long *__restrict__ n_sel_out,
                                                                       unsigned long long *__restrict__ n_unpred,
                                                                       unsigned long long *__restrict__ unpred_pos,
                                                                       T *__restrict__ unpred_val) {
    const int d = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nc = N + 1;
    if (d >= nc) return;
    const QuantParams qp = d < N ? q_liner : q_indep;
    const unsigned full = 0xffffffffu;
    T r_cur = 0;
    uint64_t b0 = 0;
    T c_pref = lane < nblocks ? c_fit[static_cast<uint64_t>(lane) * nc + d] : static_cast<T>(0);
    uint64_t pref_b0 = 0;
    while (b0 < nblocks) {
        const uint64_t b = b0 + lane;
        const bool in = b < nblocks;
        const T c = pref_b0 == b0 ? c_pref : (in ? c_fit[b * nc + d] : static_cast<T>(0));
        {   // prefetch the window that follows a fully accepted one
            const uint64_t nb = b0 + 32 + lane;
            c_pref = nb < nblocks ? c_fit[nb * nc + d] : static_cast<T>(0);
            pref_b0 = b0 + 32;
        }
        // 1. signed lattice index relative to r_cur
        const T df = c - r_cur;
        const double v = fabs(static_cast<double>(df)) * qp.ebr;
        const bool sane = v < 1.0e9;
        const int half = sane ? (trunc_to_int(v) + 1) >> 1 : 0;
        const int K = df < 0 ? -half : half;
        int Kp = __shfl_up_sync(full, K, 1);
        if (lane == 0) Kp = 0;
        const int dk = K - Kp;
        const double dd = int_to_double(2 * dk) * qp.eb;
        // 2. running value, rebuilt by every lane; keep r_{lane-1}
        T r = r_cur, pred = r_cur;
#pragma unroll
        for (int i = 0; i < 32; i++) {
            const double di = __shfl_sync(full, dd, i);
            r = static_cast<T>(static_cast<double>(r) + di);
            if (i + 1 == lane) pred = r;
        }
        // 3. the real quantizer against the speculated predecessor
        T rec;
        const int qv = quantize<T>(c, pred, qp, rec);
        const bool ok = sane && qv != 0 && qv == qp.radius + dk;
        const unsigned mism = __ballot_sync(full, in && !ok);
        const unsigned n_in = nblocks - b0 < 32 ? static_cast<unsigned>(nblocks - b0) : 32u;
        const unsigned first_bad = mism ? static_cast<unsigned>(__ffs(mism) - 1) : 32u;
        const unsigned n_acc = first_bad + 1 < n_in ? first_bad + 1 : n_in;
        if (static_cast<unsigned>(lane) < n_acc) {
            coef_q[b * nc + d] = qv;
            c_rec[b * nc + d] = rec;
            if (qv == 0) {
                const unsigned long long slot = atomicAdd(n_unpred, 1ull);
                unpred_pos[slot] = b * nc + d;
                unpred_val[slot] = c;
            }
        }
        r_cur = __shfl_sync(full, rec, n_acc - 1);
        b0 += n_acc;
    }
    if (threadIdx.x == 0) *n_sel_out = nblocks;
}

// ---------------------------------------------------------------------------------------------------------------------
// Same speculation over super-windows of kSW blocks, so that the only serial work per block is the one dependent add:
// lanes own interleaved blocks (i = j*32 + lane), guess all lattice indices, one lane rebuilds the running values
// from shared memory back to back, then all lanes verify their blocks and the prefix up to the first disagreement is
// committed with the real quantizer results.
constexpr int kSW = 512;
constexpr int kSWPer = kSW / 32;

template <class T>
__global__ void __launch_bounds__(32 * (kMaxDim + 1)) k_reg_chain_spec2(const T *__restrict__ c_fit, uint64_t nblocks, int N,
                                                                        QuantParams q_liner, QuantParams q_indep,
                                                                        int32_t *__restrict__ coef_q, T *__restrict__ c_rec,
                                                                        unsigned long long *__restrict__ n_sel_out,
                                                                        unsigned long long *__restrict__ n_unpred,
                                                                        unsigned long long *__restrict__ unpred_pos,
                                                                        T *__restrict__ unpred_val, const T *__restrict__ init,
                                                                        unsigned long long pos_base) {
    __shared__ double s_d[kMaxDim + 1][kSW];
    __shared__ T s_r[kMaxDim + 1][kSW + 1];
    __shared__ T s_last[kMaxDim + 1];
    const int d = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nc = N + 1;
    if (d >= nc) return;
    const QuantParams qp = d < N ? q_liner : q_indep;
    const unsigned full = 0xffffffffu;
    double *sd = s_d[d];
    T *sr = s_r[d];
    T r_cur = init ? init[d] : static_cast<T>(0);   // chain continued from an earlier launch
    uint64_t b0 = 0;
    while (b0 < nblocks) {
        const unsigned n_in = nblocks - b0 < kSW ? static_cast<unsigned>(nblocks - b0) : kSW;
        T c[kSWPer];
        int dk[kSWPer];
        bool sane_all = true;
        // 1. lattice indices relative to r_cur, steps between consecutive blocks
        int Kprev_row = 0;   // K of block (j-1)*32 + 31
#pragma unroll
        for (int j = 0; j < kSWPer; j++) {
            const unsigned i = j * 32 + lane;
            c[j] = i < n_in ? c_fit[(b0 + i) * nc + d] : static_cast<T>(0);
            const T df = c[j] - r_cur;
            const double v = fabs(static_cast<double>(df)) * qp.ebr;
            const bool sane = v < 1.0e9;
            sane_all = sane_all && (sane || i >= n_in);
            const int half = sane ? (trunc_to_int(v) + 1) >> 1 : 0;
            const int K = df < 0 ? -half : half;
            const int up = __shfl_up_sync(full, K, 1);
            const int Kp = lane ? up : Kprev_row;
            Kprev_row = __shfl_sync(full, K, 31);
            dk[j] = K - Kp;
            sd[i] = int_to_double(2 * dk[j]) * qp.eb;
        }
        __syncwarp();
        // 2. the serial part: one dependent add per block
        if (lane == 0) {
            T r = r_cur;
            sr[0] = r;
            // batches of 8: the shared-memory loads of a batch are issued together, so the loop runs at the
            // latency of the dependent adds (entries beyond n_in are unused garbage)
            for (unsigned i0 = 0; i0 < n_in; i0 += 8) {
                double dv[8];
#pragma unroll
                for (int k = 0; k < 8; k++) dv[k] = sd[i0 + k];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    r = static_cast<T>(static_cast<double>(r) + dv[k]);
                    sr[i0 + k + 1] = r;
                }
            }
        }
        __syncwarp();
        // 3. verify every block against its speculated predecessor
        unsigned first_bad = kSW;
        int qv[kSWPer];
        T rec[kSWPer];
#pragma unroll
        for (int j = 0; j < kSWPer; j++) {
            const unsigned i = j * 32 + lane;
            qv[j] = 0;
            rec[j] = 0;
            if (i < n_in) {
                qv[j] = quantize<T>(c[j], sr[i], qp, rec[j]);
                const bool ok = sane_all && qv[j] != 0 && qv[j] == qp.radius + dk[j];
                if (!ok && i < first_bad) first_bad = i;
            }
        }
        first_bad = __reduce_min_sync(full, first_bad);
        const unsigned n_acc = first_bad + 1 < n_in ? first_bad + 1 : n_in;
#pragma unroll
        for (int j = 0; j < kSWPer; j++) {
            const unsigned i = j * 32 + lane;
            if (i < n_acc) {
                const uint64_t pos = (b0 + i) * nc + d;
                coef_q[pos] = qv[j];
                c_rec[pos] = rec[j];
                if (qv[j] == 0) {
                    const unsigned long long slot = atomicAdd(n_unpred, 1ull);
                    unpred_pos[slot] = pos_base + pos;
                    unpred_val[slot] = c[j];
                }
                if (i + 1 == n_acc) s_last[d] = rec[j];
            }
        }
        __syncwarp();
        r_cur = s_last[d];
        b0 += n_acc;
        __syncwarp();
    }
    if (threadIdx.x == 0) *n_sel_out = nblocks;
}

// ---------------------------------------------------------------------------------------------------------------------
// One CTA = one chunk of kRegChunk consecutive elements of one row (a row = all coordinates but the fastest fixed).
// Row coordinates come from the grid indices, so no thread ever divides a 64-bit element index.
constexpr int kRegThreads = 128;
constexpr int kRegChunk = 512;
constexpr int kRegRows = 16;   // consecutive rows (along the second-fastest dim) per CTA: amortises the histogram window

template <class T, class QT>
__global__ void __launch_bounds__(kRegThreads) k_reg_predict(const T *__restrict__ data, BlockShape bs, uint32_t nchunks,
                                                            uint32_t mgB, const T *__restrict__ c_rec, QuantParams qp,
                                                            QT *__restrict__ q, T *__restrict__ unpred_tmp,
                                                            unsigned long long *__restrict__ hist) {
    __shared__ unsigned shist[kHistWindow];
    DevCtx2 ctx(shist, hist, qp.radius);
    ctx.clear();
    const int N = bs.N;
    uint32_t xr[kMaxDim] = {0, 0, 0, 0};
    uint32_t chunk = blockIdx.x, row0 = 0, nrows = 1;
    if (N >= 2) {
        const uint32_t rg = blockIdx.x / nchunks;
        chunk = blockIdx.x - rg * nchunks;
        row0 = rg * kRegRows;
        nrows = bs.dims[N - 2] - row0 < static_cast<uint32_t>(kRegRows) ? bs.dims[N - 2] - row0 : kRegRows;
    }
    if (N >= 3) xr[N - 3] = blockIdx.y;
    if (N >= 4) xr[N - 4] = blockIdx.z;
    const uint32_t len = bs.dims[N - 1];
    const uint32_t x0 = chunk * kRegChunk;
    const int nc = N + 1;
    for (uint32_t r = 0; r < nrows; r++) {
        if (N >= 2) xr[N - 2] = row0 + r;
        RegRow rr;
        reg_row_setup(bs, xr, rr);
        uint64_t row_off = 0;
        for (int d = 0; d < N - 1; d++) row_off += xr[d] * bs.stride[d];
#pragma unroll
        for (int k = 0; k < kRegChunk / kRegThreads; k++) {
            const uint32_t x = x0 + k * kRegThreads + threadIdx.x;
            const bool active = x < len;
            int qv = 0;
            if (active) {
                uint64_t blin, pos;
                uint32_t li[kMaxDim] = {rr.li[0], rr.li[1], rr.li[2], rr.li[3]};
                reg_row_locate(bs, rr, x, mgB, &blin, &li[N - 1], &pos);
                const T pred = reg_predict<T>(N, c_rec + blin * nc, li);
                const T orig = data[row_off + x];
                T rec;
                qv = quantize<T>(orig, pred, qp, rec);
                q[pos] = static_cast<QT>(qv);
                if (qv == 0) unpred_tmp[pos] = orig;
            }
            ctx.hist_add(qv, active);
        }
    }
    ctx.pass_end();   // at most kRegRows * kRegChunk / kRegThreads = 64 points per thread: the 8-bit counters hold
    ctx.flush();
}

// ---------------------------------------------------------------------------------------------------------------------
template <class T>
void launch_reg_fit(const T *data, const BlockShape &bs, T *c_fit, uint8_t *valid, cudaStream_t st) {
    const unsigned grid = static_cast<unsigned>((bs.nblocks + 127) / 128);
    k_reg_fit<T><<<grid, 128, 0, st>>>(data, bs, c_fit, valid);
}
template <class T>
void launch_reg_chain(const T *c_fit, const uint8_t *sel, uint64_t nblocks, int N, const QuantParams &q_liner,
                      const QuantParams &q_indep, int32_t *coef_q, T *c_rec, unsigned long long *counters,
                      unsigned long long *unpred_pos, T *unpred_val, cudaStream_t st, const T *init,
                      unsigned long long pos_base) {
    // dense (every block selected): the speculative chain -- its lattice arithmetic is floating point; integer element
    // types walk the chain block by block with the plain quantizer
    if (sel == nullptr && std::is_floating_point<T>::value)
        k_reg_chain_spec2<T><<<1, 32 * (N + 1), 0, st>>>(c_fit, nblocks, N, q_liner, q_indep, coef_q, c_rec, counters,
                                                        counters + 1, unpred_pos, unpred_val, init, pos_base);
    else
        k_reg_chain<T><<<1, 32, 0, st>>>(c_fit, sel, nblocks, N, q_liner, q_indep, coef_q, c_rec, counters, counters + 1,
                                        unpred_pos, unpred_val, sel ? nullptr : init, sel ? 0ull : pos_base);
}
template <class T, class QT>
const char *launch_reg_predict(const T *data, const BlockShape &bs, const T *c_rec, const QuantParams &qp, QT *q,
                               T *unpred_tmp, unsigned long long *hist, cudaStream_t st) {
    const int N = bs.N;
    const uint32_t len = bs.dims[N - 1];
    const uint32_t nchunks = (len + kRegChunk - 1) / kRegChunk;
    const uint64_t gx = static_cast<uint64_t>(nchunks) * (N >= 2 ? (bs.dims[N - 2] + kRegRows - 1) / kRegRows : 1);
    const uint32_t gy = N >= 3 ? bs.dims[N - 3] : 1, gz = N >= 4 ? bs.dims[N - 4] : 1;
    if (gx > 0x7fffffffull || gy > 65535u || gz > 65535u) return "array shape exceeds the launch grid of the regression kernel";
    // multiply-high division by B is exact while x * B < 2^32
    const uint32_t mgB = (bs.B > 1 && static_cast<uint64_t>(len) * bs.B < (1ull << 32)) ? 0xffffffffu / bs.B + 1u : 0u;
    dim3 grid(static_cast<unsigned>(gx), gy, gz);
    k_reg_predict<T, QT><<<grid, kRegThreads, 0, st>>>(data, bs, nchunks, mgB, c_rec, qp, q, unpred_tmp, hist);
    return nullptr;
}

#define SZ3B_INST_BW(T)                                                                                              \
    template void launch_reg_fit<T>(const T *, const BlockShape &, T *, uint8_t *, cudaStream_t);                    \
    template void launch_reg_chain<T>(const T *, const uint8_t *, uint64_t, int, const QuantParams &,                \
                                      const QuantParams &, int32_t *, T *, unsigned long long *,                    \
                                      unsigned long long *, T *, cudaStream_t, const T *, unsigned long long);       \
    template const char *launch_reg_predict<T, uint16_t>(const T *, const BlockShape &, const T *,                  \
                                                         const QuantParams &, uint16_t *, T *, unsigned long long *, \
                                                         cudaStream_t);                                              \
    template const char *launch_reg_predict<T, uint32_t>(const T *, const BlockShape &, const T *,                  \
                                                         const QuantParams &, uint32_t *, T *, unsigned long long *, \
                                                         cudaStream_t);
SZ3B_INST_BW(float)
SZ3B_INST_BW(double)
SZ3B_INST_BW(int32_t)
SZ3B_INST_BW(int64_t)

}  // namespace sz3b
