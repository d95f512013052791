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
                                                  unsigned long long *__restrict__ unpred_pos, T *__restrict__ unpred_val) {
    __shared__ T s_fit[kChainChunk * (kMaxDim + 1)];
    __shared__ T s_rec[kChainChunk * (kMaxDim + 1)];
    __shared__ int32_t s_q[kChainChunk * (kMaxDim + 1)];
    __shared__ uint8_t s_sel[kChainChunk];
    const int lane = threadIdx.x;
    const int nc = N + 1;
    const QuantParams qp = lane < N ? q_liner : q_indep;
    T prev = 0;
    uint64_t nsel = 0;   // selected blocks so far (same on every lane)
    for (uint64_t base = 0; base < nblocks; base += kChainChunk) {
        const uint32_t cnt = static_cast<uint32_t>(nblocks - base < kChainChunk ? nblocks - base : kChainChunk);
        for (uint32_t i = lane; i < cnt * nc; i += 32) s_fit[i] = c_fit[base * nc + i];
        for (uint32_t i = lane; i < cnt; i += 32) s_sel[i] = sel[base + i];
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
                    unpred_pos[slot] = (nsel + local_sel) * nc + lane;
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
template <class T, class QT>
__global__ void __launch_bounds__(256) k_reg_predict(const T *__restrict__ data, BlockShape bs,
                                                     const T *__restrict__ c_rec, QuantParams qp, QT *__restrict__ q,
                                                     T *__restrict__ unpred_tmp, unsigned long long *__restrict__ hist) {
    __shared__ unsigned shist[kHistWindow];
    DevCtx ctx(shist, hist, qp.radius);
    ctx.clear();
    const uint64_t gid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool active = gid < bs.num;
    int qv = 0;
    if (active) {
        uint64_t blin, pos;
        uint32_t li[kMaxDim];
        reg_locate(bs, gid, &blin, li, &pos);
        const T pred = reg_predict<T>(bs.N, c_rec + blin * (bs.N + 1), li);
        const T orig = data[gid];
        T rec;
        qv = quantize<T>(orig, pred, qp, rec);
        q[pos] = static_cast<QT>(qv);
        if (qv == 0) unpred_tmp[pos] = orig;
    }
    ctx.hist_add(qv, active);
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
                      unsigned long long *unpred_pos, T *unpred_val, cudaStream_t st) {
    k_reg_chain<T><<<1, 32, 0, st>>>(c_fit, sel, nblocks, N, q_liner, q_indep, coef_q, c_rec, counters, counters + 1,
                                    unpred_pos, unpred_val);
}
template <class T, class QT>
void launch_reg_predict(const T *data, const BlockShape &bs, const T *c_rec, const QuantParams &qp, QT *q, T *unpred_tmp,
                        unsigned long long *hist, cudaStream_t st) {
    const unsigned grid = static_cast<unsigned>((bs.num + 255) / 256);
    k_reg_predict<T, QT><<<grid, 256, 0, st>>>(data, bs, c_rec, qp, q, unpred_tmp, hist);
}

#define SZ3B_INST_BW(T)                                                                                              \
    template void launch_reg_fit<T>(const T *, const BlockShape &, T *, uint8_t *, cudaStream_t);                    \
    template void launch_reg_chain<T>(const T *, const uint8_t *, uint64_t, int, const QuantParams &,                \
                                      const QuantParams &, int32_t *, T *, unsigned long long *,                    \
                                      unsigned long long *, T *, cudaStream_t);                                      \
    template void launch_reg_predict<T, uint16_t>(const T *, const BlockShape &, const T *, const QuantParams &,     \
                                                  uint16_t *, T *, unsigned long long *, cudaStream_t);             \
    template void launch_reg_predict<T, uint32_t>(const T *, const BlockShape &, const T *, const QuantParams &,     \
                                                  uint32_t *, T *, unsigned long long *, cudaStream_t);
SZ3B_INST_BW(float)
SZ3B_INST_BW(double)

}  // namespace sz3b
