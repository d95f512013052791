// sz3_b200/csrc/lorenzo.cuh -- one block of BlockwiseDecomposition with a Lorenzo predictor in the stack (reference
// include/SZ3/decomposition/BlockwiseDecomposition.hpp:28-67, predictor/LorenzoPredictor.hpp:17-94,
// predictor/ComposedPredictor.hpp:25-50, utils/BlockwiseIterator.hpp:103-184).
//
// Dependencies of the reference's sequential walk: a Lorenzo prediction reads the RECONSTRUCTED values of the lower
// neighbours, so a block needs the finished reconstruction of its (up to 2^N - 1) lower neighbour blocks, and a point
// needs its lower neighbours inside the block.  Both are hyperplane wavefronts: blocks with the same coordinate sum
// are independent (one CTA each, one kernel launch per front), points of a block with the same index sum are
// independent (the lanes of the CTA's warp, one barrier per diagonal).  The working array is the reference's
// zero-padded copy (2 cells on the low side of every dimension, BlockwiseIterator.hpp:205-222), modified in place.
//
// With the regression predictor in the stack the reconstructed coefficients form a serial chain over the SELECTED
// blocks in row-major order (RegressionPredictor.hpp:57-60), which is not a wavefront order.  pipeline.cu therefore
// iterates  selection guess -> exact chain over the guessed selection -> exact wavefront pass that re-derives the
// selection  until the selection reproduces itself; the first block (row-major) whose guess was wrong is always
// corrected by a pass, so the fixed point is the reference's result.
//
// Plain inline code: the same body runs on the device (one warp per block) and in tests/emul (nl = 1).
#pragma once
#include <float.h>

#include "blockwise.cuh"

namespace sz3b {

constexpr int kBwPad = 2;   // LorenzoPredictor::get_padding (LorenzoPredictor.hpp:54)
enum { BW_SPEC = 0, BW_EXACT = 1, BW_DECODE = 2, BW_SERIAL = 3 };
enum { PK_LORENZO1 = 0, PK_LORENZO2 = 1, PK_REG = 2 };

#if defined(__CUDA_ARCH__)
#define SZ_WARP_SYNC() __syncwarp()
#else
#define SZ_WARP_SYNC() ((void)0)
#endif

template <class T, class QT>
struct BwArgs {
    BlockShape bs;
    uint64_t pstride[kMaxDim];   // element strides of the padded working array (dims[d] + 2 per dimension)
    T *W;                        // padded working array; element (0,..,0) of the data sits at +2 in every dimension
    T *out;                      // BW_DECODE: the caller's (unpadded) output array
    QuantParams qp;
    T noise[2];                  // estimate_error noise of the first / second order Lorenzo predictor
    int kinds[3];                // enabled predictors in the reference's order (SZAlgoLorenzoReg.hpp:33-61)
    int nk;
    int mode;
    const T *c_fit;              // unquantized fit per block [nblocks * (N+1)] (regression estimate_error)
    const uint8_t *fit_valid;
    const T *c_spec;             // lattice guess of the reconstructed coefficients per block
    const T *c_rec;              // exact reconstructed coefficients, dense over the regression-selected blocks
    const uint32_t *rank;        // dense rank of a block among the regression-selected ones (per sel_in)
    const uint8_t *sel_in;       // BW_EXACT: the guessed selection (may be null); BW_DECODE: the stream's selection
    uint8_t *sel_out;            // BW_SPEC / BW_EXACT: the selection derived by this pass
    unsigned *mismatch;          // BW_EXACT: [0] number of blocks whose derived selection differs from sel_in,
                                 //           [1] the first such block in row-major order (atomicMin; init ~0u)
    QT *q;                       // quantization indices, block-major (written by BW_EXACT, read by BW_DECODE)
    T *unpred_tmp;               // position-indexed unpredictable values (same direction as q)
    uint64_t b_lo, b_hi;         // window of row-major block indices processed by a front launch
    // batch of independent arrays of identical shape (the tuner's sampled blocks): strides between batch members
    uint32_t nbatch;
    uint64_t w_bstride, q_bstride, sel_bstride;
    // points of a FULL block (every extent == B) ordered by index sum: entry = tile offset (relative to the block's
    // first point) << 16 | row-major rank inside the block; diag_start[d] .. diag_start[d + 1] = diagonal d
    const uint32_t *diag_tab;
    const uint16_t *diag_start;
    const uint32_t *diag_idx;    // same order: the point's in-block index, 8 bits per dimension (dimension d at bits 8d)
    // BW_SERIAL (row-major walk with the coefficient chain inline): chain outputs, dense over the selected blocks
    QuantParams q_liner, q_indep;
    int32_t *coef_q;
    T *c_rec_out;                // reconstructed coefficients (same dense array the exact passes read as c_rec)
    unsigned long long *n_unpred_coef, *unpred_pos;   // unpredictable coefficients: count, dense position
    T *unpred_val;
};

// chain state of the row-major walk (shared memory of the walking CTA)
template <class T>
struct BwSerial {
    T prev[kMaxDim + 1];          // reconstructed coefficients of the last selected block (prev_coeffs)
    unsigned long long nsel;      // regression-selected blocks so far
};

// LorenzoPredictor constructor (LorenzoPredictor.hpp:17-38): noise is stored as T
template <class T>
SZ_HD T lorenzo_noise(int N, int L, double eb) {
    const double n1[4] = {0.5, 0.81, 1.22, 1.79}, n2[3] = {1.08, 2.76, 6.8};
    if (L == 1) return static_cast<T>(n1[N - 1] * eb);
    if (N <= 3) return static_cast<T>(n2[N - 1] * eb);
    return static_cast<T>(0);
}

// LorenzoPredictor::predict (LorenzoPredictor.hpp:60-94) on a tile with strides ts[] (ts[N-1] == 1); the reference's
// prevN helpers pair their arguments with the strides in the order (.., ds[1], ds[0], 1) -- kept as is, the
// summation order decides the bits.
template <class T>
SZ_HD T lorenzo_predict_tile(const T *d, int N, int L, const uint32_t ts[kMaxDim]) {
#define P1(i) (*(d - (i)))
#define P2(j, i) (*(d - ((j) * s0 + (i))))
#define P3(k, j, i) (*(d - ((k) * s1 + (j) * s0 + (i))))
#define P4(t, k, j, i) (*(d - ((t) * s2 + (k) * s1 + (j) * s0 + (i))))
    const int s0 = static_cast<int>(ts[0]), s1 = static_cast<int>(ts[1]), s2 = static_cast<int>(ts[2]);
    const T c2 = static_cast<T>(2), c4 = static_cast<T>(4), c8 = static_cast<T>(8);
    if (L == 1) {
        if (N == 1) return P1(1);
        if (N == 2) return P2(0, 1) + P2(1, 0) - P2(1, 1);
        if (N == 3) return P3(0, 0, 1) + P3(0, 1, 0) + P3(1, 0, 0) - P3(0, 1, 1) - P3(1, 0, 1) - P3(1, 1, 0) + P3(1, 1, 1);
        return P4(0, 0, 0, 1) + P4(0, 0, 1, 0) - P4(0, 0, 1, 1) + P4(0, 1, 0, 0) - P4(0, 1, 0, 1) - P4(0, 1, 1, 0) +
               P4(0, 1, 1, 1) + P4(1, 0, 0, 0) - P4(1, 0, 0, 1) - P4(1, 0, 1, 0) + P4(1, 0, 1, 1) - P4(1, 1, 0, 0) +
               P4(1, 1, 0, 1) + P4(1, 1, 1, 0) - P4(1, 1, 1, 1);
    }
    if (N == 1) return c2 * P1(1) - P1(2);
    if (N == 2)
        return c2 * P2(0, 1) - P2(0, 2) + c2 * P2(1, 0) - c4 * P2(1, 1) + c2 * P2(1, 2) - P2(2, 0) + c2 * P2(2, 1) -
               P2(2, 2);
    if (N == 3)
        return c2 * P3(0, 0, 1) - P3(0, 0, 2) + c2 * P3(0, 1, 0) - c4 * P3(0, 1, 1) + c2 * P3(0, 1, 2) - P3(0, 2, 0) +
               c2 * P3(0, 2, 1) - P3(0, 2, 2) + c2 * P3(1, 0, 0) - c4 * P3(1, 0, 1) + c2 * P3(1, 0, 2) -
               c4 * P3(1, 1, 0) + c8 * P3(1, 1, 1) - c4 * P3(1, 1, 2) + c2 * P3(1, 2, 0) - c4 * P3(1, 2, 1) +
               c2 * P3(1, 2, 2) - P3(2, 0, 0) + c2 * P3(2, 0, 1) - P3(2, 0, 2) + c2 * P3(2, 1, 0) - c4 * P3(2, 1, 1) +
               c2 * P3(2, 1, 2) - P3(2, 2, 0) + c2 * P3(2, 2, 1) - P3(2, 2, 2);
    return static_cast<T>(0);   // 4-D second order: the reference's predict() returns T(0) as well (:91-92)
#undef P1
#undef P2
#undef P3
#undef P4
}

// Lattice guess of a reconstructed regression coefficient: what the quantizer chain started at 0 would produce if
// no rounding error accumulated (RegressionPredictor.hpp:148-155 with prev = 0).
template <class T>
SZ_HD T coef_lattice_guess(T c, const QuantParams &qp) {
    const double v = fabs(static_cast<double>(c)) * qp.ebr;
    if (!(v < 1.0e9)) return c;
    const int half = (trunc_to_int(v) + 1) >> 1;
    const int K2 = c < 0 ? -(half << 1) : (half << 1);
    return static_cast<T>(0.0 + int_to_double(K2) * qp.eb);
}

// Guess for a block that is NOT in the chain: the value the chain would give it if it were inserted after its
// `before` selected predecessors -- exact whenever the selection of all earlier blocks is already right.
template <class T>
SZ_HD void bw_respec_block(const T *c_fit, const T *c_rec_dense, uint64_t b, uint32_t before, int nc, const QuantParams &q_liner,
                           const QuantParams &q_indep, T *c_spec) {
    for (int d = 0; d < nc; d++) {
        const T prev = before ? c_rec_dense[static_cast<uint64_t>(before - 1) * nc + d] : static_cast<T>(0);
        T rec;
        quantize<T>(c_fit[b * nc + d], prev, d < nc - 1 ? q_liner : q_indep, rec);
        c_spec[b * nc + d] = rec;
    }
}

struct BwGeom {
    uint32_t lo[kMaxDim], ext[kMaxDim], ts[kMaxDim];
    uint64_t b, pos0, wbase, obase;
    uint32_t npts, tile_size;
};

template <class T, class QT>
SZ_HD void bw_geometry(const BwArgs<T, QT> &A, const uint32_t bi[kMaxDim], BwGeom &g) {
    const BlockShape &bs = A.bs;
    const int N = bs.N;
    uint64_t extprod = 1;
    g.b = 0;
    g.pos0 = 0;
    g.wbase = 0;
    g.obase = 0;
    for (int d = 0; d < kMaxDim; d++) {
        g.lo[d] = 0;
        g.ext[d] = 1;
        g.ts[d] = 0;
    }
    for (int d = 0; d < N; d++) {
        g.lo[d] = bi[d] * bs.B;
        g.ext[d] = bs.dims[d] - g.lo[d] < bs.B ? bs.dims[d] - g.lo[d] : bs.B;
        g.b = g.b * bs.nb[d] + bi[d];
        g.pos0 += extprod * g.lo[d] * bs.stride[d];   // closed form of reg_locate (blockwise.cuh); stride = trailing extent product
        extprod *= g.ext[d];
        g.wbase += static_cast<uint64_t>(g.lo[d]) * A.pstride[d];
        g.obase += static_cast<uint64_t>(g.lo[d]) * bs.stride[d];
    }
    g.npts = static_cast<uint32_t>(extprod);
    uint32_t acc = 1;
    for (int d = N - 1; d >= 0; d--) {
        g.ts[d] = acc;
        acc *= g.ext[d] + kBwPad;
    }
    g.tile_size = acc;
}

// One block.  tile: (ext+2)^N elements, est: nk * (sample points) elements.  Lanes lane, lane+nl, ... of one warp.
template <class T, class QT>
SZ_HD void bw_process_block(const BwArgs<T, QT> &A, const uint32_t bi[kMaxDim], T *tile, T *est, int lane, int nl,
                            BwSerial<T> *chain = nullptr) {
    const BlockShape &bs = A.bs;
    const int N = bs.N, nc = N + 1;
    BwGeom g;
    bw_geometry(A, bi, g);
    // ---- 1. tile = halo (finished reconstruction of the lower neighbour blocks, or the zero padding) + the block;
    //         row by row, four independent loads in flight per lane
    {
        const uint32_t rowlen = g.ext[N - 1] + kBwPad, nrows = g.tile_size / rowlen;
        for (uint32_t row = lane; row < nrows; row += nl) {
            uint32_t r = row;
            uint64_t w = g.wbase;
            for (int d = N - 2; d >= 0; d--) {
                const uint32_t te = g.ext[d] + kBwPad;
                w += static_cast<uint64_t>(r % te) * A.pstride[d];
                r /= te;
            }
            const T *src = A.W + w;
            T *dst = tile + row * rowlen;
            uint32_t x = 0;
            for (; x + 4 <= rowlen; x += 4) {
                const T v0 = src[x], v1 = src[x + 1], v2 = src[x + 2], v3 = src[x + 3];
                dst[x] = v0;
                dst[x + 1] = v1;
                dst[x + 2] = v2;
                dst[x + 3] = v3;
            }
            for (; x < rowlen; x++) dst[x] = src[x];
        }
    }
    SZ_WARP_SYNC();
    // offset of in-block point (0,..,0) inside the tile
    uint32_t t00 = 0;
    for (int d = 0; d < N; d++) t00 += kBwPad * g.ts[d];
    // ---- 2. predictor selection (ComposedPredictor::precompress, :25-45)
    int sid = 0;
    if (A.nk > 1) {
        if (A.mode == BW_DECODE) {
            sid = A.sel_in[g.b];
        } else {
            uint32_t m = g.ext[0];
            for (int d = 1; d < N; d++) m = g.ext[d] < m ? g.ext[d] : m;
            const uint32_t P = N == 1 ? 2u : m << (N - 1);
            const bool reg_ok = A.fit_valid ? A.fit_valid[g.b] != 0 : false;
            for (uint32_t item = lane; item < P * A.nk; item += nl) {
                const uint32_t k = item / P, p = item - k * P;
                uint32_t idx[kMaxDim] = {0, 0, 0, 0};
                if (N == 1) {
                    idx[0] = p ? m - 1 : 0;
                } else {   // foreach_sampling (BlockwiseIterator.hpp:150-184)
                    const uint32_t i = p >> (N - 1), cb = p & ((1u << (N - 1)) - 1), j = m - 1 - i;
                    idx[0] = i;
                    for (int d = 1; d < N; d++) idx[d] = ((cb >> (N - 1 - d)) & 1u) ? j : i;
                }
                uint32_t off = t00;
                for (int d = 0; d < N; d++) off += idx[d] * g.ts[d];
                const T *ptr = tile + off;
                const int kind = A.kinds[k];
                T e = 0;
                if (kind == PK_REG) {
                    if (reg_ok) e = static_cast<T>(fabs(static_cast<double>(static_cast<T>(*ptr - reg_predict<T>(N, A.c_fit + g.b * nc, idx)))));
                } else {
                    e = static_cast<T>(fabs(static_cast<double>(static_cast<T>(*ptr - lorenzo_predict_tile<T>(ptr, N, kind + 1, g.ts)))));
                    e = e + A.noise[kind];
                }
                est[item] = e;
            }
            SZ_WARP_SYNC();
            // double sums in sample order (ComposedPredictor.hpp:31-33); with a warp, lane k sums predictor k
            double best = 0, mine = 0;
            const int k_lo = nl > 1 ? lane : 0, k_hi = nl > 1 ? lane + 1 : A.nk;
            double errs[3] = {0, 0, 0};
            for (int k = k_lo; k < k_hi && k < A.nk; k++) {
                double err = 0;
                if (A.kinds[k] == PK_REG && !reg_ok) {
                    err = DBL_MAX;
                } else {
                    for (uint32_t p = 0; p < P; p++) err += static_cast<double>(est[k * P + p]);
                }
                errs[k < 3 ? k : 0] = err;
                mine = err;
            }
            for (int k = 0; k < A.nk; k++) {
                double err = errs[k];
#if defined(__CUDA_ARCH__)
                if (nl > 1) err = __shfl_sync(0xffffffffu, mine, k);
#endif
                if (k == 0 || err < best) {   // std::min_element: the first minimum wins
                    best = err;
                    sid = k;
                }
            }
        }
    }
    const int use = A.kinds[sid];
    const bool emit = A.mode == BW_EXACT || A.mode == BW_SERIAL;
    if (lane == 0 && A.mode != BW_DECODE && A.nk > 1) {
        if (A.mode == BW_EXACT && A.sel_in && A.sel_in[g.b] != sid) {
#if defined(__CUDA_ARCH__)
            atomicAdd(&A.mismatch[0], 1u);
            atomicMin(&A.mismatch[1], static_cast<unsigned>(g.b));
#else
            A.mismatch[0] += 1u;
            if (static_cast<unsigned>(g.b) < A.mismatch[1]) A.mismatch[1] = static_cast<unsigned>(g.b);
#endif
        }
        A.sel_out[g.b] = static_cast<uint8_t>(sid);
    }
    // ---- 3. predict + quantize (or recover), in place in the tile
    const uint32_t extL = g.ext[N - 1];
    if (use == PK_REG) {
        T cf[kMaxDim + 1];
        if (A.mode == BW_SERIAL) {
            // precompress_block_commit (RegressionPredictor.hpp:57-60,148-155), redundantly on every lane
            const unsigned long long at = chain->nsel * nc;
            for (int d = 0; d < nc; d++) {
                const T c = A.c_fit[g.b * nc + d];
                const int qv = quantize<T>(c, chain->prev[d], d < N ? A.q_liner : A.q_indep, cf[d]);
                if (lane == 0) {
                    A.coef_q[at + d] = qv;
                    A.c_rec_out[at + d] = cf[d];
                    if (qv == 0) {
#if defined(__CUDA_ARCH__)
                        const unsigned long long slot = atomicAdd(A.n_unpred_coef, 1ull);
#else
                        const unsigned long long slot = (*A.n_unpred_coef)++;
#endif
                        A.unpred_pos[slot] = at + d;
                        A.unpred_val[slot] = c;
                    }
                }
            }
            SZ_WARP_SYNC();
            if (lane == 0) {
                for (int d = 0; d < nc; d++) chain->prev[d] = cf[d];
                chain->nsel++;
            }
        } else {
            const T *coef;
            if (A.mode == BW_SPEC)
                coef = A.c_spec + g.b * nc;
            else if (A.mode == BW_EXACT)
                coef = (A.sel_in && A.kinds[A.sel_in[g.b]] == PK_REG) ? A.c_rec + static_cast<uint64_t>(A.rank[g.b]) * nc
                                                                      : A.c_spec + g.b * nc;
            else
                coef = A.c_rec + static_cast<uint64_t>(A.rank[g.b]) * nc;
            for (int d = 0; d < nc; d++) cf[d] = coef[d];
        }
        bool fullr = A.diag_tab != nullptr;
        for (int d = 0; d < N; d++) fullr = fullr && g.ext[d] == bs.B;
        for (uint32_t it = lane; it < g.npts; it += nl) {
            uint32_t idx[kMaxDim] = {0, 0, 0, 0};
            uint32_t e = it, off = t00;
            if (fullr) {   // table order (any order is fine here: the points of a regression block are independent)
                const uint32_t te = A.diag_tab[it], pk = A.diag_idx[it];
                off += te >> 16;
                e = te & 0xffffu;
                for (int d = 0; d < N; d++) idx[d] = (pk >> (8 * d)) & 0xffu;
            } else {
                uint32_t r = it;
                for (int d = N - 1; d >= 0; d--) {
                    idx[d] = r % g.ext[d];
                    r /= g.ext[d];
                    off += idx[d] * g.ts[d];
                }
            }
            const T pred = reg_predict<T>(N, cf, idx);
            if (A.mode == BW_DECODE) {
                const int qv = static_cast<int>(A.q[g.pos0 + e]);
                tile[off] = qv ? recover_pred<T>(pred, qv, A.qp) : A.unpred_tmp[g.pos0 + e];
            } else {
                const T orig = tile[off];
                T rec;
                const int qv = quantize<T>(orig, pred, A.qp, rec);
                tile[off] = rec;
                if (emit) {
                    A.q[g.pos0 + e] = static_cast<QT>(qv);
                    if (qv == 0) A.unpred_tmp[g.pos0 + e] = orig;
                }
            }
        }
    } else {
        const int L = use + 1;
        bool full = A.diag_tab != nullptr;
        for (int d = 0; d < N; d++) full = full && g.ext[d] == bs.B;
        uint32_t ndiag = 1;
        for (int d = 0; d < N; d++) ndiag += g.ext[d] - 1;
        const uint32_t nlead = g.npts / extL;   // index tuples over the dimensions before the last
        for (uint32_t diag = 0; diag < ndiag; diag++) {
            const uint32_t i0 = full ? A.diag_start[diag] : 0u, i1 = full ? A.diag_start[diag + 1] : nlead;
            for (uint32_t item = i0 + lane; item < i1; item += nl) {
                uint32_t off, within;
                if (full) {
                    const uint32_t e = A.diag_tab[item];
                    off = t00 + (e >> 16);
                    within = e & 0xffffu;
                } else {
                    uint32_t r = item, s = 0;
                    off = t00;
                    for (int d = N - 2; d >= 0; d--) {
                        const uint32_t i = r % g.ext[d];
                        r /= g.ext[d];
                        s += i;
                        off += i * g.ts[d];
                    }
                    if (diag < s || diag - s >= extL) continue;
                    const uint32_t last = diag - s;
                    off += last;
                    within = item * extL + last;
                }
                const T pred = lorenzo_predict_tile<T>(tile + off, N, L, g.ts);
                if (A.mode == BW_DECODE) {
                    const int qv = static_cast<int>(A.q[g.pos0 + within]);
                    tile[off] = qv ? recover_pred<T>(pred, qv, A.qp) : A.unpred_tmp[g.pos0 + within];
                } else {
                    const T orig = tile[off];
                    T rec;
                    const int qv = quantize<T>(orig, pred, A.qp, rec);
                    tile[off] = rec;
                    if (emit) {
                        A.q[g.pos0 + within] = static_cast<QT>(qv);
                        if (qv == 0) A.unpred_tmp[g.pos0 + within] = orig;
                    }
                }
            }
            SZ_WARP_SYNC();
        }
    }
    SZ_WARP_SYNC();
    // ---- 4. the block's reconstruction goes back to the working array (halo of the blocks of later fronts)
    bool fullw = A.diag_tab != nullptr;
    for (int d = 0; d < N; d++) fullw = fullw && g.ext[d] == bs.B;
    for (uint32_t e = lane; e < g.npts; e += nl) {
        uint32_t off = t00;
        uint64_t w = g.wbase, o = g.obase;
        if (fullw) {
            const uint32_t pk = A.diag_idx[e];
            off += A.diag_tab[e] >> 16;
            for (int d = 0; d < N; d++) {
                const uint32_t i = (pk >> (8 * d)) & 0xffu;
                w += static_cast<uint64_t>(i + kBwPad) * A.pstride[d];
                o += static_cast<uint64_t>(i) * bs.stride[d];
            }
        } else {
            uint32_t r = e;
            for (int d = N - 1; d >= 0; d--) {
                const uint32_t i = r % g.ext[d];
                r /= g.ext[d];
                off += i * g.ts[d];
                w += static_cast<uint64_t>(i + kBwPad) * A.pstride[d];
                o += static_cast<uint64_t>(i) * bs.stride[d];
            }
        }
        const T v = tile[off];
        A.W[w] = v;
        if (A.mode == BW_DECODE) A.out[o] = v;
    }
    SZ_WARP_SYNC();
}

// Diagonal table of a full block (see BwArgs::diag_tab).  tab: B^N entries, start: N * (B - 1) + 2 entries.  Returns
// false when the offsets do not fit the 16-bit fields (the generic enumeration is used then).
SZ_HD bool bw_build_diag_table(int N, uint32_t B, uint32_t *tab, uint16_t *start, uint32_t *idx_tab) {
    uint64_t npts = 1, tile = 1;
    for (int d = 0; d < N; d++) {
        npts *= B;
        tile *= B + kBwPad;
    }
    if (npts > 0xffffu || tile > 0xffffu || B > 255u) return false;
    uint32_t ts[kMaxDim] = {0, 0, 0, 0};
    uint32_t acc = 1;
    for (int d = N - 1; d >= 0; d--) {
        ts[d] = acc;
        acc *= B + kBwPad;
    }
    const uint32_t ndiag = static_cast<uint32_t>(N) * (B - 1) + 1;
    uint32_t at = 0;
    for (uint32_t diag = 0; diag < ndiag; diag++) {
        start[diag] = static_cast<uint16_t>(at);
        for (uint32_t e = 0; e < npts; e++) {   // row-major order inside a diagonal
            uint32_t r = e, s = 0, off = 0, packed = 0;
            for (int d = N - 1; d >= 0; d--) {
                const uint32_t i = r % B;
                r /= B;
                s += i;
                off += i * ts[d];
                packed |= i << (8 * d);
            }
            if (s == diag) {
                idx_tab[at] = packed;
                tab[at++] = (off << 16) | e;
            }
        }
    }
    start[ndiag] = static_cast<uint16_t>(at);
    return true;
}

// number of block fronts (hyperplanes of constant block-coordinate sum)
SZ_HD uint32_t bw_num_fronts(const BlockShape &bs) {
    uint32_t f = 1;
    for (int d = 0; d < bs.N; d++) f += bs.nb[d] - 1;
    return f;
}

// shared memory of one block: tile + estimate scratch, in elements of T
SZ_HD size_t bw_scratch_elems(const BlockShape &bs, int nk) {
    size_t tile = 1;
    for (int d = 0; d < bs.N; d++) tile *= (bs.dims[d] < bs.B ? bs.dims[d] : bs.B) + kBwPad;
    uint32_t m = bs.B;
    const size_t P = bs.N == 1 ? 2 : static_cast<size_t>(m) << (bs.N - 1);
    return tile + P * static_cast<size_t>(nk > 1 ? nk : 0) + 8;
}

}  // namespace sz3b
