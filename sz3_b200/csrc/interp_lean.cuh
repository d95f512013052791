// sz3_b200/csrc/interp_lean.cuh -- row-mapped per-pass schedule of the interpolation predict+quantize (and recover)
// loops of InterpolationDecomposition (reference include/SZ3/decomposition/InterpolationDecomposition.hpp:26-76,
// :79-147, :309-454), for N >= 3 (the N <= 2 variant :247-293 keeps the point-mapped kernels of interp_body.cuh).
//
// Same "global-pass schedule" as pass_point (level -> pass -> all blocks, neighbours from a full-size working array;
// SURVEY.md Appendix B), re-cut so that nothing per point depends on a 64-bit index decode:
//
//   * a CTA works inside ONE block-row (fixed block indices of all dims but the fastest, L) on a chunk of its lattice
//     rows; everything that depends on the block -- traversal base of the pass, extents, sub-phase sizes -- is put in
//     a small shared-memory table indexed by the block index along L (<= 256 entries, computed by the first threads
//     with the generic geometry functions of core.cuh), and everything that depends on the row in a row descriptor;
//   * a thread then walks lattice points along L: block index by shift (blocks are 32*s wide), table lookup, a few
//     integer operations, coalesced loads of the stencil taps (direction != L: whole warps read consecutive x), the
//     quantizer, one 16-bit index store at the closed-form traversal position.
//
// RECOVER = true is the decompression side (LinearQuantizer::recover): same walk, index read instead of written.
#pragma once
#include "core.cuh"
#include "interp_body.cuh"

namespace sz3b {

constexpr int kLeanThreads = 256;
constexpr int kLeanRows = 64;        // lattice rows of one block-row per CTA
constexpr int kLeanMaxBlocks = 256;  // table entries (blocks along the fastest dim): fastest extent up to 8192 at stride 1

struct LeanEntry {          // per block index along L
    uint64_t base;          // traversal position of this pass inside the block (block_base + earlier passes)
    uint32_t cnt, lo;       // L != D: owned lattice points along L and the first owned coordinate
    uint32_t other;         // product of the owned counts of all dims but D (size of one boundary sub-phase)
    uint32_t begin, n;      // L == D: block origin and points along L
    uint32_t main_cnt, nbnd;
    uint32_t bnd[3];
};

struct LeanRowDesc {        // per lattice row of the chunk
    uint64_t off;           // element offset of the row (coordinate 0 along L)
    uint32_t R;             // traversal rank of the row over the dims before L (this point's sub-phase)
    uint32_t R2;            // linear tail only: rank of the row of local index i-2 (main sub-phase)
    uint32_t i, n;          // D != L: local index along D and points along D
    uint32_t subk;          // D != L, boundary: main_cnt + index of the boundary sub-phase
    uint32_t in_main;       // D != L
    uint32_t valid;
};

struct LeanShared {
    LeanEntry tab[kLeanMaxBlocks];
    LeanRowDesc rows[kLeanRows];
    // block-row constants (thread 0)
    uint32_t bidx[kMaxDim];
    BlockGeom g;            // dims < L filled
    uint32_t nrows, row0;   // rows of this block-row, first row of this CTA's chunk
    uint32_t cntd[kMaxDim], lod[kMaxDim], stepd[kMaxDim];   // per dim < L, d != D
    uint32_t cD, mainD, nbndD, bndD[3], nD;                 // D < L
};

template <class T, class QT>
struct LeanArgs {
    InterpArgs<T, QT> A;     // shape, data, work, q, unpred_tmp, hist, qp, s, nb, block_base (+ batch strides)
    int p;                   // pass
    uint32_t chunks_per_brow;
    uint32_t nchunks_L;      // chunks along L per row (grid.y)
    uint32_t write_work;     // 0: nobody reads this pass's reconstructions (last pass of the finest level)
    uint32_t lg_s;           // log2 of the level stride (strides and blocks are powers of two: divisions become shifts)
    const T *unpred_in;      // RECOVER: position-indexed unpredictable values
};

// number of lattice targets along L in one row of pass p
template <class T, class QT>
SZ_HD uint32_t lean_row_len(const InterpArgs<T, QT> &A, int p) {
    const InterpShape &sh = A.sh;
    const int L = sh.N - 1;
    int qL = 0;
    for (int q = 0; q < sh.N; q++)
        if (sh.perm[q] == L) qL = q;
    if (qL == p) return ((sh.dims[L] - 1) / A.s + 1) / 2;
    return (sh.dims[L] - 1) / (qL < p ? A.s : 2 * A.s) + 1;
}

// rows of a FULL block-row (upper bound used to size the grid)
template <class T, class QT>
SZ_HD uint32_t lean_max_rows(const InterpArgs<T, QT> &A, int p) {
    const InterpShape &sh = A.sh;
    const int L = sh.N - 1, D = sh.perm[p];
    uint32_t rows = 1;
    for (int q = 0; q < sh.N; q++) {
        const int d = sh.perm[q];
        if (d == L) continue;
        // block bound (n <= 33: 16 odd local indices; 33 / 17 owned points at step s / 2s) and array bound
        const uint32_t npts = (sh.dims[d] - 1) / A.s + 1;
        if (d == D) rows *= npts / 2 < 16u ? npts / 2 : 16u;
        else if (q < p) rows *= npts < 33u ? npts : 33u;
        else rows *= (npts + 1) / 2 < 17u ? (npts + 1) / 2 : 17u;
    }
    return rows;
}

// thread 0 of the CTA: block-row constants
template <class T, class QT>
SZ_HD void lean_brow_setup(const LeanArgs<T, QT> &P, uint32_t brow, uint32_t chunk, LeanShared &S) {
    const InterpArgs<T, QT> &A = P.A;
    const InterpShape &sh = A.sh;
    const int N = sh.N, L = N - 1, D = sh.perm[P.p];
    const uint32_t s = A.s;
    uint32_t r = brow;
    for (int d = L - 1; d >= 0; d--) {
        S.bidx[d] = r % A.nb[d];
        r /= A.nb[d];
    }
    S.bidx[L] = 0;
    block_geom(sh, s, S.bidx, S.g);   // entry L is overwritten per table entry
    uint32_t nrows = 1;
    for (int q = 0; q < N; q++) {
        const int d = sh.perm[q];
        if (d == L) continue;
        if (d == D) {
            S.nD = S.g.n[d];
            S.cD = S.nD / 2;
            nrows *= S.nD <= 1 ? 0u : S.cD;
        } else {
            const uint32_t st = q < P.p ? s : 2 * s;
            S.stepd[d] = st;
            S.lod[d] = S.g.begin[d] ? S.g.begin[d] + st : 0;
            S.cntd[d] = q < P.p ? S.g.c1[d] : S.g.c2[d];
            nrows *= S.cntd[d];
        }
    }
    S.mainD = S.nbndD = 0;
    if (D != L) {
        // sub-phase structure along D (same rules as pass_geom)
        const uint32_t n = S.nD;
        if (n > 1) {
            if (sh.cubic) {
                S.mainD = n >= 7 ? (n - 7) / 2 + 1 : 0;
                S.bndD[S.nbndD++] = 1;
                if ((n & 1) && n > 3) S.bndD[S.nbndD++] = n - 2;
                if (!(n & 1) && n > 4) S.bndD[S.nbndD++] = n - 3;
                if (!(n & 1) && n > 2) S.bndD[S.nbndD++] = n - 1;
            } else {
                S.mainD = (n - 1) / 2;
                if (!(n & 1)) S.bndD[S.nbndD++] = n - 1;
            }
        }
    }
    S.nrows = nrows;
    S.row0 = chunk * kLeanRows;
}

// thread b (< nb[L]): table entry of block b along L
template <class T, class QT>
SZ_HD void lean_entry_setup(const LeanArgs<T, QT> &P, uint32_t brow, uint32_t b, const LeanShared &S, LeanEntry &E) {
    const InterpArgs<T, QT> &A = P.A;
    const InterpShape &sh = A.sh;
    const int L = sh.N - 1, D = sh.perm[P.p];
    uint32_t bidx[kMaxDim] = {S.bidx[0], S.bidx[1], S.bidx[2], S.bidx[3]};
    bidx[L] = b;
    BlockGeom g;
    block_geom(sh, A.s, bidx, g);
    uint64_t base = A.block_base[static_cast<uint64_t>(brow) * A.nb[L] + b];
    PassGeom pg;
    for (int pp = 0; pp < P.p; pp++) {
        pass_geom(sh, A.s, g, pp, pg);
        base += pg.size;
    }
    pass_geom(sh, A.s, g, P.p, pg);
    E.base = base;
    E.other = static_cast<uint32_t>(pg.other);
    E.begin = g.begin[L];
    E.n = g.n[L];
    E.main_cnt = pg.main_cnt;
    E.nbnd = pg.nbnd;
    for (int k = 0; k < 3; k++) E.bnd[k] = k < static_cast<int>(pg.nbnd) ? pg.bnd[k] : 0xffffffffu;
    if (D != L) {
        E.cnt = pg.cnt[L];
        E.lo = pg.lo[L];
    } else {
        E.cnt = 0;
        E.lo = 0;
    }
}

// lane r (< kLeanRows): descriptor of row row0 + r of the block-row
template <class T, class QT>
SZ_HD void lean_row_setup(const LeanArgs<T, QT> &P, uint32_t r, const LeanShared &S, LeanRowDesc &W) {
    const InterpArgs<T, QT> &A = P.A;
    const InterpShape &sh = A.sh;
    const int N = sh.N, L = N - 1, D = sh.perm[P.p];
    const uint32_t s = A.s;
    const uint32_t row = S.row0 + r;
    W.valid = row < S.nrows;
    if (!W.valid) return;
    // decode the row over dims < L in natural order (last of them fastest)
    uint32_t j[kMaxDim] = {0, 0, 0, 0};
    uint32_t rr = row;
    for (int d = L - 1; d >= 0; d--) {
        const uint32_t ext = d == D ? S.cD : S.cntd[d];
        j[d] = rr % ext;
        rr /= ext;
    }
    uint64_t off = 0;
    uint32_t i = 0, in_main = 1, idxD = 0, subk = 0;
    if (D != L) {
        i = 2 * j[D] + 1;
        const uint32_t n = S.nD;
        in_main = sh.cubic ? (i >= 3 && i + 3 < n) : (i + 1 < n);
        idxD = sh.cubic ? (i - 3) >> 1 : (i - 1) >> 1;
        if (!in_main) {
            uint32_t k = 0;
            while (k < S.nbndD && S.bndD[k] != i) k++;
            subk = S.mainD + k;
        }
    }
    uint32_t R = 0, R2 = 0;
    for (int d = 0; d < L; d++) {
        if (d == D) {
            off += static_cast<uint64_t>(S.g.begin[d] + i * s) * sh.stride[d];
            R = in_main ? R * S.mainD + idxD : R;          // boundary: extent 1, index 0
            R2 = R2 * S.mainD + ((i >= 2 ? i - 2 : 0) - 1) / 2;   // row of i-2 in the (linear) main sub-phase
        } else {
            off += static_cast<uint64_t>(S.lod[d] + j[d] * S.stepd[d]) * sh.stride[d];
            R = R * S.cntd[d] + j[d];
            R2 = R2 * S.cntd[d] + j[d];
        }
    }
    W.off = off;
    W.R = R;
    W.R2 = R2;
    W.i = i;
    W.n = S.nD;
    W.subk = subk;
    W.in_main = in_main;
}

// one lattice point of row W at lattice index `idx` along L
template <class T, class QT, class Ctx, bool RECOVER>
SZ_HD void lean_point(const LeanArgs<T, QT> &P, Ctx &ctx, const LeanShared &S, const LeanRowDesc &W, uint32_t idx,
                      uint32_t row_len, uint32_t batch, int qL) {
    const InterpArgs<T, QT> &A = P.A;
    const InterpShape &sh = A.sh;
    const int L = sh.N - 1, D = sh.perm[P.p];
    const uint32_t s = A.s;
    const uint32_t lgS = P.lg_s, lgB = P.lg_s + 5;   // blocks are 32*s wide
    bool active = W.valid && idx < row_len;
    int qv = 0;
    if (active) {
        const T *dat = A.data + batch * A.data_bstride;
        T *wk = A.work + batch * A.data_bstride;
        QT *qo = A.q + batch * A.q_bstride;
        uint32_t xL, i, n;
        uint64_t pos;
        bool in_main;
        int64_t sd;   // element stride of one local step along D
        uint64_t pos2 = 0;
        if (D != L) {
            const uint32_t stepL = qL < P.p ? s : 2 * s;
            xL = idx * stepL;
            const uint32_t b = xL ? (xL - 1) >> lgB : 0;
            const LeanEntry &E = S.tab[b];
            const uint32_t rL = (xL - E.lo) >> (qL < P.p ? lgS : lgS + 1);
            i = W.i;
            n = W.n;
            in_main = W.in_main != 0;
            pos = E.base + (in_main ? 0u : W.subk * E.other) + static_cast<uint64_t>(W.R) * E.cnt + rL;
            pos2 = E.base + static_cast<uint64_t>(W.R2) * E.cnt + rL;
            sd = static_cast<int64_t>(s) * static_cast<int64_t>(sh.stride[D]);
        } else {
            xL = (2 * idx + 1) * s;
            const uint32_t b = xL >> lgB;
            const LeanEntry &E = S.tab[b];
            i = (xL - E.begin) >> lgS;
            n = E.n;
            in_main = sh.cubic ? (i >= 3 && i + 3 < n) : (i + 1 < n);
            if (in_main) {
                const uint32_t idxD = sh.cubic ? (i - 3) >> 1 : (i - 1) >> 1;
                pos = E.base + static_cast<uint64_t>(W.R) * E.main_cnt + idxD;
            } else {
                const uint32_t k = i == E.bnd[0] ? 0u : (i == E.bnd[1] ? 1u : 2u);
                pos = E.base + static_cast<uint64_t>(E.main_cnt + k) * E.other + W.R;
            }
            pos2 = E.base + static_cast<uint64_t>(W.R) * E.main_cnt + (((i >= 2 ? i - 2 : 0) - 1) >> 1);
            sd = static_cast<int64_t>(s);
        }
        const int64_t off = static_cast<int64_t>(W.off + xL);
        auto v = [&](uint32_t k) -> T { return wk[off + (static_cast<int64_t>(k) - static_cast<int64_t>(i)) * sd]; };
        T r2 = 0;
        if (!sh.cubic && i + 1 == n && n >= 4) {
            // linear tail: the reconstruction of i-2 belongs to another thread of this same pass; rebuild it
            const T p2 = interp_linear<T>(v(i - 3), v(i - 1));
            if (RECOVER) {
                const int q2 = static_cast<int>(qo[pos2]);
                r2 = q2 ? recover_pred<T>(p2, q2, A.qp) : P.unpred_in[batch * A.q_bstride + pos2];
            } else {
                quantize<T>(dat[off - 2 * sd], p2, A.qp, r2);
            }
        }
        const T pred = predict_line<T>(sh.cubic, i, n, v, r2);
        if (RECOVER) {
            qv = static_cast<int>(qo[pos]);
            wk[off] = qv ? recover_pred<T>(pred, qv, A.qp) : P.unpred_in[batch * A.q_bstride + pos];
        } else {
            const T orig = dat[off];
            T rec;
            qv = quantize<T>(orig, pred, A.qp, rec);
            if (P.write_work) wk[off] = rec;
            qo[pos] = static_cast<QT>(qv);
            if (qv == 0) A.unpred_tmp[batch * A.q_bstride + pos] = orig;
        }
    }
    if (!RECOVER) ctx.hist_add(qv, active);
}

// whole CTA: (brow, chunk) from cta_x, chunk along L from cta_y
template <class T, class QT, class Ctx, bool RECOVER>
SZ_HD void lean_cta(const LeanArgs<T, QT> &P, Ctx &ctx, LeanShared &S, uint32_t cta_x, uint32_t cta_y, uint32_t batch) {
    const InterpArgs<T, QT> &A = P.A;
    const InterpShape &sh = A.sh;
    const int L = sh.N - 1;
    const uint32_t tid = ctx.tid(), nt = ctx.nthreads();
    const uint32_t brow = cta_x / P.chunks_per_brow, chunk = cta_x - brow * P.chunks_per_brow;
    int qL = 0;
    for (int q = 0; q < sh.N; q++)
        if (sh.perm[q] == L) qL = q;
    if (tid == 0) lean_brow_setup(P, brow, chunk, S);
    ctx.sync();
    if (S.row0 < S.nrows) {   // uniform: this chunk has rows
        for (uint32_t b = tid; b < A.nb[L]; b += nt) lean_entry_setup(P, brow, b, S, S.tab[b]);
        // the row descriptors are built by the LAST threads so that they overlap with the table entries
        for (uint32_t r = nt - 1 - tid; r < static_cast<uint32_t>(kLeanRows); r += nt) lean_row_setup(P, r, S, S.rows[r]);
    }
    ctx.sync();
    if (S.row0 >= S.nrows) return;
    const uint32_t row_len = lean_row_len(A, P.p);
    const uint32_t idx = cta_y * nt + tid;
    for (uint32_t r = 0; r < static_cast<uint32_t>(kLeanRows); r++)
        lean_point<T, QT, Ctx, RECOVER>(P, ctx, S, S.rows[r], idx, row_len, batch, qL);
}

}  // namespace sz3b
