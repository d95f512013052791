// sz3_b200/csrc/interp_body.cuh -- bodies of the fused interpolation-predict + LinearQuantizer kernels.
//
// Replaces InterpolationDecomposition::compress (reference include/SZ3/decomposition/InterpolationDecomposition.hpp
// :79-147, :215-222, :309-454).  Two schedules, both bit-identical to the reference traversal:
//
//  * tile_body   (N == 3): one CTA per closed level block (<= 33^3 points at spacing s).  The sub-lattice that is even
//                along the LAST pass dimension lives in shared memory; coarse points (all local indices even) come from
//                the compact reconstruction array `recon2`, everything else from the immutable input.  Three passes
//                separated by CTA barriers; the last pass streams its targets straight from global memory.  Valid
//                because a closed tile recomputed from {original} U {final coarse points} reproduces the reference
//                values bit for bit ("closed-tile self-containment").
//  * pass_point  (any N, old and new API): one thread per predicted point of one (level, pass), neighbours read from a
//                full-size working array.  Valid because level -> pass -> all blocks is equivalent to the reference's
//                level -> block -> passes order ("global-pass schedule").
//
// Both write each quantization index at its reference traversal position (closed form, core.cuh), so the index
// stream is identical to the reference's std::vector<int> quant_inds.
//
// The bodies are templates over a context type so tests/emul can run them with host threads; the product only ever
// instantiates them inside __global__ kernels (interp_kernels.cu).
#pragma once
#include "core.cuh"

namespace sz3b {

template <class T, class QT>
struct InterpArgs {
    InterpShape sh;
    const T *data;          // immutable input; batch element b starts at data + b * data_bstride
    T *work;                // pass_point: full-size reconstruction array (same layout as data)
    T *recon2;              // tile_body: reconstruction at even coordinates, dims2 = (dims-1)/2+1, row-major
    uint64_t data_bstride;
    uint64_t recon2_bstride;
    uint64_t q_bstride;     // == number of elements of one array
    uint32_t dims2[kMaxDim];
    uint64_t stride2[kMaxDim];
    QT *q;                  // quantization indices, reference traversal order
    T *unpred_tmp;          // sparse: unpred_tmp[pos] = original value where q[pos] == 0
    unsigned long long *hist;  // global histogram, 2*radius bins
    QuantParams qp;         // quantizer of this level
    uint32_t s;             // level stride
    uint32_t nb[kMaxDim];   // blocks per dimension at this level
    const uint64_t *block_base;  // position (inside one array) of the first index each block emits
    uint32_t tile0;         // tile kernels: index of the first tile of this launch (partial launches of one level)
};

template <class T, class QT, class Ctx>
SZ_HD void emit(const InterpArgs<T, QT> &A, Ctx &ctx, uint64_t pos, int qv, T orig, bool active) {
    if (active) {
        A.q[pos] = static_cast<QT>(qv);
        if (qv == 0) A.unpred_tmp[pos] = orig;
    }
    ctx.hist_add(qv, active);
}

// ---------------------------------------------------------------------------------------------------------------------
// Generic schedule: thread `gid` of pass `p` at level stride A.s, array `batch`.
// ---------------------------------------------------------------------------------------------------------------------
template <class T, class QT>
SZ_HD uint64_t pass_points(const InterpArgs<T, QT> &A, int p) {
    const InterpShape &sh = A.sh;
    const uint32_t s = A.s;
    uint64_t total = 1;
    for (int q = 0; q < sh.N; q++) {
        int d = sh.perm[q];
        uint32_t ext;
        if (q == p)
            ext = ((sh.dims[d] - 1) / s + 1) / 2;
        else
            ext = (sh.dims[d] - 1) / (q < p ? s : 2 * s) + 1;
        total *= ext;
    }
    return total;
}

template <class T, class QT, class Ctx>
SZ_HD void pass_point(const InterpArgs<T, QT> &A, Ctx &ctx, int p, uint64_t gid, uint64_t total, uint32_t batch) {
    const InterpShape &sh = A.sh;
    const uint32_t s = A.s;
    const int D = sh.perm[p];
    bool active = gid < total;
    int qv = 0;
    uint64_t pos = 0;
    T orig = 0;
    if (active) {
        uint32_t step[kMaxDim], ext[kMaxDim], x[kMaxDim], bidx[kMaxDim];
        for (int q = 0; q < sh.N; q++) {
            int d = sh.perm[q];
            step[d] = q < p ? s : 2 * s;
            ext[d] = q == p ? ((sh.dims[d] - 1) / s + 1) / 2 : (sh.dims[d] - 1) / step[d] + 1;
        }
        const uint32_t B = kInterpBlock * s;
        uint64_t off = 0, blin = 0;
        if (total <= 0xffffffffull) {   // 32-bit index arithmetic (a 64-bit division costs ~10x more on the GPU)
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
                uint32_t idx = static_cast<uint32_t>(r % ext[d]);
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
        const T *dat = A.data + batch * A.data_bstride;
        T *wk = A.work + batch * A.data_bstride;
        const int64_t sd = static_cast<int64_t>(s) * static_cast<int64_t>(sh.stride[D]);
        auto v = [&](uint32_t k) -> T {
            return wk[static_cast<int64_t>(off) + (static_cast<int64_t>(k) - static_cast<int64_t>(i)) * sd];
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
                // linear tail: needs the reconstruction of i-2, predicted in this same pass by another thread;
                // recompute it here from immutable inputs instead of racing on the working array.
                T p2 = interp_linear<T>(v(i - 3), v(i - 1));
                quantize<T>(dat[static_cast<int64_t>(off) - 2 * sd], p2, A.qp, r2);
            }
            pred = predict_line<T>(sh.cubic, i, n, v, r2);
            in_pass = pass_offset(sh, pg, x, i);
        }
        orig = dat[off];
        T rec;
        qv = quantize<T>(orig, pred, A.qp, rec);
        wk[off] = rec;
        pos = batch * A.q_bstride + base + in_pass;
    }
    emit(A, ctx, pos, qv, orig, active);
}

// ---------------------------------------------------------------------------------------------------------------------
// Tile schedule (N == 3).  smem holds at most 33*33*17 elements of T.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kTileSmemElems = 33 * 33 * 17;

struct TileGeom {
    BlockGeom g;
    PassGeom pg[3];
    uint64_t pass_base[3];   // absolute position of each pass (batch offset included)
    uint32_t E[3];           // smem extents, natural order (dimension a2 halved)
    uint32_t sst[3];         // smem strides, natural order
    int a[3];                // pass order: a[p] = sh.perm[p]
};

template <class T, class QT>
SZ_HD void tile_geom(const InterpArgs<T, QT> &A, uint32_t tile, uint32_t batch, TileGeom &tg) {
    const InterpShape &sh = A.sh;
    uint32_t bidx[kMaxDim];
    uint32_t r = tile;
    for (int d = 2; d >= 0; d--) {
        bidx[d] = r % A.nb[d];
        r /= A.nb[d];
    }
    block_geom(sh, A.s, bidx, tg.g);
    uint64_t base = batch * A.q_bstride + A.block_base[tile];
    for (int p = 0; p < 3; p++) {
        tg.a[p] = sh.perm[p];
        pass_geom(sh, A.s, tg.g, p, tg.pg[p]);
        tg.pass_base[p] = base;
        base += tg.pg[p].size;
    }
    for (int d = 0; d < 3; d++) tg.E[d] = d == tg.a[2] ? (tg.g.n[d] + 1) / 2 : tg.g.n[d];
    tg.sst[2] = 1;
    tg.sst[1] = tg.E[2];
    tg.sst[0] = tg.E[2] * tg.E[1];
}

}  // namespace sz3b
