// sz3_b200/csrc/interp_fast.cuh -- lean N == 3 tile schedule of the fused interpolation-predict + LinearQuantizer
// kernel (the hot loop of InterpolationDecomposition::compress, reference
// include/SZ3/decomposition/InterpolationDecomposition.hpp:79-147, :309-454).
//
// Same mathematics and the same closed-tile argument as tile_body (interp_body.cuh); what changes is the cost per
// point.  Everything that depends only on (tile, pass) is folded once per CTA into a table of small integers in
// shared memory (FastPass); a point then costs two multiply-high "divisions", a handful of integer multiply-adds,
// four shared-memory loads, the stencil, the quantizer and one 16-bit store at its reference traversal position.
//
//   * smem holds the sub-lattice that is even along the LAST pass dimension (<= 33*33*17 values); passes 0 and 1 work
//     in place, pass 2 streams its targets from global memory.
//   * low faces shared with the previous tile are recomputed only where a later pass of THIS tile reads them
//     (dimensions interpolated later); faces of dimensions already interpolated are skipped unless owned.
//   * every thread walks items in emission order, so the 16-bit index stores of a warp are contiguous.
#pragma once
#include "core.cuh"
#include "interp_body.cuh"

namespace sz3b {

struct FastPass {
    uint32_t c1, c2;          // item counts of natural dims 1 and 2 (dim 0 follows from total)
    uint32_t total;           // items of this pass
    uint32_t mg1, mg2;        // ceil(2^32 / c1), ceil(2^32 / c2)
    uint32_t mul[3], add[3];  // local index (units of s) l_d = i_d * mul_d + add_d
    uint32_t skip[3];         // dims interpolated LATER: items with i_d < skip_d lie on a low face this tile does not own
    uint32_t sP[3];           // smem offset = sum_d i_d * sP_d + sC  (address of the point itself in passes 0/1,
    uint32_t sC;              //   of the neighbour l_D - 1 in pass 2)
    uint32_t sD;              // smem stride of one neighbour step along the pass direction
    uint32_t D, n;            // direction (natural dim) and points along it
    uint32_t main_cnt, nbnd, bnd[3];
    uint32_t X1m, X2m, X1b, X2b;  // emission rank = (r0 * X1 + r1) * X2 + r2, main / boundary sub-phases
    uint32_t bbase[3];        // start of each boundary sub-phase inside the pass
    uint64_t base;            // absolute position of the pass in the index stream
};

struct FastTile {
    FastPass ps[3];
    uint32_t E[3];            // smem extents, natural order
    uint32_t mgE1, mgE2;
    uint32_t last;            // dimension of the last pass (halved in smem)
    uint64_t gbase, g2base;   // element offset of the tile origin in data / recon2
    uint64_t gS[3], g2S[3];   // element stride of one local-index step in data / recon2 (g2S valid for s >= 2)
};

SZ_HD uint32_t magic_u32(uint32_t c) {  // floor(x / c) == umulhi(x, magic) for x * c < 2^32 (x < 2^16 here)
    return c <= 1 ? 0u : static_cast<uint32_t>(((1ull << 32) + c - 1) / c);
}
SZ_HD uint32_t fast_div(uint32_t x, uint32_t mg) {
#if defined(__CUDA_ARCH__)
    return mg ? __umulhi(x, mg) : x;
#else
    return mg ? static_cast<uint32_t>((static_cast<uint64_t>(x) * mg) >> 32) : x;
#endif
}

template <class T, class QT>
SZ_HD void fast_tile_setup(const InterpArgs<T, QT> &A, uint32_t tile, uint32_t batch, FastTile &ft) {
    const InterpShape &sh = A.sh;
    const uint32_t s = A.s;
    TileGeom tg;
    tile_geom(A, tile, batch, tg);
    const int last = tg.a[2];
    ft.last = last;
    uint32_t sst[3];
    for (int d = 0; d < 3; d++) ft.E[d] = tg.E[d];
    sst[2] = 1;
    sst[1] = ft.E[2];
    sst[0] = ft.E[2] * ft.E[1];
    ft.mgE1 = magic_u32(ft.E[1]);
    ft.mgE2 = magic_u32(ft.E[2]);
    ft.gbase = batch * A.data_bstride;
    ft.g2base = batch * A.recon2_bstride;
    for (int d = 0; d < 3; d++) {
        ft.gbase += static_cast<uint64_t>(tg.g.begin[d]) * sh.stride[d];
        ft.g2base += static_cast<uint64_t>(tg.g.begin[d] >> 1) * A.stride2[d];
        ft.gS[d] = static_cast<uint64_t>(s) * sh.stride[d];
        ft.g2S[d] = static_cast<uint64_t>(s >> 1) * A.stride2[d];
    }
    for (int p = 0; p < 3; p++) {
        FastPass &fp = ft.ps[p];
        const PassGeom &pg = tg.pg[p];
        const int D = tg.a[p];
        fp.D = D;
        fp.n = pg.n;
        fp.base = tg.pass_base[p];
        fp.main_cnt = pg.main_cnt;
        fp.nbnd = pg.nbnd;
        uint32_t c[3];
        fp.sC = 0;
        for (int q = 0; q < 3; q++) {
            const int d = tg.a[q];
            const uint32_t n = tg.g.n[d];
            const uint32_t low = tg.g.begin[d] ? 1u : 0u;
            const uint32_t unit = d == last ? 0 : 1;   // smem index of local l: l (other dims) or l / 2 (last dim)
            if (q == p) {            // targets: odd local indices
                c[d] = n / 2;
                fp.mul[d] = 2;
                fp.add[d] = 1;
                fp.skip[d] = 0;
                if (p < 2) {         // point itself at 2 i + 1
                    fp.sP[d] = 2 * sst[d];
                    fp.sC += sst[d];
                    fp.sD = sst[d];
                } else {             // last pass: neighbour l - 1 = 2 i lives at smem index i
                    fp.sP[d] = sst[d];
                    fp.sD = sst[d];
                }
            } else if (q < p) {      // already refined to step s; low face only if owned
                c[d] = n - low;
                fp.mul[d] = 1;
                fp.add[d] = low;
                fp.skip[d] = 0;
                // q < p <= 2 means d != last
                fp.sP[d] = sst[d];
                fp.sC += low * sst[d];
            } else {                 // still on the 2s lattice; low face needed by the later pass along d
                c[d] = (n + 1) / 2;
                fp.mul[d] = 2;
                fp.add[d] = 0;
                fp.skip[d] = low;
                fp.sP[d] = unit ? 2 * sst[d] : sst[d];
            }
        }
        fp.c1 = c[1];
        fp.c2 = c[2];
        fp.total = pg.n <= 1 ? 0 : c[0] * c[1] * c[2];
        fp.mg1 = magic_u32(c[1]);
        fp.mg2 = magic_u32(c[2]);
        for (uint32_t k = 0; k < 3; k++) {
            fp.bnd[k] = k < pg.nbnd ? pg.bnd[k] : 0xffffffffu;
            fp.bbase[k] = static_cast<uint32_t>((pg.main_cnt + k) * pg.other);
        }
        // emission extents (owned points): direction -> main_cnt or 1, others -> pg.cnt
        const uint32_t e1m = D == 1 ? pg.main_cnt : pg.cnt[1], e2m = D == 2 ? pg.main_cnt : pg.cnt[2];
        const uint32_t e1b = D == 1 ? 1 : pg.cnt[1], e2b = D == 2 ? 1 : pg.cnt[2];
        fp.X1m = e1m;
        fp.X2m = e2m;
        fp.X1b = e1b;
        fp.X2b = e2b;
    }
}

// Per-pass constants pulled into registers once per thread (the table itself sits in shared memory, which the
// compiler must otherwise re-read after every shared-memory store of a reconstruction).
struct FastPassRegs {
    uint32_t c1, c2, mg1, mg2;
    uint32_t sP0, sP1, sP2, sC, sD;
    uint32_t D, n;
    uint32_t skip0, skip1, skip2;
    uint32_t X1m, X2m, X1b, X2b;
    uint32_t bnd0, bnd1, bb0, bb1, bb2;
    uint32_t l0m, l0a, l1m, l1a, l2m, l2a;
    uint64_t base;
};

SZ_HD void fast_pass_regs(const FastPass &fp, FastPassRegs &r) {
    r.c1 = fp.c1; r.c2 = fp.c2; r.mg1 = fp.mg1; r.mg2 = fp.mg2;
    r.sP0 = fp.sP[0]; r.sP1 = fp.sP[1]; r.sP2 = fp.sP[2]; r.sC = fp.sC; r.sD = fp.sD;
    r.D = fp.D; r.n = fp.n;
    r.skip0 = fp.skip[0]; r.skip1 = fp.skip[1]; r.skip2 = fp.skip[2];
    r.X1m = fp.X1m; r.X2m = fp.X2m; r.X1b = fp.X1b; r.X2b = fp.X2b;
    r.bnd0 = fp.bnd[0]; r.bnd1 = fp.bnd[1]; r.bb0 = fp.bbase[0]; r.bb1 = fp.bbase[1]; r.bb2 = fp.bbase[2];
    r.l0m = fp.mul[0]; r.l0a = fp.add[0]; r.l1m = fp.mul[1]; r.l1a = fp.add[1]; r.l2m = fp.mul[2]; r.l2a = fp.add[2];
    r.base = fp.base;
}

// One item of pass P (0/1: in-place in smem, 2: targets streamed from global memory).
template <class T, class QT, class Ctx, bool LAST>
SZ_HD void fast_item(const InterpArgs<T, QT> &A, Ctx &ctx, const FastTile &ft, const FastPassRegs &fp, T *sm,
                     uint32_t it, bool active, bool cubic, bool write2) {
    int qv = 0;
    uint64_t pos = 0;
    T orig = 0;
    bool owned = false;
    if (active) {
        const uint32_t r = fast_div(it, fp.mg2);
        const uint32_t i2 = it - r * fp.c2;
        const uint32_t i0 = fast_div(r, fp.mg1);
        const uint32_t i1 = r - i0 * fp.c1;
        const uint32_t D = fp.D, n = fp.n;
        const uint32_t iD = D == 0 ? i0 : (D == 1 ? i1 : i2);
        const uint32_t i = 2 * iD + 1;
        const bool tail = !cubic && i + 1 == n && n >= 4;   // linear i == n-1: handled by the i == n-3 item
        if (!tail) {
            const uint32_t soff = i0 * fp.sP0 + i1 * fp.sP1 + i2 * fp.sP2 + fp.sC;
            const uint32_t sD = fp.sD;
            // neighbours at local l-3, l-1, l+1, l+3 along D
            const T *nb = LAST ? sm + soff : sm + soff - sD;   // -> l-1
            const uint32_t st = LAST ? sD : 2 * sD;
            T pred;
            if (cubic) {
                if (i >= 3) {
                    if (i + 3 < n) pred = interp_cubic<T>(nb[-static_cast<int>(st)], nb[0], nb[st], nb[2 * st]);
                    else if (i + 1 < n) pred = interp_quad_2<T>(nb[-static_cast<int>(st)], nb[0], nb[st]);
                    else pred = interp_linear1<T>(nb[-static_cast<int>(st)], nb[0]);
                } else {
                    if (i + 3 < n) pred = interp_quad_1<T>(nb[0], nb[st], nb[2 * st]);
                    else if (i + 1 < n) pred = interp_linear<T>(nb[0], nb[st]);
                    else pred = nb[0];
                }
            } else {
                if (i + 1 < n) pred = interp_linear<T>(nb[0], nb[st]);
                else pred = nb[0];   // n < 3 (n >= 4 is the tail, done below by its predecessor)
            }
            // global offset of the point (needed for the original value in the last pass and for recon2)
            uint64_t goff = 0, g2off = 0;
            const uint32_t l0 = i0 * fp.l0m + fp.l0a, l1 = i1 * fp.l1m + fp.l1a, l2 = i2 * fp.l2m + fp.l2a;
            if (LAST) goff = ft.gbase + l0 * ft.gS[0] + l1 * ft.gS[1] + l2 * ft.gS[2];
            if (write2) g2off = ft.g2base + l0 * ft.g2S[0] + l1 * ft.g2S[1] + l2 * ft.g2S[2];
            orig = LAST ? A.data[goff] : sm[soff];
            T rec;
            qv = quantize<T>(orig, pred, A.qp, rec);
            if (!LAST) sm[soff] = rec;
            // emission position
            owned = i0 >= fp.skip0 && i1 >= fp.skip1 && i2 >= fp.skip2;
            const bool in_main = cubic ? (i >= 3 && i + 3 < n) : (i + 1 < n);
            const uint32_t rD = in_main ? (cubic ? iD - 1 : iD) : 0;
            const uint32_t r0 = D == 0 ? rD : i0 - fp.skip0;
            const uint32_t r1 = D == 1 ? rD : i1 - fp.skip1;
            const uint32_t r2 = D == 2 ? rD : i2 - fp.skip2;
            const uint32_t sub = in_main ? 0 : (i == fp.bnd0 ? fp.bb0 : (i == fp.bnd1 ? fp.bb1 : fp.bb2));
            const uint32_t X1 = in_main ? fp.X1m : fp.X1b, X2 = in_main ? fp.X2m : fp.X2b;
            pos = fp.base + (sub + (r0 * X1 + r1) * X2 + r2);
            if (owned && write2) A.recon2[g2off] = rec;
            if (!cubic && i + 3 == n && n >= 4 && !(n & 1)) {
                // flush this point, then the linear tail i+2 = n-1: linear1(recon(i), value(i+1))
                emit(A, ctx, pos, qv, orig, owned);
                T pred2 = interp_linear1<T>(rec, nb[st]);
                const uint32_t t_soff = soff + 2 * sD;   // passes 0/1 only (smem address of i+2)
                const uint64_t gSD = D == 0 ? ft.gS[0] : (D == 1 ? ft.gS[1] : ft.gS[2]);
                const uint64_t g2SD = D == 0 ? ft.g2S[0] : (D == 1 ? ft.g2S[1] : ft.g2S[2]);
                orig = LAST ? A.data[goff + 2 * gSD] : sm[t_soff];
                qv = quantize<T>(orig, pred2, A.qp, rec);
                if (!LAST) sm[t_soff] = rec;
                // i+2 = n-1 is the (single) boundary sub-phase of linear mode
                const uint32_t t0 = D == 0 ? 0 : r0, t1 = D == 1 ? 0 : r1, t2 = D == 2 ? 0 : r2;
                pos = fp.base + (fp.bb0 + (t0 * fp.X1b + t1) * fp.X2b + t2);
                if (owned && write2) A.recon2[g2off + 2 * g2SD] = rec;
            }
        }
    }
    emit(A, ctx, pos, qv, orig, active && owned);
}

template <class T, class QT, class Ctx>
SZ_HD void fast_tile_body(const InterpArgs<T, QT> &A, Ctx &ctx, T *sm, const FastTile &ft) {
    const uint32_t tid = ctx.tid(), nt = ctx.nthreads();
    const bool cubic = A.sh.cubic != 0;
    const bool write2 = A.s >= 2;
    // ---- load the sub-lattice even along the last pass dimension: coarse points (all local indices even) from
    //      recon2, everything else from the immutable input ------------------------------------------------------
    {
        const uint32_t total = ft.E[0] * ft.E[1] * ft.E[2];
        const uint32_t last = ft.last;
        for (uint32_t it = tid; it < total; it += nt) {
            const uint32_t r = fast_div(it, ft.mgE2);
            const uint32_t e2 = it - r * ft.E[2];
            const uint32_t e0 = fast_div(r, ft.mgE1);
            const uint32_t e1 = r - e0 * ft.E[1];
            const uint32_t l0 = last == 0 ? 2 * e0 : e0, l1 = last == 1 ? 2 * e1 : e1, l2 = last == 2 ? 2 * e2 : e2;
            const bool coarse = !((l0 | l1 | l2) & 1);
            T v;
            if (coarse) {
                // recon2 index of local l (even) is origin2 + l/2 * (s * stride2 / ... ): (begin + l*s)/2
                const uint64_t o = ft.g2base + (l0 >> 1) * (A.s * A.stride2[0]) + (l1 >> 1) * (A.s * A.stride2[1]) +
                                   (l2 >> 1) * (A.s * A.stride2[2]);
                v = A.recon2[o];
            } else {
                v = A.data[ft.gbase + l0 * ft.gS[0] + l1 * ft.gS[1] + l2 * ft.gS[2]];
            }
            sm[it] = v;
        }
    }
    ctx.sync();
    for (int p = 0; p < 3; p++) {
        FastPassRegs fp;
        fast_pass_regs(ft.ps[p], fp);
        const uint32_t total = ft.ps[p].total;
        const uint32_t rounds = (total + nt - 1) / nt;
        for (uint32_t rd = 0; rd < rounds; rd++) {
            const uint32_t it = rd * nt + tid;
            if (p < 2)
                fast_item<T, QT, Ctx, false>(A, ctx, ft, fp, sm, it, it < total, cubic, write2);
            else
                fast_item<T, QT, Ctx, true>(A, ctx, ft, fp, sm, it, it < total, cubic, write2);
        }
        if (p < 2) ctx.sync();
    }
}

}  // namespace sz3b
