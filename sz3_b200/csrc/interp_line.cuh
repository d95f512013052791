// sz3_b200/csrc/interp_line.cuh -- line-walker N == 3 tile schedule (fallback of the box schedule) of the fused interpolation-predict +
// LinearQuantizer kernel (InterpolationDecomposition::compress, reference
// include/SZ3/decomposition/InterpolationDecomposition.hpp:79-147, :309-454).
//
// Same closed-tile mathematics as tile_body (interp_body.cuh); the work is re-cut so that a thread never decodes an
// index per point.  A pass along direction D is executed as
//
//   line items (D != fastest dim): the thread owns a segment of one line along D; lanes of a warp sit on consecutive
//       x, so shared-memory reads, the 16-bit index stores and the recon2 stores are contiguous; the four stencil
//       taps slide through registers (one new shared-memory load per point), every offset advances by a constant.
//   row items  (D == fastest dim): 16 lanes cover the 16 targets of one row along x, the thread walks over rows
//       (y, z) with constant increments; its stencil type (cubic / quad_1 / quad_2 / linear / copy) is fixed, so it
//       is evaluated branch-free as a 4-tap filter whose coefficients reproduce the reference's evaluation order.
//
// The quantization-index histogram of HuffmanEncoder::init is accumulated in packed per-thread registers for the
// 8 bins around the radius (Ctx::hist_add) and reduced once per pass.
#pragma once
#include "core.cuh"
#include "interp_body.cuh"

namespace sz3b {

// floor(x / c) == umulhi(x, magic(c)) for x * c < 2^32 (magic 0 stands for c <= 1)
SZ_HD uint32_t fast_div(uint32_t x, uint32_t mg) {
#if defined(__CUDA_ARCH__)
    return mg ? __umulhi(x, mg) : x;
#else
    return mg ? static_cast<uint32_t>((static_cast<uint64_t>(x) * mg) >> 32) : x;
#endif
}


struct LinePass {
    uint32_t kind;             // 0 = line items, 1 = row items
    uint32_t D, n, cD, jmax;   // direction, points along it, targets (n/2), largest even local index / 2
    uint32_t cu, cv;           // items along the two other dims (u slower than v in natural order)
    uint32_t mgu, mgv;
    uint32_t skipu, skipv;     // items below skip lie on a low face owned by the previous tile
    uint32_t sU, sV, sC;       // smem offset of (iu, iv, local index 0 along D)
    uint32_t hS, tS;           // neighbour j (local 2j) at + j*hS; target k (local 2k+1) at + k*hS + tS (passes 0/1)
    uint32_t mU, mV, mD;       // emission rank multipliers, main sub-phase
    uint32_t bU, bV;           // ... boundary sub-phases
    uint32_t kfirst;           // k of the first main target (cubic 1, linear 0)
    uint32_t bnd[3], bbase[3]; // boundary sub-phases: local index and start offset inside the pass
    uint32_t nseg, seglen;     // line items: segments per line
    uint64_t base;             // position of the pass in the index stream
    uint64_t gU, gV, gD, gC;   // data element offset (from the tile origin) = iu*gU + iv*gV + gC + l_D*gD
    uint64_t hU, hV, hD, hC;   // same for recon2 (level stride >= 2)
};

struct LineTile {
    LinePass ps[3];
};

// Divisors inside a tile are its extents (<= 33): their magic numbers come from a table in constant memory.  Every
// thread needs four of them before its first load of the fill; as integer divisions they were 1 % of the instructions
// and 13 % of the stall samples of the level-1 launch (profiles/r1y_full_interp_ltile.md).
struct MagicTab {
    uint32_t v[65];
};
constexpr MagicTab make_magic_tab() {
    MagicTab t{};
    for (uint32_t c = 0; c <= 64; c++) t.v[c] = c <= 1 ? 0u : 0xffffffffu / c + 1u;
    return t;
}
#ifdef __CUDACC__
static __constant__ MagicTab kMagicTab = make_magic_tab();
#endif
SZ_HD uint32_t magic_u32_fast(uint32_t c) {
#ifdef __CUDA_ARCH__
    if (c <= 64u) return kMagicTab.v[c];
#endif
    return c <= 1 ? 0u : 0xffffffffu / c + 1u;
}

// rare path (an unpredictable point): kept out of line so that the hot loops carry a branch, not predicated stores
template <class T>
SZ_NOINLINE void store_unpred(T *ub, uint32_t pos, T orig) {
    ub[pos] = orig;
}

// Tile geometry every thread keeps in registers (fill phase).
struct LineGeom {
    uint32_t begin[3], n[3], E[3];
    uint32_t last;
    uint64_t gbase, g2base;
};

template <class T, class QT>
SZ_HD void line_geom(const InterpArgs<T, QT> &A, uint32_t tile, uint32_t batch, LineGeom &lg) {
    const InterpShape &sh = A.sh;
    const uint32_t s = A.s;
    uint32_t bidx[3];
    uint32_t r = tile;
    bidx[2] = r % A.nb[2];
    r /= A.nb[2];
    bidx[1] = r % A.nb[1];
    bidx[0] = r / A.nb[1];
    lg.last = static_cast<uint32_t>(sh.perm[2]);
    lg.gbase = batch * A.data_bstride;
    lg.g2base = batch * A.recon2_bstride;
    const uint32_t B = kInterpBlock * s;
    for (int d = 0; d < 3; d++) {
        const uint32_t b = bidx[d] * B;
        uint32_t e = b + B;
        if (e > sh.dims[d] - 1) e = sh.dims[d] - 1;
        lg.begin[d] = b;
        lg.n[d] = (e - b) / s + 1;
        lg.E[d] = static_cast<uint32_t>(d) == lg.last ? (lg.n[d] + 1) / 2 : lg.n[d];
        lg.gbase += static_cast<uint64_t>(b) * sh.stride[d];
        lg.g2base += static_cast<uint64_t>(b >> 1) * A.stride2[d];
    }
}

// Pass table p of one tile (thread p of the CTA runs this while the others fill shared memory).
template <class T, class QT>
SZ_HD void line_pass_setup(const InterpArgs<T, QT> &A, uint32_t tile, uint32_t batch, int p, uint32_t nthreads,
                           LinePass &P) {
    const InterpShape &sh = A.sh;
    const uint32_t s = A.s;
    TileGeom tg;
    tile_geom(A, tile, batch, tg);
    const int last = tg.a[2];
    const int D = tg.a[p];
    const PassGeom &pg = tg.pg[p];
    P.D = static_cast<uint32_t>(D);
    P.n = pg.n;
    P.cD = pg.n / 2;
    P.jmax = (pg.n - 1) / 2;
    P.kind = D == 2 ? 1u : 0u;
    P.base = tg.pass_base[p];
    P.kfirst = sh.cubic ? 1u : 0u;
    for (uint32_t k = 0; k < 3; k++) {
        P.bnd[k] = k < pg.nbnd ? pg.bnd[k] : 0xffffffffu;
        P.bbase[k] = static_cast<uint32_t>((pg.main_cnt + k) * pg.other);
    }
    // the two other dims in natural order
    int u = -1, v = -1;
    for (int d = 0; d < 3; d++)
        if (d != D) {
            if (u < 0) u = d; else v = d;
        }
    uint32_t cnt[3] = {0, 0, 0}, mul[3] = {0, 0, 0}, add[3] = {0, 0, 0}, skip[3] = {0, 0, 0};
    for (int q = 0; q < 3; q++) {
        const int d = tg.a[q];
        const uint32_t n = tg.g.n[d];
        const uint32_t low = tg.g.begin[d] ? 1u : 0u;
        if (q == p) continue;
        if (q < p) {          // refined to step s; the low face is neither owned nor read later
            cnt[d] = n - low; mul[d] = 1; add[d] = low; skip[d] = 0;
        } else {              // still on the 2s lattice; low face feeds the later pass along d
            cnt[d] = (n + 1) / 2; mul[d] = 2; add[d] = 0; skip[d] = low;
        }
    }
    // smem: local index l of dim d sits at l (d != last) or l/2 (d == last, l even)
    auto sm_step = [&](int d, uint32_t lstep) -> uint32_t {
        return d == last ? (lstep / 2) * tg.sst[d] : lstep * tg.sst[d];
    };
    P.cu = cnt[u]; P.cv = cnt[v];
    P.mgu = magic_u32_fast(P.cu); P.mgv = magic_u32_fast(P.cv);
    P.skipu = skip[u]; P.skipv = skip[v];
    // (add is 0 or 1 and only 1 when the dim is not `last`... a dim with add == 1 was interpolated earlier, the
    //  last pass dim never is)
    P.sU = sm_step(u, mul[u]); P.sV = sm_step(v, mul[v]);
    P.sC = add[u] * tg.sst[u] + add[v] * tg.sst[v];
    if (p < 2) {
        P.hS = 2 * tg.sst[D];
        P.tS = tg.sst[D];
    } else {
        P.hS = tg.sst[D];
        P.tS = 0;
    }
    // emission multipliers: rank = (r0*X1 + r1)*X2 + r2 over natural dims, r_D = main index (or 0 on a boundary)
    {
        const uint32_t em[3] = {D == 0 ? pg.main_cnt : pg.cnt[0], D == 1 ? pg.main_cnt : pg.cnt[1],
                                D == 2 ? pg.main_cnt : pg.cnt[2]};
        const uint32_t eb[3] = {D == 0 ? 1u : pg.cnt[0], D == 1 ? 1u : pg.cnt[1], D == 2 ? 1u : pg.cnt[2]};
        const uint32_t mm[3] = {em[1] * em[2], em[2], 1u};
        const uint32_t bm[3] = {eb[1] * eb[2], eb[2], 1u};
        P.mU = mm[u]; P.mV = mm[v]; P.mD = mm[D];
        P.bU = bm[u]; P.bV = bm[v];
    }
    // global offsets relative to the tile origin
    const uint64_t gs[3] = {static_cast<uint64_t>(s) * sh.stride[0], static_cast<uint64_t>(s) * sh.stride[1],
                            static_cast<uint64_t>(s) * sh.stride[2]};
    const uint64_t hs[3] = {static_cast<uint64_t>(s >> 1) * A.stride2[0], static_cast<uint64_t>(s >> 1) * A.stride2[1],
                            static_cast<uint64_t>(s >> 1) * A.stride2[2]};
    P.gU = mul[u] * gs[u]; P.gV = mul[v] * gs[v]; P.gD = gs[D];
    P.gC = add[u] * gs[u] + add[v] * gs[v];
    P.hU = mul[u] * hs[u]; P.hV = mul[v] * hs[v]; P.hD = hs[D];
    P.hC = add[u] * hs[u] + add[v] * hs[v];
    // segments of a line: keep every round of the CTA reasonably full, never shorter than 2 targets
    P.nseg = 1;
    P.seglen = P.cD;
    if (P.kind == 0 && P.cD >= 4) {
        const uint32_t lines = P.cu * P.cv;
        uint32_t best = 1;
        uint64_t best_cost = ~0ull;
        for (uint32_t ns = 1; ns <= 4 && P.cD / ns >= 2; ns *= 2) {
            const uint32_t items = lines * ns;
            const uint32_t rounds = (items + nthreads - 1) / nthreads;
            const uint32_t len = P.cD / ns + P.cD % ns;
            const uint64_t cost = static_cast<uint64_t>(rounds) * (len + 2);   // +2: window priming per segment
            if (cost < best_cost) {
                best_cost = cost;
                best = ns;
            }
        }
        P.nseg = best;
        P.seglen = P.cD / best;
    }
    if (pg.n <= 1) P.cu = 0;   // nothing to do along a degenerate direction
}

// ---------------------------------------------------------------------------------------------------------------------
// line items: the thread walks targets k0 <= k < k1 of one line along D
// ---------------------------------------------------------------------------------------------------------------------
template <class T, class QT, class Ctx, bool LAST, bool WRITE2>
SZ_HD void line_items(const InterpArgs<T, QT> &A, Ctx &ctx, const LineGeom &lg, const LinePass &P, T *sm, bool cubic) {
    const uint32_t tid = ctx.tid(), nt = ctx.nthreads();
    // pass constants -> registers (the table lives in shared memory)
    const uint32_t cu = P.cu, cv = P.cv, mgu = P.mgu, mgv = P.mgv, n = P.n, cD = P.cD, jmax = P.jmax;
    const uint32_t hS = P.hS, tS = P.tS, mD = P.mD, nseg = P.nseg, seglen = P.seglen;
    const uint32_t bnd0 = P.bnd[0], bnd1 = P.bnd[1], bb0 = P.bbase[0], bb1 = P.bbase[1], bb2 = P.bbase[2];
    const uint32_t kfirst = P.kfirst, skipu = P.skipu, skipv = P.skipv;
    const uint32_t gstep = static_cast<uint32_t>(2 * P.gD), hstep = static_cast<uint32_t>(2 * P.hD);
    QT *const qb = A.q + P.base;
    T *const ub = A.unpred_tmp + P.base;
    const T *const gb = A.data + lg.gbase + P.gC + P.gD;
    T *const hb = A.recon2 + lg.g2base + P.hC + P.hD;
    const QuantParams qp = A.qp;
    const uint32_t nitems = nseg * cu * cv;
    for (uint32_t it = tid; it < nitems; it += nt) {
        const uint32_t t = fast_div(it, mgv);
        const uint32_t iv = it - t * cv;
        const uint32_t sg = fast_div(t, mgu);
        const uint32_t iu = t - sg * cu;
        const uint32_t k0 = sg * seglen;
        const uint32_t k1 = sg + 1 == nseg ? cD : k0 + seglen;
        const bool owned = iu >= skipu && iv >= skipv;
        const uint32_t ru = iu - skipu, rv = iv - skipv;
        uint32_t pm = ru * P.mU + rv * P.mV + (k0 - kfirst) * mD;
        const uint32_t pb = ru * P.bU + rv * P.bV;
        uint32_t goff = iu * static_cast<uint32_t>(P.gU) + iv * static_cast<uint32_t>(P.gV) + k0 * gstep;
        uint32_t hoff = iu * static_cast<uint32_t>(P.hU) + iv * static_cast<uint32_t>(P.hV) + k0 * hstep;
        T *nb = sm + (iu * P.sU + iv * P.sV + P.sC + k0 * hS);   // neighbour j = k (local index i-1)
        T w0 = 0, w1 = nb[0], w2 = 0, w3 = 0;
        if (k0 >= 1) w0 = nb[-static_cast<int>(hS)];
        if (k0 + 1 <= jmax) w2 = nb[hS];
        if (k0 + 2 <= jmax) w3 = nb[2 * hS];
        T prev_rec = 0;
        T nxt = LAST ? gb[goff] : static_cast<T>(0);
        for (uint32_t k = k0; k < k1; k++) {
            const uint32_t i = 2 * k + 1;
            const T orig = LAST ? nxt : nb[tS];
            if (LAST && k + 1 < k1) nxt = gb[goff + gstep];
            T pred;
            bool in_main;
            if (cubic) {
                if (k >= 1) {
                    if (i + 3 < n) {
                        pred = interp_cubic<T>(w0, w1, w2, w3);
                        in_main = true;
                    } else {
                        in_main = false;
                        pred = i + 1 < n ? interp_quad_2<T>(w0, w1, w2) : interp_linear1<T>(w0, w1);
                    }
                } else {
                    in_main = false;
                    pred = i + 3 < n ? interp_quad_1<T>(w1, w2, w3) : (i + 1 < n ? interp_linear<T>(w1, w2) : w1);
                }
            } else {
                if (i + 1 < n) {
                    pred = interp_linear<T>(w1, w2);
                    in_main = true;
                } else {
                    in_main = false;
                    pred = n < 3 ? w1 : interp_linear1<T>(prev_rec, w1);
                }
            }
            T rec;
            const int qv = quantize<T>(orig, pred, qp, rec);
            if (!LAST) nb[tS] = rec;
            const uint32_t pos = in_main ? pm : pb + (i == bnd0 ? bb0 : (i == bnd1 ? bb1 : bb2));
            if (owned) {
                qb[pos] = static_cast<QT>(qv);
                if (qv == 0) store_unpred(ub, pos, orig);
                if (WRITE2) hb[hoff] = rec;
            }
            ctx.hist_add(qv, owned);
            prev_rec = rec;
            pm += mD;
            w0 = w1;
            w1 = w2;
            w2 = w3;
            nb += hS;
            w3 = k + 3 <= jmax ? nb[2 * hS] : static_cast<T>(0);
            goff += gstep;
            hoff += hstep;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// row items (D == 2): 16 lanes on the targets of one row, the thread walks over rows.
// SIMPLE: n odd (no double-precision linear1 lane, no merged linear tail) and the rows a thread visits keep the same
// iv (the stride over rows is a multiple of cv): everything advances by constants, no wrap test.
// ---------------------------------------------------------------------------------------------------------------------
template <class T, class QT, class Ctx, bool LAST, bool WRITE2, bool SIMPLE>
SZ_HD void row_items(const InterpArgs<T, QT> &A, Ctx &ctx, const LineGeom &lg, const LinePass &P, T *sm, bool cubic) {
    const uint32_t tid = ctx.tid(), nt = ctx.nthreads();
    const uint32_t cu = P.cu, cv = P.cv, n = P.n, cD = P.cD;
    constexpr uint32_t hS = LAST ? 1u : 2u;   // D == 2: sst[D] == 1
    constexpr uint32_t tS = LAST ? 0u : 1u;
    const uint32_t k = tid & 15u, slot = tid >> 4, nslots = nt >> 4;
    const uint32_t i = 2 * k + 1;
    const bool merge_tail = !SIMPLE && !cubic && !(n & 1u) && n >= 4;   // linear: target n-1 is done by the lane of n-3
    const bool act = k < cD && !(merge_tail && k + 1 == cD);
    // stencil of this lane as a 4-tap filter over (l-3, l-1, l+1, l+3); unused taps have coefficient 0 and are not
    // loaded (an infinite neighbour times 0 would poison the sum)
    T c0 = 0, c1 = 0, c2 = 0, c3 = 0, sc = 1;
    bool use_l1 = false, in_main = false;
    if (cubic) {
        if (k >= 1) {
            if (i + 3 < n) {
                c0 = -1; c1 = 9; c2 = 9; c3 = -1; sc = static_cast<T>(0.0625);
                in_main = true;
            } else if (i + 1 < n) {
                c0 = -1; c1 = 6; c2 = 3; sc = static_cast<T>(0.125);
            } else {
                use_l1 = true;
            }
        } else {
            if (i + 3 < n) {
                c1 = 3; c2 = 6; c3 = -1; sc = static_cast<T>(0.125);
            } else if (i + 1 < n) {
                c1 = 1; c2 = 1; sc = static_cast<T>(0.5);
            } else {
                c1 = 1;
            }
        }
    } else {
        if (i + 1 < n) {
            c1 = 1; c2 = 1; sc = static_cast<T>(0.5);
            in_main = true;
        } else {
            c1 = 1;   // n < 3 (n >= 4 is the merged tail)
        }
    }
    if (SIMPLE) use_l1 = false;
    const bool ld0 = c0 != 0 || use_l1, ld2 = c2 != 0, ld3 = c3 != 0;
    const uint32_t bsel = i == P.bnd[0] ? P.bbase[0] : (i == P.bnd[1] ? P.bbase[1] : P.bbase[2]);
    const uint32_t posk = in_main ? (k - P.kfirst) * P.mD : bsel;
    const uint32_t pU = in_main ? P.mU : P.bU, pV = in_main ? P.mV : P.bV;
    const uint32_t skipu = P.skipu, skipv = P.skipv;
    const uint32_t sU = P.sU, sV = P.sV;
    const uint32_t gU = static_cast<uint32_t>(P.gU), gV = static_cast<uint32_t>(P.gV);
    const uint32_t hU = static_cast<uint32_t>(P.hU), hV = static_cast<uint32_t>(P.hV);
    const uint32_t gD2 = static_cast<uint32_t>(2 * P.gD), hD2 = static_cast<uint32_t>(2 * P.hD);
    const uint32_t tail_pos = P.bbase[0], bU = P.bU, bV = P.bV;
    QT *const qb = A.q + P.base;
    T *const ub = A.unpred_tmp + P.base;
    const T *const gb = A.data + lg.gbase + P.gC + static_cast<uint64_t>(i) * P.gD;
    T *const hb = A.recon2 + lg.g2base + P.hC + static_cast<uint64_t>(i) * P.hD;
    T *const sb = sm + P.sC + k * hS;
    const QuantParams qp = A.qp;
    const uint32_t nrows = cu * cv;
    const uint32_t dU = cv ? nslots / cv : 0, dV = cv ? nslots % cv : 0;
    uint32_t iu = fast_div(slot, P.mgv), iv = slot - iu * cv;
    const bool do_tail = merge_tail && k + 2 == cD;
    if (!act) return;
    if (SIMPLE) {
        // dV == 0: the thread stays on one iv; every offset advances by a constant per visited row
        if (slot >= nrows) return;
        const uint32_t rows = (nrows - slot + nslots - 1) / nslots;
        const bool v_owned = LAST || iv >= skipv;
        // every address is a pointer advanced by a constant
        T *nb = sb + (iu * sU + iv * sV);
        const T *gp = gb + (iu * gU + iv * gV);
        T *hp = hb + (iu * hU + iv * hV);
        // signed: rows on a low face this tile does not own start "before" the pass (never stored)
        const int32_t pos0 = static_cast<int32_t>((iu - skipu) * pU + (iv - skipv) * pV + posk);
        QT *qp_ = qb + pos0;
        T *up = ub + pos0;
        const uint32_t so_inc = dU * sU, go_inc = dU * gU, ho_inc = dU * hU, pos_inc = dU * pU;
        // last pass: the original values stream in from global memory two rows ahead of their use
        T nxt = LAST ? *gp : static_cast<T>(0);
        T nxt2 = LAST && rows > 1 ? gp[go_inc] : static_cast<T>(0);
        for (uint32_t r = rows; r > 0; r--) {
            const T orig = LAST ? nxt : nb[tS];
            nxt = nxt2;
            if (LAST && r > 2) nxt2 = gp[2 * go_inc];
            gp += go_inc;
            const T w1 = nb[0];
            const T w0 = ld0 ? nb[-static_cast<int>(hS)] : static_cast<T>(0);
            const T w2 = ld2 ? nb[hS] : static_cast<T>(0);
            const T w3 = ld3 ? nb[2 * hS] : static_cast<T>(0);
            const T pred = (((c0 * w0 + c1 * w1) + c2 * w2) + c3 * w3) * sc;
            T rec;
            const int qv = quantize<T>(orig, pred, qp, rec);
            if (!LAST) nb[tS] = rec;
            const bool owned = LAST || (v_owned && iu >= skipu);
            if (owned) {
                *qp_ = static_cast<QT>(qv);
                if (qv == 0) store_unpred(up, 0u, orig);
                if (WRITE2) *hp = rec;
            }
            ctx.hist_add(qv, owned);
            nb += so_inc;
            hp += ho_inc;
            qp_ += pos_inc;
            up += pos_inc;
            iu += dU;
        }
        return;
    }
    T nxt = LAST && slot < nrows ? gb[iu * gU + iv * gV] : static_cast<T>(0);
    for (uint32_t row = slot; row < nrows; row += nslots) {
        const uint32_t cu_ = iu, cv_ = iv;   // this row
        iv += dV;
        iu += dU;
        if (iv >= cv) {
            iv -= cv;
            iu++;
        }
        T *const nb = sb + (cu_ * sU + cv_ * sV);
        T orig = LAST ? nxt : nb[tS];
        if (LAST && row + nslots < nrows) nxt = gb[iu * gU + iv * gV];
        const T w1 = nb[0];
        const T w0 = ld0 ? nb[-static_cast<int>(hS)] : static_cast<T>(0);
        const T w2 = ld2 ? nb[hS] : static_cast<T>(0);
        const T w3 = ld3 ? nb[2 * hS] : static_cast<T>(0);
        T pred = (((c0 * w0 + c1 * w1) + c2 * w2) + c3 * w3) * sc;
        if (use_l1) pred = interp_linear1<T>(w0, w1);
        T rec;
        int qv = quantize<T>(orig, pred, qp, rec);
        if (!LAST) nb[tS] = rec;
        const bool owned = LAST || (cu_ >= skipu && cv_ >= skipv);
        const uint32_t ru = cu_ - skipu, rv = cv_ - skipv;
        uint32_t pos = ru * pU + rv * pV + posk;
        const uint32_t hoff = cu_ * hU + cv_ * hV;
        if (owned) {
            qb[pos] = static_cast<QT>(qv);
            if (qv == 0) store_unpred(ub, pos, orig);
            if (WRITE2) hb[hoff] = rec;
        }
        ctx.hist_add(qv, owned);
        if (do_tail) {
            // the linear tail i+2 = n-1: linear1(recon(i), value(i+1))
            const T pred2 = interp_linear1<T>(rec, w2);
            orig = LAST ? gb[cu_ * gU + cv_ * gV + gD2] : nb[tS + hS];
            qv = quantize<T>(orig, pred2, qp, rec);
            if (!LAST) nb[tS + hS] = rec;
            pos = ru * bU + rv * bV + tail_pos;
            if (owned) {
                qb[pos] = static_cast<QT>(qv);
                if (qv == 0) store_unpred(ub, pos, orig);
                if (WRITE2) hb[hoff + hD2] = rec;
            }
            ctx.hist_add(qv, owned);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// fill: the sub-lattice even along the last pass dimension; coarse points (all local indices even) from recon2,
// everything else from the immutable input.  Elements go from global to shared memory as asynchronous copies
// (cp.async, no register staging): a thread issues all of its ~46 copies back to back and waits once, instead of
// paying a global-memory round trip per batch of four loads.  The two loops write disjoint cells, so they need no
// barrier between them.
// ---------------------------------------------------------------------------------------------------------------------
template <class T>
SZ_HD void fill_copy(T *smem_dst, const T *gsrc) {
#ifdef __CUDA_ARCH__
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "cp.async element size");
    const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    if (sizeof(T) == 4)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
#else
    *smem_dst = *gsrc;
#endif
}
SZ_HD void fill_copy_wait() {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

template <class T, class QT, class Ctx>
SZ_HD void line_fill(const InterpArgs<T, QT> &A, Ctx &ctx, const LineGeom &lg, T *sm) {
    const uint32_t tid = ctx.tid(), nt = ctx.nthreads();
    const uint32_t s = A.s;
    const uint32_t m0 = lg.last == 0 ? 2u : 1u, m1 = lg.last == 1 ? 2u : 1u, m2 = lg.last == 2 ? 2u : 1u;
    {   // (1) every element of the sub-lattice that is not a coarse point, from the input
        const uint32_t o0 = lg.last == 0 ? 0u : 1u, o1 = lg.last == 1 ? 0u : 1u, o2 = lg.last == 2 ? 0u : 1u;
        const uint32_t E1 = lg.E[1], E2 = lg.E[2];
        const uint32_t total = lg.E[0] * E1 * E2;
        const uint32_t mg1 = magic_u32_fast(E1), mg2 = magic_u32_fast(E2);
        const uint32_t g0 = m0 * s * static_cast<uint32_t>(A.sh.stride[0]), g1 = m1 * s * static_cast<uint32_t>(A.sh.stride[1]),
                       g2 = m2 * s;
        const T *dat = A.data + lg.gbase;
#pragma unroll 4
        for (uint32_t it = tid; it < total; it += nt) {
            const uint32_t r = fast_div(it, mg2);
            const uint32_t e2 = it - r * E2;
            const uint32_t e0 = fast_div(r, mg1);
            const uint32_t e1 = r - e0 * E1;
            if (((e0 & o0) | (e1 & o1) | (e2 & o2)) & 1u) fill_copy(sm + it, dat + (e0 * g0 + e1 * g1 + e2 * g2));
        }
    }
    {   // (2) coarse points (all local indices even) come as their reconstruction from recon2:
        //     local index 2c of dim d sits at smem coordinate 2c (d != last) or c (d == last) and at recon2 offset
        //     (begin + 2c*s)/2 = begin/2 + c*s
        const uint32_t C0 = (lg.n[0] + 1) / 2, C1 = (lg.n[1] + 1) / 2, C2 = (lg.n[2] + 1) / 2;
        const uint32_t total = C0 * C1 * C2;
        const uint32_t mg1 = magic_u32_fast(C1), mg2 = magic_u32_fast(C2);
        const uint32_t h0 = s * static_cast<uint32_t>(A.stride2[0]), h1 = s * static_cast<uint32_t>(A.stride2[1]), h2 = s;
        const uint32_t t2 = 2u / m2, t1 = (2u / m1) * lg.E[2], t0 = (2u / m0) * lg.E[2] * lg.E[1];
        const T *rc2 = A.recon2 + lg.g2base;
#pragma unroll 4
        for (uint32_t it = tid; it < total; it += nt) {
            const uint32_t r = fast_div(it, mg2);
            const uint32_t c2 = it - r * C2;
            const uint32_t c0 = fast_div(r, mg1);
            const uint32_t c1 = r - c0 * C1;
            fill_copy(sm + (c0 * t0 + c1 * t1 + c2 * t2), rc2 + (c0 * h0 + c1 * h1 + c2 * h2));
        }
    }
    fill_copy_wait();   // the caller's barrier publishes the tile
}

template <class T, class QT, class Ctx, bool LAST, bool WRITE2>
SZ_HD void line_pass_run(const InterpArgs<T, QT> &A, Ctx &ctx, T *sm, const LineGeom &lg, const LinePass &P,
                         bool cubic) {
    if (P.kind == 0) {
        line_items<T, QT, Ctx, LAST, WRITE2>(A, ctx, lg, P, sm, cubic);
    } else {
        const uint32_t nslots = ctx.nthreads() >> 4;
        const bool simple = (P.n & 1u) && P.cv != 0 && nslots % P.cv == 0;
        if (simple)
            row_items<T, QT, Ctx, LAST, WRITE2, true>(A, ctx, lg, P, sm, cubic);
        else
            row_items<T, QT, Ctx, LAST, WRITE2, false>(A, ctx, lg, P, sm, cubic);
    }
}

template <class T, class QT, class Ctx>
SZ_HD void line_tile_passes(const InterpArgs<T, QT> &A, Ctx &ctx, T *sm, const LineGeom &lg, const LineTile &lt) {
    const bool cubic = A.sh.cubic != 0;
    const bool write2 = A.s >= 2;
    for (int p = 0; p < 3; p++) {
        const LinePass &P = lt.ps[p];
        if (P.cu != 0 && P.cv != 0 && P.cD != 0) {
            if (p < 2) {
                if (write2)
                    line_pass_run<T, QT, Ctx, false, true>(A, ctx, sm, lg, P, cubic);
                else
                    line_pass_run<T, QT, Ctx, false, false>(A, ctx, sm, lg, P, cubic);
            } else {
                if (write2)
                    line_pass_run<T, QT, Ctx, true, true>(A, ctx, sm, lg, P, cubic);
                else
                    line_pass_run<T, QT, Ctx, true, false>(A, ctx, sm, lg, P, cubic);
            }
        }
        ctx.pass_end();
        if (p < 2) ctx.sync();
    }
}

}  // namespace sz3b
