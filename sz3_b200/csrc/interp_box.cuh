// sz3_b200/csrc/interp_box.cuh -- box schedule: the N == 3 tile schedule of the fused interpolation-predict +
// LinearQuantizer kernel (InterpolationDecomposition::compress, reference
// include/SZ3/decomposition/InterpolationDecomposition.hpp:99-143, :309-402; LinearQuantizer.hpp:43-71) for the
// configuration the auto-tuner picks most often: float data, pass order z, y, x (interpDirection 0), tiles whose
// three extents are 32 or 33 points (every tile of an array whose dims are multiples of 32, e.g. 512^3 / 2048^3).
//
// Same closed-tile mathematics as the earlier generations (a level block recomputed standalone
// from {original values} U {final reconstruction of its coarse points} reproduces the reference bit for bit), but the
// work is cut so that no thread ever decodes an index:
//
//   phase A (CTA)   EE = the sub-lattice even in y and x (33 x 17 x 17): coarse points from recon2, the rest from the
//                   input; thread c owns the z-line c = y' * 17 + x': it fills the line, then predicts + quantizes its
//                   16 targets from registers (pass 0, fully unrolled, stencil per target known at compile time).
//   phase B (warp)  after pass 0 the z-planes are independent.  A warp owns planes z = low + w, + 8, ...; the raw plane
//                   (33 rows x 36 floats) arrives in the warp's slot by ONE TMA box copy (cp.async.bulk.tensor), the
//                   even-even points are overwritten with EE, pass 1 runs on lanes (x', half of the targets), pass 2
//                   on lanes = rows: a lane reads its whole row with nine 16-byte shared-memory loads (row pitch of
//                   nine 16-byte chunks: conflict-free) and walks its 16 targets from registers.
//
// Index positions are closed forms of (z, y, x) with per-tile multipliers (BoxTile); the 14 (or 13..16) main-phase
// indices of a row are staged in shared memory so that a plane's run leaves as 16-byte stores.
//
// The per-lane phase functions below are plain inline code: tests/emul runs them lane by lane on the CPU (with a
// host copy standing in for the TMA box) against the reference before any GPU time is spent.
#pragma once
#include "core.cuh"
#include "interp_body.cuh"
#include "interp_line.cuh"

namespace sz3b {

constexpr int kBoxWarps = 8;
constexpr int kBoxThreads = kBoxWarps * 32;
constexpr int kBoxPitch = 36;                    // floats per slot row: 144 B = nine 16-byte chunks
constexpr int kBoxSlotElems = 33 * kBoxPitch;    // one raw plane
constexpr int kBoxSlotStride = 1216;             // floats between the slots of two warps (4864 B, a multiple of 128 B)
constexpr int kBoxEEPlane = 17 * 17;
constexpr int kBoxEEElems = 33 * kBoxEEPlane;
constexpr int kBoxStageU16 = 576;                // per warp: 8 (misalignment) + 33 rows * 16 indices, rounded up

using BoxArgs = InterpArgs<float, uint16_t>;

#ifdef __CUDACC__
using BoxChunk = uint4;
#else
struct alignas(16) BoxChunk {
    uint32_t w[4];
};
#endif

// Where a level's input values come from: the input array itself (one lattice step = s elements per dim) or a compact
// copy of the level's lattice (interp_box.cu: k_box_compact; unit steps, so that coarse levels read dense memory and
// can use the TMA path).  Lattice point (l0, l1, l2) of the tile with origin `begin` is
// p[sum_d (begin[d] / odiv) * ost[d] + l[d] * st[d]].
struct BoxSrc {
    const float *p;
    uint64_t ost[3];   // stride of the tile origin (per unit of begin / odiv)
    uint64_t st[3];    // one lattice step
    uint32_t odiv;
    uint32_t tma;      // planes arrive by TMA box copies (unit x step, 16-byte aligned rows) or by per-element copies
    uint32_t split;    // CTAs per tile (coarse levels with fewer tiles than SMs): every CTA of a tile runs phase A, the
                       // first one emits it, and the planes of phase B are dealt out over split * kBoxWarps warps
};

// Per-tile constants (shared memory; written by one thread while the others fill EE).
struct BoxTile {
    uint32_t n[3];       // points per dim (32 or 33), natural order z, y, x
    uint32_t low[3];     // 1 when the low face belongs to the previous tile (begin > 0)
    uint32_t C[3];       // lattice points at step 2: (n + 1) / 2
    uint32_t c1[3];      // owned points at step 1: n - low
    uint32_t c2[3];      // owned lattice points at step 2: C - low
    uint32_t mainc[3];   // targets of the main sub-phase along the dim
    uint32_t pb[3];      // first index of each pass, relative to qbase
    uint32_t other[3];   // lines of each pass (product of the owned counts of the two other dims)
    uint64_t qbase;      // position of the tile's first index in the stream
    uint64_t sbase;      // element offset of the tile's origin in the source array (BoxSrc)
    uint64_t g2base;     // ... and in recon2
};

// What every thread needs before BoxTile is ready (a few integer operations, recomputed per thread).
struct BoxOrigin {
    uint32_t begin[3], n[3];
    uint64_t sbase, g2base;
};

SZ_HD void box_origin(const BoxArgs &A, const BoxSrc &S, uint32_t tile, BoxOrigin &o) {
    uint32_t bidx[3];
    uint32_t r = tile;
    bidx[2] = r % A.nb[2];
    r /= A.nb[2];
    bidx[1] = r % A.nb[1];
    bidx[0] = r / A.nb[1];
    const uint32_t B = kInterpBlock * A.s;
    o.sbase = 0;
    o.g2base = 0;
    for (int d = 0; d < 3; d++) {
        const uint32_t b = bidx[d] * B;
        uint32_t e = b + B;
        if (e > A.sh.dims[d] - 1) e = A.sh.dims[d] - 1;
        o.begin[d] = b;
        o.n[d] = (e - b) / A.s + 1;
        o.sbase += static_cast<uint64_t>(b / S.odiv) * S.ost[d];
        o.g2base += static_cast<uint64_t>(b >> 1) * A.stride2[d];
    }
}

// Host + device: can this tile take the box schedule?
SZ_HD bool box_tile_ok(const BoxOrigin &o) {
    return (o.n[0] == 32 || o.n[0] == 33) && (o.n[1] == 32 || o.n[1] == 33) && (o.n[2] == 32 || o.n[2] == 33);
}

template <bool CUBIC>
SZ_HD void box_tile_setup(const BoxArgs &A, uint32_t tile, const BoxOrigin &o, BoxTile &T) {
    for (int d = 0; d < 3; d++) {
        const uint32_t n = o.n[d], low = o.begin[d] ? 1u : 0u;
        T.n[d] = n;
        T.low[d] = low;
        T.C[d] = (n + 1) / 2;
        T.c1[d] = n - low;
        T.c2[d] = (n + 1) / 2 - low;
        T.mainc[d] = CUBIC ? (n - 7) / 2 + 1 : (n - 1) / 2;
    }
    // boundary sub-phases along a dim: cubic i = 1 and (n odd: n-2 | n even: n-3, n-1); linear n even: n-1
    uint32_t nbnd[3];
    for (int d = 0; d < 3; d++) nbnd[d] = CUBIC ? ((T.n[d] & 1u) ? 2u : 3u) : ((T.n[d] & 1u) ? 0u : 1u);
    T.other[0] = T.c2[1] * T.c2[2];
    T.other[1] = T.c1[0] * T.c2[2];
    T.other[2] = T.c1[0] * T.c1[1];
    T.pb[0] = 0;
    T.pb[1] = (T.mainc[0] + nbnd[0]) * T.other[0];
    T.pb[2] = T.pb[1] + (T.mainc[1] + nbnd[1]) * T.other[1];
    T.qbase = A.block_base[tile];
    T.sbase = o.sbase;
    T.g2base = o.g2base;
}

// Sub-phase of target k (local index 2k + 1) on a line of n = 33 (n_odd) or 32 points: true = main sub-phase with
// index idx inside it, false = boundary sub-phase number idx (reference order, InterpolationDecomposition.hpp:334-400).
// With k a compile-time constant everything but the last two targets folds away.
template <bool CUBIC>
SZ_HD bool box_class(uint32_t k, bool n_odd, uint32_t &idx) {
    if (CUBIC) {
        if (k == 0) {
            idx = 0;
            return false;
        }
        if (k <= 13) {
            idx = k - 1;
            return true;
        }
        if (k == 14) {            // i = 29: main for n = 33, the n-3 boundary for n = 32
            idx = n_odd ? 13u : 1u;
            return n_odd;
        }
        idx = n_odd ? 1u : 2u;    // i = 31: n-2 (n = 33) or n-1 (n = 32)
        return false;
    }
    if (k <= 14) {
        idx = k;
        return true;
    }
    idx = n_odd ? 15u : 0u;       // i = 31: main for n = 33, the tail for n = 32
    return n_odd;
}

// Prediction of target k from the line's values v[local index] (n = 32 or 33 points); rec_prev = reconstruction of
// target k - 1 (the linear tail of an even line needs it).  k is a compile-time constant after unrolling.
template <bool CUBIC>
SZ_HD float box_pred(const float *v, int k, bool n_odd, float rec_prev) {
    if (CUBIC) {
        if (k == 0) return interp_quad_1<float>(v[0], v[2], v[4]);
        if (k <= 13) return interp_cubic<float>(v[2 * k - 2], v[2 * k], v[2 * k + 2], v[2 * k + 4]);
        if (k == 14)
            return n_odd ? interp_cubic<float>(v[26], v[28], v[30], v[32]) : interp_quad_2<float>(v[26], v[28], v[30]);
        return n_odd ? interp_quad_2<float>(v[28], v[30], v[32]) : interp_linear1<float>(v[28], v[30]);
    }
    if (k <= 14) return interp_linear<float>(v[2 * k], v[2 * k + 2]);
    return n_odd ? interp_linear<float>(v[30], v[32]) : interp_linear1<float>(rec_prev, v[30]);
}

// Where a line's indices go: the main sub-phase at qm[idx * main_mul], boundary sub-phase j at qm[bnd0 + j * other]
// (element offsets relative to qm; everything a register after inlining).  um is the same place in unpred_tmp.
struct BoxEmit {
    uint16_t *qm;
    float *um;
    uint32_t main_mul, bnd0, other;
};
SZ_HD uint32_t box_off(const BoxEmit &E, bool in_main, uint32_t idx) {
    return in_main ? idx * E.main_mul : E.bnd0 + idx * E.other;
}
SZ_HD void box_store(const BoxEmit &E, bool in_main, uint32_t idx, int qv, float orig, bool owned) {
    const uint32_t off = box_off(E, in_main, idx);
    if (owned) {
        E.qm[off] = static_cast<uint16_t>(qv);
        if (qv == 0) E.um[off] = orig;
    }
}
// runtime-k twin of box_class (leftover lines: lane = target)
template <bool CUBIC>
SZ_HD bool box_class_rt(uint32_t k, bool n_odd, uint32_t &idx) {
    if (CUBIC) {
        const bool in_main = (k >= 1 && k <= 13) || (k == 14 && n_odd);
        idx = in_main ? k - 1 : (k == 0 ? 0u : (k == 14 ? 1u : (n_odd ? 1u : 2u)));
        return in_main;
    }
    const bool in_main = k <= 14 || n_odd;
    idx = in_main ? k : 0u;
    return in_main;
}

// Indices of a run of N targets handled by one thread.  The box schedule does not count symbols itself (the histogram
// of HuffmanEncoder::init is taken by a pass over the index stream afterwards, encode_kernels.cu: k_hist_u16 --
// counting here cost a quarter of the kernel's time); what it has to notice is index 0, an unpredictable point, whose
// original value goes to unpred_tmp: the run keeps the minimum of its indices, and only a run that saw a 0 looks for
// it (`zero(k)`; the caller re-reads the value from the input array, nothing stays in registers for it).
template <class Ctx, int N>
struct BoxHist {
    int q[N];
    int qmin = 0x7fffffff;
    SZ_HD void add(Ctx &, int k, int qv, bool) {
        q[k] = qv;
        qmin = qv < qmin ? qv : qmin;
    }
    template <class Zero>
    SZ_HD void finish(Ctx &, bool owned, Zero &&zero) {
        if (qmin == 0 && owned) {
#pragma unroll
            for (int k = 0; k < N; k++)
                if (q[k] == 0) zero(k);
        }
    }
};

// stencils whose kind depends on the lane: four taps with coefficients (unused taps must be zero, never NaN)
enum { BOX_ST_CUBIC = 0, BOX_ST_QUAD1 = 1, BOX_ST_QUAD2 = 2, BOX_ST_LINEAR1 = 3 };
SZ_HD float box_stencil4(uint32_t kind, float w0, float w1, float w2, float w3) {
    // (((c0*w0 + c1*w1) + c2*w2) + c3*w3) * sc evaluates cubic / quad_1 / quad_2 in the reference's operation order
    const float c0 = kind == BOX_ST_QUAD1 ? 0.0f : -1.0f;
    const float c1 = kind == BOX_ST_CUBIC ? 9.0f : (kind == BOX_ST_QUAD1 ? 3.0f : 6.0f);
    const float c2 = kind == BOX_ST_CUBIC ? 9.0f : (kind == BOX_ST_QUAD1 ? 6.0f : 3.0f);
    const float c3 = kind == BOX_ST_QUAD2 ? 0.0f : -1.0f;
    const float sc = kind == BOX_ST_CUBIC ? 0.0625f : 0.125f;
    const float p = (((c0 * w0 + c1 * w1) + c2 * w2) + c3 * w3) * sc;
    return kind == BOX_ST_LINEAR1 ? interp_linear1<float>(w0, w1) : p;
}

// One target with a run-time k (lane = target) without divergent code: the few lines / rows / columns that do not
// fit the lane mappings (the 17th lattice column, a 33rd owned row, z-lines 256..288).  ld(l) = current value at local
// index l (only indices inside the line are requested), st(l, v) stores a reconstruction, em(k, qv, orig) emits.
// Same arithmetic as box_pred: the stencil kinds of the reference (InterpolationDecomposition.hpp:334-400) for
// n = 32 / 33, evaluated through box_stencil4.
template <bool CUBIC, class Ld, class St, class Em>
SZ_HD void box_target_rt(uint32_t k, bool n_odd, const QuantParams &qp, Ld &&ld, St &&st, Em &&em) {
    const bool tail_pair = !CUBIC && !n_odd;      // linear, even n: target 15 needs the reconstruction of target 14
    if (tail_pair && k == 15) return;             // done together with its predecessor
    float pred;
    if (CUBIC) {
        const uint32_t kind = k == 0 ? BOX_ST_QUAD1
                                     : (k <= 13 ? BOX_ST_CUBIC
                                                : (k == 14 ? (n_odd ? BOX_ST_CUBIC : BOX_ST_QUAD2) : (n_odd ? BOX_ST_QUAD2 : BOX_ST_LINEAR1)));
        const float w0 = k >= 1 ? ld(2 * k - 2) : 0.0f;
        const float w1 = ld(2 * k);
        const float w2 = kind != BOX_ST_LINEAR1 ? ld(2 * k + 2) : 0.0f;
        const float w3 = (kind == BOX_ST_CUBIC || kind == BOX_ST_QUAD1) ? ld(2 * k + 4) : 0.0f;
        pred = box_stencil4(kind, w0, w1, w2, w3);
    } else {
        pred = interp_linear<float>(ld(2 * k), ld(2 * k + 2));
    }
    float orig = ld(2 * k + 1), rec;
    int qv = quantize_f32(orig, pred, qp, rec);
    st(2 * k + 1, rec);
    em(k, qv, orig);
    if (tail_pair && k == 14) {
        pred = interp_linear1<float>(rec, ld(30));
        orig = ld(31);
        qv = quantize_f32(orig, pred, qp, rec);
        st(31, rec);
        em(15, qv, orig);
    }
}

// Eight consecutive targets k = 8h + t of a line of n = 32 | 33 points, the unit every pass is cut into (one compact
// body instead of sixteen unrolled targets per pass: the kernel has to stay inside the instruction cache).
//   nb[m]  value at lattice index 8h - 1 + m, i.e. local index 2 (8h - 1 + m); entries outside the line must be 0
//   og[t]  original value of target t (local index 16h + 2t + 1)
// Results: rc[t] reconstructions, H.q[t] indices (counted in the histogram when `owned`).
template <bool CUBIC, class Ctx>
SZ_HD void box_run8(const float (&nb)[11], const float (&og)[8], uint32_t h, bool n_odd, const QuantParams &qp, Ctx &ctx,
                    bool owned, float (&rc)[8], BoxHist<Ctx, 8> &H) {
    // Predictions first, then the eight quantizer chains (interleaved by the compiler).  The stencils that differ from
    // the plain cubic sit at the two ends of a line; which one applies depends on (h, n_odd) only -- uniform over the
    // warp, so they are ordinary branches around the exact interpolator instead of a four-tap form with selected
    // coefficients (which also evaluated the double-precision linear tail for every line).
    float pred[8];
    if (CUBIC) {
#pragma unroll
        for (int t = 1; t <= 5; t++) pred[t] = interp_cubic<float>(nb[t], nb[t + 1], nb[t + 2], nb[t + 3]);
        if (h == 0) {
            pred[0] = interp_quad_1<float>(nb[1], nb[2], nb[3]);
            pred[6] = interp_cubic<float>(nb[6], nb[7], nb[8], nb[9]);
            pred[7] = interp_cubic<float>(nb[7], nb[8], nb[9], nb[10]);
        } else {
            pred[0] = interp_cubic<float>(nb[0], nb[1], nb[2], nb[3]);
            if (n_odd) {
                pred[6] = interp_cubic<float>(nb[6], nb[7], nb[8], nb[9]);
                pred[7] = interp_quad_2<float>(nb[7], nb[8], nb[9]);
            } else {
                pred[6] = interp_quad_2<float>(nb[6], nb[7], nb[8]);
                pred[7] = interp_linear1<float>(nb[7], nb[8]);
            }
        }
    } else {
#pragma unroll
        for (int t = 0; t < 8; t++) pred[t] = interp_linear<float>(nb[t + 1], nb[t + 2]);
    }
    const bool lin_tail = !CUBIC && h == 1 && !n_odd;   // linear, even line: the last target leans on its predecessor
#pragma unroll
    for (int t = 0; t < 8; t++) {
        float p = pred[t];
        if (!CUBIC && t == 7 && lin_tail) p = interp_linear1<float>(rc[6], nb[8]);
        const int qv = quantize_f32(og[t], p, qp, rc[t]);
        H.add(ctx, t, qv, owned);
    }
}

// Element offset (relative to E.qm) of target k = 8h + t: the regular main-phase ones advance by main_mul, the first
// and the last two of a line sit in boundary sub-phases.
template <bool CUBIC>
SZ_HD uint32_t box_off8(const BoxEmit &E, uint32_t h, bool n_odd, int t) {
    const uint32_t reg = (8u * h + static_cast<uint32_t>(t) - (CUBIC ? 1u : 0u)) * E.main_mul;
    const uint32_t b1 = E.bnd0 + E.other, b2 = b1 + E.other;
    if (CUBIC) {
        if (t == 0) return h ? reg : E.bnd0;
        if (t == 6) return (h && !n_odd) ? b1 : reg;
        if (t == 7) return h ? (n_odd ? b1 : b2) : reg;
        return reg;
    }
    if (t == 7) return (h && !n_odd) ? E.bnd0 : reg;
    return reg;
}

// ---------------------------------------------------------------------------------------------------------------------
// phase A: EE fill, thread c fills z-line c (c = y' * 17 + x'); asynchronous 4-byte copies, one wait at the end
// ---------------------------------------------------------------------------------------------------------------------
SZ_HD void box_fill_column(const BoxArgs &A, const BoxSrc &S, const BoxOrigin &o, uint32_t c, float *EE) {
    const uint32_t yl = c / 17u, xl = c - yl * 17u;
    if (yl >= (o.n[1] + 1) / 2 || xl >= (o.n[2] + 1) / 2) return;
    const uint32_t s = A.s;
    // coarse points (z even): recon2 offset of local (2a, 2b, 2c) is g2base + (a*stride2[0] + b*stride2[1] + c) * s
    const float *rc = A.recon2 + o.g2base + static_cast<uint64_t>(yl) * s * A.stride2[1] + static_cast<uint64_t>(xl) * s;
    const uint64_t rstep = static_cast<uint64_t>(s) * A.stride2[0];
    // originals (z odd) at local (z, 2y', 2x')
    const uint64_t gstep = S.st[0];
    const float *gp = S.p + o.sbase + 2 * yl * S.st[1] + 2 * xl * S.st[2] + gstep;
    float *dst = EE + c;
#pragma unroll
    for (int z = 0; z < 32; z += 2) {
        fill_copy(dst + z * kBoxEEPlane, rc);
        fill_copy(dst + (z + 1) * kBoxEEPlane, gp);
        rc += rstep;
        gp += 2 * gstep;
    }
    if (o.n[0] & 1u) fill_copy(dst + 32 * kBoxEEPlane, rc);
}

// One raw plane of the tile into a warp's slot by per-element copies (levels whose source cannot feed the TMA path:
// the input array at stride 2, unaligned rows): lane = x, walking over the rows; then the 33rd column.
SZ_HD void box_gather_plane(const BoxSrc &S, const BoxTile &T, uint32_t lane, uint32_t z, float *slot) {
    const float *src = S.p + T.sbase + z * S.st[0];
    const uint32_t ny = T.n[1], nx = T.n[2];
    if (lane < nx) {
        const float *p = src + lane * S.st[2];
        float *d = slot + lane;
        for (uint32_t y = 0; y < ny; y++) {
            fill_copy(d, p);
            d += kBoxPitch;
            p += S.st[1];
        }
    }
    if (nx == 33) {
        for (uint32_t y = lane; y < ny; y += 32) fill_copy(slot + y * kBoxPitch + 32, src + y * S.st[1] + 32 * S.st[2]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// pass 0 (along z), thread c < 256 on its own z-line; everything from registers
// ---------------------------------------------------------------------------------------------------------------------
template <class Ctx>
SZ_HD void box_pass0_emit(const BoxArgs &A, const BoxTile &T, uint32_t yl, uint32_t xl, BoxEmit &E, bool &owned) {
    const uint32_t lowy = T.low[1], lowx = T.low[2], other = T.other[0];
    owned = yl >= lowy && xl >= lowx;
    const uint32_t line_rank = (yl - lowy) * T.c2[2] + (xl - lowx);
    E.qm = A.q + T.qbase + line_rank;   // pb[0] == 0
    E.um = A.unpred_tmp + T.qbase + line_rank;
    E.main_mul = other;
    E.bnd0 = T.mainc[0] * other;
    E.other = other;
}

template <bool CUBIC, class Ctx>
SZ_HD void box_pass0_line(const BoxArgs &A, const BoxSrc &S, Ctx &ctx, const BoxTile &T, uint32_t c, float *EE, bool emit) {
    const uint32_t yl = c / 17u, xl = c - yl * 17u;
    if (yl >= T.C[1] || xl >= T.C[2]) return;
    const bool n_odd = T.n[0] & 1u;
    BoxEmit E;
    bool owned;
    box_pass0_emit<Ctx>(A, T, yl, xl, E, owned);
    owned = owned && emit;
    const QuantParams qp = A.qp;
    // where the line's originals sit in the source (rare path: value of an unpredictable point)
    const uint64_t gstep = S.st[0];
    const float *const gcol = S.p + T.sbase + 2 * yl * S.st[1] + 2 * xl * S.st[2];
#pragma unroll 1
    for (uint32_t h = 0; h < 2; h++) {
        float *const col = EE + c + h * (16 * kBoxEEPlane);   // local z = 16h
        float nb[11], og[8], rc[8];
        nb[0] = h ? col[-2 * kBoxEEPlane] : 0.0f;
#pragma unroll
        for (int m = 1; m < 9; m++) nb[m] = col[(2 * m - 2) * kBoxEEPlane];
        nb[9] = (h == 0 || n_odd) ? col[16 * kBoxEEPlane] : 0.0f;
        nb[10] = h ? 0.0f : col[18 * kBoxEEPlane];
#pragma unroll
        for (int t = 0; t < 8; t++) og[t] = col[(2 * t + 1) * kBoxEEPlane];
        BoxHist<Ctx, 8> H;
        box_run8<CUBIC>(nb, og, h, n_odd, qp, ctx, owned, rc, H);
#pragma unroll
        for (int t = 0; t < 8; t++) {
            col[(2 * t + 1) * kBoxEEPlane] = rc[t];
            if (owned) E.qm[box_off8<CUBIC>(E, h, n_odd, t)] = static_cast<uint16_t>(H.q[t]);
        }
        H.finish(ctx, owned, [&](int t) { E.um[box_off8<CUBIC>(E, h, n_odd, t)] = gcol[static_cast<uint64_t>(16 * h + 2 * t + 1) * gstep]; });
    }
}

// z-lines 256..288 (the 17th lattice row / column spill-over of the 256-thread mapping): item e = (line, target)
template <bool CUBIC, class Ctx>
SZ_HD void box_pass0_left(const BoxArgs &A, Ctx &ctx, const BoxTile &T, uint32_t e, float *EE, bool emit) {
    const uint32_t c = 256u + (e >> 4), k = e & 15u;
    if (c >= static_cast<uint32_t>(kBoxEEPlane)) return;
    const uint32_t yl = c / 17u, xl = c - yl * 17u;
    if (yl >= T.C[1] || xl >= T.C[2]) return;
    const bool n_odd = T.n[0] & 1u;
    BoxEmit E;
    bool owned;
    box_pass0_emit<Ctx>(A, T, yl, xl, E, owned);
    owned = owned && emit;
    float *const col = EE + c;
    box_target_rt<CUBIC>(
        k, n_odd, A.qp, [&](uint32_t l) { return col[l * kBoxEEPlane]; }, [&](uint32_t l, float r) { col[l * kBoxEEPlane] = r; },
        [&](uint32_t kk, int qv, float orig) {
            uint32_t idx;
            const bool in_main = box_class_rt<CUBIC>(kk, n_odd, idx);
            box_store(E, in_main, idx, qv, orig, owned);
        });
}

// ---------------------------------------------------------------------------------------------------------------------
// phase B, per warp and plane z.  slot = the raw plane [33][36]; EEz = EE + z * 289.
// ---------------------------------------------------------------------------------------------------------------------
// even-even points of the plane take their current values (coarse points, pass-0 reconstructions)
SZ_HD void box_merge(const BoxTile &T, uint32_t lane, const float *EEz, float *slot) {
    const uint32_t xl = lane & 15u, h = lane >> 4;
    const uint32_t Cy = T.C[1];
    const float *src = EEz + h * 17 + xl;
    float *dst = slot + 2 * h * kBoxPitch + 2 * xl;
#pragma unroll
    for (int i = 0; i < 8; i++) dst[4 * i * kBoxPitch] = src[34 * i];    // y' = h + 2i <= 15
    if (h == 0 && Cy == 17) dst[32 * kBoxPitch] = src[34 * 8];           // y' = 16
    if (T.C[2] == 17 && lane < Cy) slot[2 * lane * kBoxPitch + 32] = EEz[lane * 17 + 16];
}

// pass 1 (along y): lane = (x' = lane & 15, half h of the 16 targets); neighbours from EE (conflict-free), targets
// read and overwritten in the slot
template <class Ctx>
SZ_HD void box_pass1_emit(const BoxArgs &A, const BoxTile &T, uint32_t z, uint32_t xl, BoxEmit &E, bool &owned) {
    const uint32_t lowx = T.low[2], c2x = T.c2[2], mainc = T.mainc[1], other = T.other[1];
    const uint32_t rz = z - T.low[0];
    const uint64_t base = T.qbase + T.pb[1];
    owned = xl >= lowx;
    // main: (rz * mainc + idx) * c2x + rx; boundary j: (mainc + j) * other + rz * c2x + rx
    const uint32_t line_rank = rz * c2x + (xl - lowx);
    const uint32_t m0 = rz * mainc * c2x + (xl - lowx);
    E.qm = A.q + base + m0;
    E.um = A.unpred_tmp + base + m0;
    E.main_mul = c2x;
    E.bnd0 = mainc * other + line_rank - m0;
    E.other = other;
}

template <bool CUBIC, class Ctx>
SZ_HD void box_pass1_lane(const BoxArgs &A, const BoxSrc &S, Ctx &ctx, const BoxTile &T, uint32_t lane, uint32_t z,
                          const float *EEz, float *slot) {
    const uint32_t xl = lane & 15u, h = lane >> 4;
    const bool n_odd = T.n[1] & 1u;
    float nb[11], og[8], rc[8];
    {
        const float *src = EEz + xl + (h ? 7 * 17 : -17);   // lattice index 8h - 1
        nb[0] = h ? src[0] : 0.0f;
#pragma unroll
        for (int m = 1; m < 9; m++) nb[m] = src[m * 17];
        nb[9] = (h == 0 || n_odd) ? src[9 * 17] : 0.0f;
        nb[10] = h ? 0.0f : src[10 * 17];
    }
    float *const col = slot + 2 * xl + h * (16 * kBoxPitch);
#pragma unroll
    for (int t = 0; t < 8; t++) og[t] = col[(2 * t + 1) * kBoxPitch];
    BoxEmit E;
    bool owned;
    box_pass1_emit<Ctx>(A, T, z, xl, E, owned);
    BoxHist<Ctx, 8> H;
    box_run8<CUBIC>(nb, og, h, n_odd, A.qp, ctx, owned, rc, H);
#pragma unroll
    for (int t = 0; t < 8; t++) {
        col[(2 * t + 1) * kBoxPitch] = rc[t];
        if (owned) E.qm[box_off8<CUBIC>(E, h, n_odd, t)] = static_cast<uint16_t>(H.q[t]);
    }
    H.finish(ctx, owned, [&](int t) {
        // rare: the original of an unpredictable point, re-read from the input at local (z, 2k + 1, 2x')
        const uint64_t g = T.sbase + z * S.st[0] + (16 * h + 2 * t + 1) * S.st[1] + 2 * xl * S.st[2];
        E.um[box_off8<CUBIC>(E, h, n_odd, t)] = S.p[g];
    });
}

// the 17th lattice column (x = 32) of pass 1: lane = target
template <bool CUBIC, class Ctx>
SZ_HD void box_pass1_left(const BoxArgs &A, Ctx &ctx, const BoxTile &T, uint32_t lane, uint32_t z, const float *EEz,
                          float *slot) {
    if (T.C[2] != 17 || lane >= 16) return;
    const bool n_odd = T.n[1] & 1u;
    BoxEmit E;
    bool owned;
    box_pass1_emit<Ctx>(A, T, z, 16u, E, owned);
    const float *ecol = EEz + 16;
    float *scol = slot + 32;
    box_target_rt<CUBIC>(
        lane, n_odd, A.qp, [&](uint32_t l) { return (l & 1u) ? scol[l * kBoxPitch] : ecol[(l >> 1) * 17]; },
        [&](uint32_t l, float r) { scol[l * kBoxPitch] = r; },
        [&](uint32_t kk, int qv, float orig) {
            uint32_t idx;
            const bool in_main = box_class_rt<CUBIC>(kk, n_odd, idx);
            box_store(E, in_main, idx, qv, orig, true);
        });
}

// Position of a plane's main-phase run of pass 2 and its misalignment (in indices) against 16-byte chunks.
SZ_HD uint64_t box_run_pos(const BoxTile &T, uint32_t z) {
    return T.qbase + T.pb[2] + static_cast<uint64_t>((z - T.low[0]) * T.c1[1]) * T.mainc[2];
}

// pass 2 (along x), lane = row: `v` are the row's 36 floats (registers).  Main-phase indices go to the warp's
// staging buffer at the row's place inside the plane's run; boundary sub-phases straight to global memory (lanes
// sit on consecutive positions).
struct BoxRowOut {
    uint16_t *srow;      // staging: main sub-phase of this row
    uint16_t *qb;        // global: boundary sub-phase 0 of this row
    float *um, *ub;      // unpred_tmp at the same two places
    uint32_t other;
};
SZ_HD void box_row_out(const BoxArgs &A, const BoxTile &T, uint32_t ry, uint32_t z, uint16_t *stage, BoxRowOut &R) {
    const uint32_t mainc = T.mainc[2], c1y = T.c1[1];
    const uint32_t line_rank = (z - T.low[0]) * c1y + ry;
    const uint64_t base = T.qbase + T.pb[2];
    const uint32_t mis = static_cast<uint32_t>(box_run_pos(T, z)) & 7u;
    R.srow = stage + mis + ry * mainc;
    R.other = T.other[2];
    R.qb = A.q + base + mainc * R.other + line_rank;
    R.ub = A.unpred_tmp + base + mainc * R.other + line_rank;
    R.um = A.unpred_tmp + base + line_rank * mainc;
}
SZ_HD void box_row_store(const BoxRowOut &R, bool in_main, uint32_t idx, int qv, float orig) {
    if (in_main) {
        R.srow[idx] = static_cast<uint16_t>(qv);
        if (qv == 0) R.um[idx] = orig;
    } else {
        R.qb[idx * R.other] = static_cast<uint16_t>(qv);
        if (qv == 0) R.ub[idx * R.other] = orig;
    }
}

template <bool CUBIC, class Ctx>
SZ_HD void box_pass2_row(const BoxArgs &A, const BoxSrc &S, Ctx &ctx, const BoxTile &T, uint32_t ry, uint32_t z,
                         const float *v, uint16_t *stage, float *slot_row) {
    const bool n_odd = T.n[2] & 1u;
    BoxRowOut R;
    box_row_out(A, T, ry, z, stage, R);
    const QuantParams qp = A.qp;
#pragma unroll 1
    for (uint32_t h = 0; h < 2; h++) {
        // the half's window out of the row registers: nb[m] = v[16h - 2 + 2m], og[t] = v[16h + 2t + 1]
        float nb[11], og[8], rc[8];
        nb[0] = h ? v[14] : 0.0f;
#pragma unroll
        for (int m = 1; m < 9; m++) nb[m] = h ? v[14 + 2 * m] : v[2 * m - 2];
        nb[9] = h ? (n_odd ? v[32] : 0.0f) : v[16];
        nb[10] = h ? 0.0f : v[18];
#pragma unroll
        for (int t = 0; t < 8; t++) og[t] = h ? v[17 + 2 * t] : v[2 * t + 1];
        BoxHist<Ctx, 8> H;
        box_run8<CUBIC>(nb, og, h, n_odd, qp, ctx, true, rc, H);
        if (slot_row) {   // level stride >= 2: the plane's reconstructions leave for recon2 from the slot (box_plane_out)
#pragma unroll
            for (int t = 0; t < 8; t++) slot_row[16 * h + 2 * t + 1] = rc[t];
        }
        uint16_t *const sm = R.srow + 8 * h - (CUBIC ? 1 : 0);   // regular main-phase target t at sm[t]
#pragma unroll
        for (int t = 0; t < 8; t++) {
            uint32_t i0, i1;
            const bool m0 = box_class<CUBIC>(t, n_odd, i0), m1 = box_class<CUBIC>(8 + t, n_odd, i1);
            const bool in_main = h ? m1 : m0;
            const uint32_t idx = h ? i1 : i0;
            const uint16_t qv = static_cast<uint16_t>(H.q[t]);
            if (m0 && m1) {
                sm[t] = qv;
            } else {
                if (in_main) sm[t] = qv; else R.qb[idx * R.other] = qv;
            }
        }
        H.finish(ctx, true, [&](int t) {
            // rare: the original of an unpredictable point, re-read from the input at local (z, y, 2k + 1)
            const uint32_t k = 8 * h + t;
            const uint64_t g = T.sbase + z * S.st[0] + (ry + T.low[1]) * S.st[1] + (2 * k + 1) * S.st[2];
            uint32_t idx;
            if (box_class_rt<CUBIC>(k, n_odd, idx))
                R.um[idx] = S.p[g];
            else
                R.ub[idx * R.other] = S.p[g];
        });
    }
}

// a 33rd owned row (tiles at y = 0 with 33 points): lane = target, values from the slot
template <bool CUBIC, class Ctx>
SZ_HD void box_pass2_left(const BoxArgs &A, Ctx &ctx, const BoxTile &T, uint32_t lane, uint32_t z, float *slot,
                          uint16_t *stage, bool write2) {
    if (T.c1[1] != 33 || lane >= 16) return;
    const bool n_odd = T.n[2] & 1u;
    BoxRowOut R;
    box_row_out(A, T, 32u, z, stage, R);   // low == 0 here: row 32 is local y = 32
    float *row = slot + 32 * kBoxPitch;
    box_target_rt<CUBIC>(
        lane, n_odd, A.qp, [&](uint32_t l) { return row[l]; },
        [&](uint32_t l, float r) {
            if (write2) row[l] = r;
        },
        [&](uint32_t kk, int qv, float orig) {
            uint32_t idx;
            const bool in_main = box_class_rt<CUBIC>(kk, n_odd, idx);
            box_row_store(R, in_main, idx, qv, orig);
        });
}

// Level stride >= 2: every owned point of the finished plane goes to recon2, where the next finer level reads it as
// a coarse point (rows are contiguous there at stride s / 2; lane = x, walking over the rows).
SZ_HD void box_plane_out(const BoxArgs &A, const BoxTile &T, uint32_t lane, uint32_t z, const float *slot) {
    const uint64_t h0 = static_cast<uint64_t>(A.s >> 1) * A.stride2[0], h1 = static_cast<uint64_t>(A.s >> 1) * A.stride2[1];
    const uint32_t h2 = A.s >> 1;
    float *const base = A.recon2 + T.g2base + z * h0;
    const uint32_t ny = T.n[1], nx = T.n[2], lowy = T.low[1], lowx = T.low[2];
    if (lane >= lowx && lane < nx) {
        float *p = base + lowy * h1 + lane * h2;
        const float *sl = slot + lowy * kBoxPitch + lane;
        for (uint32_t y = lowy; y < ny; y++) {
            *p = *sl;
            p += h1;
            sl += kBoxPitch;
        }
    }
    if (nx == 33) {
        for (uint32_t y = lowy + lane; y < ny; y += 32) base[y * h1 + 32 * h2] = slot[y * kBoxPitch + 32];
    }
}

// staging buffer -> index stream: the plane's run [pos, pos + len) starts `mis` indices into a 16-byte chunk, and the
// staging buffer is laid out with the same misalignment, so whole chunks move as 16-byte loads / stores
SZ_HD void box_copy_out(const BoxArgs &A, const BoxTile &T, uint32_t lane, uint32_t z, const uint16_t *stage) {
    const uint64_t pos = box_run_pos(T, z);
    const uint32_t len = T.c1[1] * T.mainc[2];
    const uint32_t mis = static_cast<uint32_t>(pos) & 7u;
    uint16_t *const g0 = A.q + (pos - mis);     // 16-byte aligned (A.q is)
    const uint32_t end = mis + len;
    const uint32_t nchunks = (end + 7) / 8;
    for (uint32_t c = lane; c < nchunks; c += 32) {
        const uint32_t a = c * 8, b = a + 8;
        if (a >= mis && b <= end) {
            *reinterpret_cast<BoxChunk *>(g0 + a) = *reinterpret_cast<const BoxChunk *>(stage + a);
        } else {
            for (uint32_t e = a < mis ? mis : a; e < (b < end ? b : end); e++) g0[e] = stage[e];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// recover, last pass (along x) of the finest level.  The box schedule as a whole does not carry over to
// decompression (a tile would need the reconstructions of its low faces from its neighbours), but its last phase
// does: once passes 0 and 1 of the level are complete everywhere, the targets of a row depend on that row's even-x
// points alone.  Same mapping as pass 2 above -- warp per plane, the plane by one TMA box copy (from the output array
// itself), lane = row, the row in registers, the plane's run of main-phase indices through the staging buffer -- with
// LinearQuantizer::recover (LinearQuantizer.hpp:74-86) instead of the quantizer.
// ---------------------------------------------------------------------------------------------------------------------
// index stream -> staging buffer: the mirror image of box_copy_out
SZ_HD void box_copy_in(const BoxArgs &A, const BoxTile &T, uint32_t lane, uint32_t z, uint16_t *stage) {
    const uint64_t pos = box_run_pos(T, z);
    const uint32_t len = T.c1[1] * T.mainc[2];
    const uint32_t mis = static_cast<uint32_t>(pos) & 7u;
    const uint16_t *const g0 = A.q + (pos - mis);   // 16-byte aligned (A.q is)
    const uint32_t end = mis + len;
    const uint32_t nchunks = (end + 7) / 8;
    for (uint32_t c = lane; c < nchunks; c += 32) {
        const uint32_t a = c * 8, b = a + 8;
        if (a >= mis && b <= end) {
            *reinterpret_cast<BoxChunk *>(stage + a) = *reinterpret_cast<const BoxChunk *>(g0 + a);
        } else {   // ragged ends: element by element (the chunk may reach outside the stream)
            for (uint32_t e = a < mis ? mis : a; e < (b < end ? b : end); e++) stage[e] = g0[e];
        }
    }
}

// One row (lane = row): v = the row's 36 floats (even x valid), results to slot_row[odd x].
template <bool CUBIC>
SZ_HD void box_recover_row(const BoxArgs &A, const BoxTile &T, uint32_t ry, uint32_t z, const float *v, const uint16_t *stage,
                           float *slot_row) {
    const bool n_odd = T.n[2] & 1u;
    BoxRowOut R;
    box_row_out(A, T, ry, z, const_cast<uint16_t *>(stage), R);
    const QuantParams qp = A.qp;
#pragma unroll 1
    for (uint32_t h = 0; h < 2; h++) {
        float nb[11];
        nb[0] = h ? v[14] : 0.0f;
#pragma unroll
        for (int m = 1; m < 9; m++) nb[m] = h ? v[14 + 2 * m] : v[2 * m - 2];
        nb[9] = h ? (n_odd ? v[32] : 0.0f) : v[16];
        nb[10] = h ? 0.0f : v[18];
        // indices of the eight targets: main sub-phase from the staging buffer, boundary sub-phases from the stream
        int qv[8];
        const uint16_t *const sm = R.srow + 8 * h - (CUBIC ? 1 : 0);
#pragma unroll
        for (int t = 0; t < 8; t++) {
            uint32_t idx;
            const bool in_main = box_class_rt<CUBIC>(8 * h + t, n_odd, idx);
            qv[t] = in_main ? sm[t] : R.qb[idx * R.other];
        }
        float pred[8];
        if (CUBIC) {
#pragma unroll
            for (int t = 1; t <= 5; t++) pred[t] = interp_cubic<float>(nb[t], nb[t + 1], nb[t + 2], nb[t + 3]);
            if (h == 0) {
                pred[0] = interp_quad_1<float>(nb[1], nb[2], nb[3]);
                pred[6] = interp_cubic<float>(nb[6], nb[7], nb[8], nb[9]);
                pred[7] = interp_cubic<float>(nb[7], nb[8], nb[9], nb[10]);
            } else {
                pred[0] = interp_cubic<float>(nb[0], nb[1], nb[2], nb[3]);
                if (n_odd) {
                    pred[6] = interp_cubic<float>(nb[6], nb[7], nb[8], nb[9]);
                    pred[7] = interp_quad_2<float>(nb[7], nb[8], nb[9]);
                } else {
                    pred[6] = interp_quad_2<float>(nb[6], nb[7], nb[8]);
                    pred[7] = interp_linear1<float>(nb[7], nb[8]);
                }
            }
        } else {
#pragma unroll
            for (int t = 0; t < 8; t++) pred[t] = interp_linear<float>(nb[t + 1], nb[t + 2]);
        }
        const bool lin_tail = !CUBIC && h == 1 && !n_odd;
        float rc[8];
#pragma unroll
        for (int t = 0; t < 8; t++) {
            float p = pred[t];
            if (!CUBIC && t == 7 && lin_tail) p = interp_linear1<float>(rc[6], nb[8]);
            float r = recover_pred<float>(p, qv[t], qp);
            if (qv[t] == 0) {   // unpredictable: the stored value, at the index's own position
                uint32_t idx;
                const bool in_main = box_class_rt<CUBIC>(8 * h + t, n_odd, idx);
                r = in_main ? R.um[idx] : R.ub[idx * R.other];
            }
            rc[t] = r;
            slot_row[16 * h + 2 * t + 1] = r;
        }
    }
}

// a 33rd owned row (tiles at y = 0 with 33 points): lane = target
template <bool CUBIC>
SZ_HD void box_recover_left(const BoxArgs &A, const BoxTile &T, uint32_t lane, uint32_t z, float *slot, const uint16_t *stage) {
    if (T.c1[1] != 33 || lane >= 16) return;
    const bool n_odd = T.n[2] & 1u;
    BoxRowOut R;
    box_row_out(A, T, 32u, z, const_cast<uint16_t *>(stage), R);
    float *row = slot + 32 * kBoxPitch;
    const bool tail_pair = !CUBIC && !n_odd;
    if (tail_pair && lane == 15) return;   // done together with its predecessor
    auto get = [&](uint32_t k, int *qv, float *un) {
        uint32_t idx;
        const bool in_main = box_class_rt<CUBIC>(k, n_odd, idx);
        *qv = in_main ? R.srow[idx] : R.qb[idx * R.other];
        *un = 0.0f;
        if (*qv == 0) *un = in_main ? R.um[idx] : R.ub[idx * R.other];
    };
    const uint32_t k = lane;
    float pred;
    if (CUBIC) {
        const uint32_t kind = k == 0 ? BOX_ST_QUAD1
                                     : (k <= 13 ? BOX_ST_CUBIC : (k == 14 ? (n_odd ? BOX_ST_CUBIC : BOX_ST_QUAD2) : (n_odd ? BOX_ST_QUAD2 : BOX_ST_LINEAR1)));
        const float w0 = k >= 1 ? row[2 * k - 2] : 0.0f;
        const float w1 = row[2 * k];
        const float w2 = kind != BOX_ST_LINEAR1 ? row[2 * k + 2] : 0.0f;
        const float w3 = (kind == BOX_ST_CUBIC || kind == BOX_ST_QUAD1) ? row[2 * k + 4] : 0.0f;
        pred = box_stencil4(kind, w0, w1, w2, w3);
    } else {
        pred = interp_linear<float>(row[2 * k], row[2 * k + 2]);
    }
    int qv;
    float un;
    get(k, &qv, &un);
    float r = qv ? recover_pred<float>(pred, qv, A.qp) : un;
    row[2 * k + 1] = r;
    if (tail_pair && k == 14) {
        pred = interp_linear1<float>(r, row[30]);
        get(15, &qv, &un);
        row[31] = qv ? recover_pred<float>(pred, qv, A.qp) : un;
    }
}

// the plane's owned odd-x points -> the output array (rows contiguous; lane = x, walking over the rows)
SZ_HD void box_recover_out(const BoxSrc &S, const BoxTile &T, uint32_t lane, uint32_t z, const float *slot, float *out) {
    float *const base = out + T.sbase + z * S.st[0];
    const uint32_t ny = T.n[1], nx = T.n[2], lowy = T.low[1];
    if ((lane & 1u) && lane < nx) {
        float *p = base + lowy * S.st[1] + lane;
        const float *sl = slot + lowy * kBoxPitch + lane;
        for (uint32_t y = lowy; y < ny; y++) {
            *p = *sl;
            p += S.st[1];
            sl += kBoxPitch;
        }
    }
}

}  // namespace sz3b
