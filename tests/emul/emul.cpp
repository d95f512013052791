// tests/emul/emul.cpp -- TEST INFRASTRUCTURE: runs the kernel *bodies* of sz3_b200/csrc/interp_body.cuh with host
// threads so that the traversal/indexing logic can be checked against the reference on a machine without a GPU.
// Never linked into the product library (which has no CPU path); built by tests/conftest.py into tests/emul/_build/.
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../../sz3_b200/csrc/interp_body.cuh"
#include "../../sz3_b200/csrc/interp_line.cuh"
#include "../../sz3_b200/csrc/interp_lean.cuh"
#include "../../sz3_b200/csrc/interp_box.cuh"
#include "../../sz3_b200/csrc/interp_plan.hpp"

using namespace sz3b;

struct HostCtx {
    uint32_t t, nt;
    std::barrier<> *bar;
    unsigned long long *hist;
    uint32_t tid() const { return t; }
    uint32_t nthreads() const { return nt; }
    void sync() { bar->arrive_and_wait(); }
    void pass_end() {}
    void hist_add(int sym, bool active) {
        if (active) __atomic_fetch_add(&hist[sym], 1ull, __ATOMIC_RELAXED);
    }
};

// Box schedule (interp_box.cuh): the per-lane phase functions run lane by lane, warp by warp, in the order the
// kernel's barriers impose; a host copy with zero fill stands in for the TMA box (out-of-bounds elements read 0).
struct SeqCtx {};   // the box schedule counts nothing itself (k_hist_u16 runs over its indices afterwards)

template <bool CUBIC>
static void box_tile_emul(const BoxArgs &A, const BoxSrc &S, const uint32_t sdims[3], uint32_t tile, unsigned long long *hist) {
    static std::vector<float> EE(kBoxEEElems), slots(kBoxWarps * kBoxSlotStride);
    static std::vector<uint16_t> stage(kBoxWarps * kBoxStageU16);
    std::fill(EE.begin(), EE.end(), std::numeric_limits<float>::quiet_NaN());   // unfilled cells must never matter
    SeqCtx ctx;
    (void)hist;
    BoxOrigin o;
    box_origin(A, S, tile, o);
    BoxTile T;
    box_tile_setup<CUBIC>(A, tile, o, T);
    const bool write2 = A.s >= 2;
    for (uint32_t t = 0; t < kBoxThreads; t++) {
        box_fill_column(A, S, o, t, EE.data());
        if (t < 33) box_fill_column(A, S, o, 256 + t, EE.data());
    }
    for (uint32_t t = kBoxThreads; t-- > 0;) box_pass0_line<CUBIC>(A, S, ctx, T, t, EE.data(), true);
    for (uint32_t t = kBoxThreads; t-- > 0;)
        for (uint32_t e = t; e < 33 * 16; e += kBoxThreads) box_pass0_left<CUBIC>(A, ctx, T, e, EE.data(), true);
    for (uint32_t w = kBoxWarps; w-- > 0;) {
        float *slot = slots.data() + w * kBoxSlotStride;
        uint16_t *stg = stage.data() + w * kBoxStageU16;
        for (uint32_t z = T.low[0] + w; z < T.n[0]; z += kBoxWarps) {
            std::fill(slot, slot + kBoxSlotStride, std::numeric_limits<float>::quiet_NaN());
            if (S.tma) {
                // TMA box (36, 33, 1) at (x0, y0, z0 + z) of the dense source array: out-of-bounds elements read 0
                for (uint32_t y = 0; y < 33; y++)
                    for (uint32_t x = 0; x < 36; x++) {
                        const uint64_t gz = o.begin[0] / S.odiv + z, gy = o.begin[1] / S.odiv + y, gx = o.begin[2] / S.odiv + x;
                        const bool in = gz < sdims[0] && gy < sdims[1] && gx < sdims[2];
                        slot[y * kBoxPitch + x] = in ? S.p[gz * S.st[0] + gy * S.st[1] + gx] : 0.0f;
                    }
            } else {
                for (uint32_t l = 32; l-- > 0;) box_gather_plane(S, T, l, z, slot);
            }
            const float *EEz = EE.data() + z * kBoxEEPlane;
            for (uint32_t l = 32; l-- > 0;) box_merge(T, l, EEz, slot);
            for (uint32_t l = 32; l-- > 0;) box_pass1_lane<CUBIC>(A, S, ctx, T, l, z, EEz, slot);
            for (uint32_t l = 32; l-- > 0;) box_pass1_left<CUBIC>(A, ctx, T, l, z, EEz, slot);
            std::fill(stg, stg + kBoxStageU16, static_cast<uint16_t>(0xdead));
            float rows[32][36];
            for (uint32_t l = 0; l < 32; l++)
                if (l < T.c1[1]) memcpy(rows[l], slot + (l + T.low[1]) * kBoxPitch, sizeof(rows[l]));
            for (uint32_t l = 32; l-- > 0;) box_pass2_left<CUBIC>(A, ctx, T, l, z, slot, stg, write2);
            for (uint32_t l = 32; l-- > 0;) {
                if (l >= T.c1[1]) continue;
                box_pass2_row<CUBIC>(A, S, ctx, T, l, z, rows[l], stg, write2 ? slot + (l + T.low[1]) * kBoxPitch : nullptr);
            }
            for (uint32_t l = 32; l-- > 0;) box_copy_out(A, T, l, z, stg);
            if (write2)
                for (uint32_t l = 32; l-- > 0;) box_plane_out(A, T, l, z, slot);
        }
    }
}

// mirrors pipeline.cu: launch_box (source of a level) with host-made compact lattices
struct BoxEmulPlan {
    std::vector<float> compact[3];
};
static bool box_level_ok(const BoxArgs &A) {
    if (A.sh.N != 3 || A.sh.perm[0] != 0 || A.sh.perm[1] != 1 || A.sh.perm[2] != 2) return false;
    for (int d = 0; d < 3; d++) {
        const uint32_t B = kInterpBlock * A.s;
        const uint32_t last_begin = ((A.sh.dims[d] - 1) / B) * B;
        const uint32_t n_last = (A.sh.dims[d] - 1 - last_begin) / A.s + 1;
        if (n_last != 32 && n_last != 33) return false;
    }
    return true;
}
static void box_level_emul(const BoxArgs &A, BoxEmulPlan &bp, uint64_t ntiles, unsigned long long *hist) {
    BoxSrc S;
    uint32_t sdims[3];
    int k = -1;
    for (int j = 0; j < 3; j++)
        if (A.s == (4u << j)) k = j;
    if (k >= 0) {
        for (int d = 0; d < 3; d++) sdims[d] = (A.sh.dims[d] - 1) / A.s + 1;
        std::vector<float> &c = bp.compact[k];
        c.resize(static_cast<size_t>(sdims[0]) * sdims[1] * sdims[2]);
        for (uint32_t z = 0; z < sdims[0]; z++)
            for (uint32_t y = 0; y < sdims[1]; y++)
                for (uint32_t x = 0; x < sdims[2]; x++)
                    c[(static_cast<size_t>(z) * sdims[1] + y) * sdims[2] + x] =
                        A.data[(z * A.sh.stride[0] + y * A.sh.stride[1] + x) * A.s];
        S.p = c.data();
        S.st[2] = 1;
        S.st[1] = sdims[2];
        S.st[0] = static_cast<uint64_t>(sdims[1]) * sdims[2];
        for (int d = 0; d < 3; d++) S.ost[d] = S.st[d];
        S.odiv = A.s;
        S.tma = (sdims[2] & 3u) == 0;
    } else {
        for (int d = 0; d < 3; d++) {
            sdims[d] = A.sh.dims[d];
            S.ost[d] = A.sh.stride[d];
            S.st[d] = A.sh.stride[d] * A.s;
        }
        S.p = A.data;
        S.odiv = 1;
        S.tma = A.s == 1 && (sdims[2] & 3u) == 0;
    }
    for (uint64_t tile = 0; tile < ntiles; tile++) {
        if (A.sh.cubic) box_tile_emul<true>(A, S, sdims, static_cast<uint32_t>(tile), hist);
        else box_tile_emul<false>(A, S, sdims, static_cast<uint32_t>(tile), hist);
    }
}

template <class T, class QT>
static int run(const sz3b_config &c, double eb, const T *data, int schedule, int nthreads, int32_t *quant_out,
               T *unpred_out, size_t *n_unpred, unsigned long long *hist_out) {
    InterpPlan pl;
    const bool box_schedule = schedule == 6;
    if (box_schedule) schedule = 4;
    if (const char *e = build_interp_plan(c, eb, schedule, pl)) {
        fprintf(stderr, "[emul] plan: %s\n", e);
        return -1;
    }
    const int radius = c.quantbinCnt / 2;
    const uint64_t n = pl.num;
    std::vector<QT> q(n + 8, static_cast<QT>(0xFFFF));
    std::vector<T> unpred_tmp(n);
    std::vector<T> recon(pl.tile ? pl.num2 : pl.num);
    std::vector<unsigned long long> hist(2 * radius, 0);
    InterpArgs<T, QT> A;
    memset(&A, 0, sizeof(A));
    A.sh = pl.sh;
    A.data = data;
    A.data_bstride = pl.num;
    A.q_bstride = pl.num;
    A.recon2_bstride = pl.num2;
    for (int d = 0; d < kMaxDim; d++) {
        A.dims2[d] = pl.dims2[d];
        A.stride2[d] = pl.stride2[d];
    }
    if (pl.tile) A.recon2 = recon.data(); else A.work = recon.data();
    A.q = q.data();
    A.unpred_tmp = unpred_tmp.data();
    A.hist = hist.data();
    A.qp = make_quant(eb, radius);
    // anchors / first element (same logic as k_interp_anchor)
    if (pl.anchor_stride == 0) {
        T rec;
        int qv = quantize<T>(data[0], static_cast<T>(0), A.qp, rec);
        q[0] = static_cast<QT>(qv);
        if (qv == 0) unpred_tmp[0] = data[0];
        recon[0] = rec;
        hist[qv]++;
    } else {
        for (uint64_t gid = 0; gid < pl.n_first; gid++) {
            uint64_t r = gid, off = 0, off2 = 0;
            for (int d = pl.sh.N - 1; d >= 0; d--) {
                uint32_t ext = (pl.sh.dims[d] - 1) / pl.anchor_stride + 1;
                uint32_t x = static_cast<uint32_t>(r % ext) * pl.anchor_stride;
                r /= ext;
                off += x * pl.sh.stride[d];
                off2 += (x >> 1) * pl.stride2[d];
            }
            q[gid] = 0;
            unpred_tmp[gid] = data[off];
            if (pl.tile) recon[off2] = data[off]; else recon[off] = data[off];
        }
        hist[0] += pl.n_first;
    }
    std::vector<T> sm(kTileSmemElems);
    for (const LevelPlan &L : pl.levels) {
        A.qp = make_quant(L.eb, radius);
        A.s = L.s;
        for (int d = 0; d < kMaxDim; d++) A.nb[d] = L.nb[d];
        A.block_base = pl.table.data() + L.table_off;
        std::barrier<> bar(nthreads);
        // box schedule (schedule 6 of the emulation): every level whose tiles all qualify, the line walker elsewhere
        bool box_level = false;
        if (box_schedule && pl.tile && sizeof(T) == 4 && sizeof(QT) == 2) {
            box_level = box_level_ok(*reinterpret_cast<const BoxArgs *>(&A));
            if (!box_level && L.s == 1) return -7;
        }
        if (box_level) {
            static BoxEmulPlan bplan;
            box_level_emul(*reinterpret_cast<const BoxArgs *>(&A), bplan, L.nblocks, hist.data());
            // what k_hist_u16 does on the device: the level's stretch of the index stream, counted afterwards
            const uint64_t lv_begin = pl.table[L.table_off];
            const uint64_t lv_end = (&L == &pl.levels.back()) ? pl.num : pl.table[(&L + 1)->table_off];
            for (uint64_t i = lv_begin; i < lv_end; i++) hist[q[i]]++;
            continue;
        }
        if (pl.tile) {
            auto worker = [&](int t) {
                HostCtx ctx{static_cast<uint32_t>(t), static_cast<uint32_t>(nthreads), &bar, hist.data()};
                for (uint64_t tile = 0; tile < L.nblocks; tile++) {
                    {
                        static LineTile lt;   // shared by the worker threads like __shared__ memory
                        LineGeom lg;
                        line_geom(A, static_cast<uint32_t>(tile), 0, lg);
                        if (t < 3) line_pass_setup(A, static_cast<uint32_t>(tile), 0, t, static_cast<uint32_t>(nthreads), lt.ps[t]);
                        line_fill(A, ctx, lg, sm.data());
                        ctx.sync();
                        line_tile_passes(A, ctx, sm.data(), lg, lt);
                    }
                    ctx.sync();
                }
            };
            std::vector<std::thread> th;
            for (int t = 1; t < nthreads; t++) th.emplace_back(worker, t);
            worker(0);
            for (auto &x : th) x.join();
        } else if (pl.lean) {
            // row-mapped per-pass schedule: every CTA of the launch grid, one after the other, with host threads
            static LeanShared S;
            for (int p = 0; p < pl.sh.N; p++) {
                if (pass_points(A, p) == 0) continue;
                LeanArgs<T, QT> P;
                P.A = A;
                P.p = p;
                P.write_work = !(L.s == 1 && p == pl.sh.N - 1);
                P.unpred_in = nullptr;
                P.lg_s = 0;
                while ((1u << P.lg_s) < A.s) P.lg_s++;
                P.chunks_per_brow = (lean_max_rows(A, p) + kLeanRows - 1) / kLeanRows;
                uint64_t nbrows = 1;
                for (int d = 0; d < pl.sh.N - 1; d++) nbrows *= A.nb[d];
                const uint32_t row_len = lean_row_len(A, p);
                const uint32_t gy = (row_len + nthreads - 1) / nthreads;
                P.nchunks_L = gy;
                auto worker = [&](int t) {
                    HostCtx ctx{static_cast<uint32_t>(t), static_cast<uint32_t>(nthreads), &bar, hist.data()};
                    for (uint64_t cx = 0; cx < nbrows * P.chunks_per_brow; cx++)
                        for (uint32_t cy = 0; cy < gy; cy++) {
                            lean_cta<T, QT, HostCtx, false>(P, ctx, S, static_cast<uint32_t>(cx), cy, 0);
                            ctx.sync();
                        }
                };
                std::vector<std::thread> th;
                for (int t = 1; t < nthreads; t++) th.emplace_back(worker, t);
                worker(0);
                for (auto &x : th) x.join();
            }
        } else {
            for (int p = 0; p < pl.sh.N; p++) {
                uint64_t total = pass_points(A, p);
                auto worker = [&](int t) {
                    HostCtx ctx{static_cast<uint32_t>(t), static_cast<uint32_t>(nthreads), &bar, hist.data()};
                    // reverse order inside each thread to shake out order dependences
                    for (uint64_t g = total; g-- > 0;)
                        if (g % nthreads == static_cast<uint64_t>(t)) pass_point(A, ctx, p, g, total, 0);
                };
                std::vector<std::thread> th;
                for (int t = 1; t < nthreads; t++) th.emplace_back(worker, t);
                worker(0);
                for (auto &x : th) x.join();
            }
        }
    }
    size_t nu = 0;
    for (uint64_t i = 0; i < n; i++) {
        quant_out[i] = static_cast<int32_t>(q[i]);
        if (q[i] == 0) unpred_out[nu++] = unpred_tmp[i];
    }
    *n_unpred = nu;
    if (hist_out) memcpy(hist_out, hist.data(), sizeof(unsigned long long) * 2 * radius);
    return 0;
}

extern "C" int emul_interp_decompose(int dtype, const sz3b_config *c, double eb, const void *data, int schedule,
                                     int nthreads, int32_t *quant_out, void *unpred_out, size_t *n_unpred,
                                     unsigned long long *hist_out) {
    sz3b_config cc = *c;
    if (cc.interpAnchorStride < 0) {
        static const int def[4] = {4096, 128, 32, 16};
        cc.interpAnchorStride = def[cc.N - 1];
    }
    if (dtype == 0)
        return run<float, uint16_t>(cc, eb, static_cast<const float *>(data), schedule, nthreads, quant_out,
                                    static_cast<float *>(unpred_out), n_unpred, hist_out);
    return run<double, uint16_t>(cc, eb, static_cast<const double *>(data), schedule, nthreads, quant_out,
                                 static_cast<double *>(unpred_out), n_unpred, hist_out);
}

// ---------------------------------------------------------------------------------------------------------------------
// Regression-only BlockwiseDecomposition: the per-thread bodies of sz3_b200/csrc/blockwise.cuh driven serially
// (fit per block -> coefficient chain -> predict+quantize per element).  Outputs: data indices in traversal order,
// coefficient indices, unpredictable data values in traversal order.
// ---------------------------------------------------------------------------------------------------------------------
#include "../../sz3_b200/csrc/blockwise.cuh"

template <class T>
static int run_reg(const sz3b_config &c, double eb, const T *data, int32_t *quant_out, int32_t *coef_q_out,
                   size_t *n_coef, T *unpred_out, size_t *n_unpred) {
    BlockShape bs;
    block_shape_init(bs, c.N, c.dims, static_cast<uint32_t>(c.blockSize));
    const int nc = c.N + 1;
    std::vector<T> c_rec(bs.nblocks * nc, 0);
    QuantParams ql = make_quant(eb / nc / static_cast<unsigned>(c.blockSize), 32768), qi = make_quant(eb / nc, 32768);
    T prev[kMaxDim + 1] = {0, 0, 0, 0, 0};
    size_t k = 0;
    for (uint64_t b = 0; b < bs.nblocks; b++) {
        T coef[kMaxDim + 1];
        if (!reg_fit_block<T>(data, bs, b, coef)) return -2;
        for (int d = 0; d < nc; d++) {
            T rec;
            coef_q_out[k++] = quantize<T>(coef[d], prev[d], d < c.N ? ql : qi, rec);
            prev[d] = rec;
            c_rec[b * nc + d] = rec;
        }
    }
    *n_coef = k;
    QuantParams qp = make_quant(eb, c.quantbinCnt / 2);
    std::vector<T> un(bs.num);
    // the kernel's mapping: one row (all coordinates but the fastest) at a time, reg_row_setup + reg_row_locate
    const uint32_t len = bs.dims[bs.N - 1];
    const uint32_t mgB = (bs.B > 1 && static_cast<uint64_t>(len) * bs.B < (1ull << 32)) ? 0xffffffffu / bs.B + 1u : 0u;
    for (uint64_t row = 0; row < bs.num / len; row++) {
        uint32_t xr[kMaxDim] = {0, 0, 0, 0};
        uint64_t r = row;
        for (int d = bs.N - 2; d >= 0; d--) {
            xr[d] = static_cast<uint32_t>(r % bs.dims[d]);
            r /= bs.dims[d];
        }
        RegRow rr;
        reg_row_setup(bs, xr, rr);
        for (uint32_t x = 0; x < len; x++) {
            const uint64_t gid = row * len + x;
            uint64_t blin, pos, blin2, pos2;
            uint32_t li[kMaxDim] = {rr.li[0], rr.li[1], rr.li[2], rr.li[3]}, li2[kMaxDim];
            reg_row_locate(bs, rr, x, mgB, &blin, &li[bs.N - 1], &pos);
            reg_locate(bs, gid, &blin2, li2, &pos2);
            if (blin != blin2 || pos != pos2) return -3;
            T pred = reg_predict<T>(bs.N, c_rec.data() + blin * nc, li);
            T rec;
            quant_out[pos] = quantize<T>(data[gid], pred, qp, rec);
            un[pos] = data[gid];
        }
    }
    size_t nu = 0;
    for (uint64_t i = 0; i < bs.num; i++)
        if (quant_out[i] == 0) unpred_out[nu++] = un[i];
    *n_unpred = nu;
    return 0;
}

extern "C" int emul_regression_decompose(int dtype, const sz3b_config *c, double eb, const void *data,
                                         int32_t *quant_out, int32_t *coef_q_out, size_t *n_coef, void *unpred_out,
                                         size_t *n_unpred) {
    if (dtype == 0)
        return run_reg<float>(*c, eb, static_cast<const float *>(data), quant_out, coef_q_out, n_coef,
                              static_cast<float *>(unpred_out), n_unpred);
    return run_reg<double>(*c, eb, static_cast<const double *>(data), quant_out, coef_q_out, n_coef,
                           static_cast<double *>(unpred_out), n_unpred);
}

// ---------------------------------------------------------------------------------------------------------------------
// BlockwiseDecomposition with a Lorenzo predictor in the stack (lorenzo.cuh): the block wavefront in front order, the
// selection-guess / exact-chain / exact-pass iteration of pipeline.cu, then (optionally) the decode pass.
// Returns the number of exact passes that were needed (>= 1), negative on error.
// ---------------------------------------------------------------------------------------------------------------------
#include "../../sz3_b200/csrc/lorenzo.cuh"

template <class T>
static void emul_fronts(const BwArgs<T, uint32_t> &A) {
    const BlockShape &bs = A.bs;
    const int N = bs.N;
    std::vector<T> scratch(bw_scratch_elems(bs, A.nk));
    size_t tile_cap = 1;
    for (int d = 0; d < N; d++) tile_cap *= (bs.dims[d] < bs.B ? bs.dims[d] : bs.B) + kBwPad;
    uint64_t grid = 1;
    for (int d = 0; d < N - 1; d++) grid *= bs.nb[d];
    const uint32_t nfronts = bw_num_fronts(bs);
    for (uint32_t f = 0; f < nfronts; f++)
        for (uint64_t cta = grid; cta-- > 0;) {   // reverse order inside a front: blocks of a front are independent
            uint32_t bi[kMaxDim] = {0, 0, 0, 0};
            uint64_t r = cta;
            uint32_t s = 0;
            for (int d = N - 2; d >= 0; d--) {
                bi[d] = static_cast<uint32_t>(r % bs.nb[d]);
                r /= bs.nb[d];
                s += bi[d];
            }
            if (f < s || f - s >= bs.nb[N - 1]) continue;
            bi[N - 1] = f - s;
            const uint64_t b = cta * bs.nb[N - 1] + bi[N - 1];
            if (b < A.b_lo || b >= A.b_hi) continue;
            bw_process_block<T, uint32_t>(A, bi, scratch.data(), scratch.data() + tile_cap, 0, 1);
        }
}

static uint64_t env_u64(const char *name, uint64_t dflt) {
    const char *v = getenv(name);
    return v ? strtoull(v, nullptr, 10) : dflt;
}

// Mirrors run_blockwise_lorenzo (pipeline.cu): returns the number of exact passes + 1000 * walks, negative on error.
template <class T>
static int run_lorenzo(const sz3b_config &c, double eb, const T *data, int32_t *quant_out, uint8_t *sel_out,
                       int32_t *coef_q_out, size_t *n_coef, T *unpred_out, size_t *n_unpred, T *decoded) {
    BlockShape bs;
    block_shape_init(bs, c.N, c.dims, static_cast<uint32_t>(c.blockSize));
    const int N = c.N, nc = N + 1;
    BwArgs<T, uint32_t> A;
    memset(&A, 0, sizeof(A));
    A.bs = bs;
    uint64_t np = 1;
    for (int d = N - 1; d >= 0; d--) {
        A.pstride[d] = np;
        np *= bs.dims[d] + kBwPad;
    }
    A.qp = make_quant(eb, c.quantbinCnt / 2);
    A.noise[0] = lorenzo_noise<T>(N, 1, eb);
    A.noise[1] = lorenzo_noise<T>(N, 2, eb);
    if (c.lorenzo) A.kinds[A.nk++] = PK_LORENZO1;
    if (c.lorenzo2) A.kinds[A.nk++] = PK_LORENZO2;
    if (c.regression) A.kinds[A.nk++] = PK_REG;
    A.b_lo = 0;
    A.b_hi = bs.nblocks;
    std::vector<uint32_t> dtab(65536), didx(65536);
    std::vector<uint16_t> dstart(N * (bs.B - 1) + 2);
    if (!getenv("EMUL_NO_DIAGTAB") && bw_build_diag_table(N, bs.B, dtab.data(), dstart.data(), didx.data())) {
        A.diag_tab = dtab.data();
        A.diag_start = dstart.data();
        A.diag_idx = didx.data();
    }
    const bool has_reg = c.regression != 0 && A.nk > 1;
    const int reg_sid = A.nk - 1;
    std::vector<T> W(np), c_fit(bs.nblocks * nc, 0), c_spec(bs.nblocks * nc, 0), c_rec(bs.nblocks * nc, 0), un(bs.num, 0);
    std::vector<uint8_t> valid(bs.nblocks, 0), selA(bs.nblocks, 0), selB(bs.nblocks, 0);
    std::vector<uint32_t> rank(bs.nblocks + 1, 0), q(bs.num, 0);
    QuantParams ql = make_quant(eb / nc / static_cast<unsigned>(c.blockSize), 32768), qi = make_quant(eb / nc, 32768);
    auto pad = [&](uint64_t b_lo, uint64_t b_hi) {   // k_bw_pad
        if (b_lo == 0 && b_hi == bs.nblocks) std::fill(W.begin(), W.end(), static_cast<T>(0));
        for (uint64_t i = 0; i < bs.num; i++) {
            uint64_t r = i, w = 0, b = 0, bmul = 1;
            for (int d = N - 1; d >= 0; d--) {
                const uint64_t x = r % bs.dims[d];
                r /= bs.dims[d];
                w += (x + kBwPad) * A.pstride[d];
                b += (x / bs.B) * bmul;
                bmul *= bs.nb[d];
            }
            if (b >= b_lo && b < b_hi) W[w] = data[i];
        }
    };
    A.W = W.data();
    A.q = q.data();
    A.unpred_tmp = un.data();
    unsigned mm[2] = {0, ~0u};
    A.mismatch = mm;
    int passes = 0, walks = 0;
    uint32_t nsel_lo = 0;
    const uint64_t min_win = env_u64("EMUL_MINWIN", 4096), walk_below = env_u64("EMUL_WALKBELOW", 24),
                   walk_len = env_u64("EMUL_WALKLEN", 512);
    if (has_reg) {
        for (uint64_t b = 0; b < bs.nblocks; b++) {
            T coef[kMaxDim + 1];
            valid[b] = reg_fit_block<T>(data, bs, b, coef) ? 1 : 0;
            if (valid[b])
                for (int d = 0; d < nc; d++) {
                    c_fit[b * nc + d] = coef[d];
                    c_spec[b * nc + d] = coef_lattice_guess<T>(coef[d], d < N ? ql : qi);
                }
        }
        std::vector<unsigned long long> upos(bs.nblocks * nc + 1);
        std::vector<T> uval(bs.nblocks * nc + 1);
        unsigned long long nuc = 0;
        A.c_fit = c_fit.data();
        A.fit_valid = valid.data();
        A.c_spec = c_spec.data();
        A.c_rec = c_rec.data();
        A.c_rec_out = c_rec.data();
        A.rank = rank.data();
        A.q_liner = ql;
        A.q_indep = qi;
        A.coef_q = coef_q_out;
        A.n_unpred_coef = &nuc;
        A.unpred_pos = upos.data();
        A.unpred_val = uval.data();
        pad(0, bs.nblocks);
        A.mode = BW_SPEC;
        A.sel_out = selA.data();
        emul_fronts(A);
        uint64_t b_lo = 0, win = bs.nblocks, last_adv = bs.nblocks;
        while (b_lo < bs.nblocks) {
            if (last_adv < walk_below) {   // k_bw_serial over the next stretch
                const uint64_t b_hi = std::min<uint64_t>(bs.nblocks, b_lo + walk_len);
                pad(b_lo, b_hi);
                BwSerial<T> st;
                st.nsel = nsel_lo;
                for (int d = 0; d < nc; d++) st.prev[d] = nsel_lo ? c_rec[static_cast<size_t>(nsel_lo - 1) * nc + d] : static_cast<T>(0);
                std::vector<T> scratch(bw_scratch_elems(bs, A.nk));
                size_t tile_cap = 1;
                for (int d = 0; d < N; d++) tile_cap *= (bs.dims[d] < bs.B ? bs.dims[d] : bs.B) + kBwPad;
                A.mode = BW_SERIAL;
                A.sel_in = nullptr;
                A.sel_out = selA.data();
                for (uint64_t b = b_lo; b < b_hi; b++) {
                    uint32_t bi[kMaxDim] = {0, 0, 0, 0};
                    uint64_t r = b;
                    for (int d = N - 1; d >= 0; d--) {
                        bi[d] = static_cast<uint32_t>(r % bs.nb[d]);
                        r /= bs.nb[d];
                    }
                    bw_process_block<T, uint32_t>(A, bi, scratch.data(), scratch.data() + tile_cap, 0, 1, &st);
                }
                nsel_lo = static_cast<uint32_t>(st.nsel);
                b_lo = b_hi;
                last_adv = walk_below;
                win = min_win;
                walks++;
                continue;
            }
            const uint64_t b_hi = std::min<uint64_t>(bs.nblocks, b_lo + win);
            // k_bw_rank + k_bw_gather_fit + k_reg_chain_spec2 (continued chain) over the window's guessed selection
            T prev[kMaxDim + 1];
            for (int d = 0; d < nc; d++) prev[d] = nsel_lo ? c_rec[static_cast<size_t>(nsel_lo - 1) * nc + d] : static_cast<T>(0);
            uint32_t nsel = nsel_lo;
            for (uint64_t b = b_lo; b < b_hi; b++) {
                rank[b] = nsel;
                if (selA[b] != reg_sid) continue;
                for (int d = 0; d < nc; d++) {
                    T rec;
                    coef_q_out[static_cast<size_t>(nsel) * nc + d] = quantize<T>(c_fit[b * nc + d], prev[d], d < N ? ql : qi, rec);
                    prev[d] = rec;
                    c_rec[static_cast<size_t>(nsel) * nc + d] = rec;
                }
                nsel++;
            }
            pad(b_lo, b_hi);
            mm[0] = 0;
            mm[1] = ~0u;
            A.mode = BW_EXACT;
            A.b_lo = b_lo;
            A.b_hi = b_hi;
            A.sel_in = selA.data();
            A.sel_out = selB.data();
            emul_fronts(A);
            passes++;
            uint64_t boundary = b_hi;
            uint32_t nsel_after = nsel;
            if (mm[0]) {
                boundary = mm[1];
                nsel_after = rank[boundary];
                for (uint64_t b = boundary; b < b_hi; b++) selA[b] = selB[b];
            }
            if (getenv("EMUL_VERBOSE"))
                fprintf(stderr, "pass %d: window [%llu, %llu) %u mismatches, final up to %llu of %llu\n", passes,
                        (unsigned long long)b_lo, (unsigned long long)b_hi, mm[0], (unsigned long long)boundary,
                        (unsigned long long)bs.nblocks);
            last_adv = boundary - b_lo;
            b_lo = boundary;
            nsel_lo = nsel_after;
            win = std::max<uint64_t>(min_win, 4 * last_adv);
        }
        A.b_lo = 0;
        A.b_hi = bs.nblocks;
    } else {
        pad(0, bs.nblocks);
        A.mode = BW_EXACT;
        A.sel_in = nullptr;
        A.sel_out = selA.data();
        emul_fronts(A);
        passes = 1;
    }
    *n_coef = static_cast<size_t>(nsel_lo) * nc;
    for (uint64_t i = 0; i < bs.num; i++) quant_out[i] = static_cast<int32_t>(q[i]);
    memcpy(sel_out, selA.data(), bs.nblocks);
    size_t nu = 0;
    for (uint64_t i = 0; i < bs.num; i++)
        if (q[i] == 0) unpred_out[nu++] = un[i];
    *n_unpred = nu;
    if (decoded) {
        std::fill(W.begin(), W.end(), static_cast<T>(0));
        A.mode = BW_DECODE;
        A.sel_in = selA.data();
        if (has_reg) {   // dense chain recover over the final selection (k_reg_chain_recover + k_bw_rank)
            T cur[kMaxDim + 1] = {0, 0, 0, 0, 0};
            uint32_t nsel = 0;
            size_t k = 0;
            for (uint64_t b = 0; b < bs.nblocks; b++) {
                rank[b] = nsel;
                if (selA[b] != reg_sid) continue;
                for (int d = 0; d < nc; d++) {
                    const int qv = coef_q_out[k++];
                    cur[d] = qv ? recover_pred<T>(cur[d], qv, d < N ? ql : qi) : c_fit[b * nc + d];   // 0: stored exactly
                    c_rec[static_cast<size_t>(nsel) * nc + d] = cur[d];
                }
                nsel++;
            }
        }
        A.out = decoded;
        emul_fronts(A);
    }
    return passes + 1000 * walks;
}

extern "C" int emul_lorenzo_decompose(int dtype, const sz3b_config *c, double eb, const void *data, int32_t *quant_out,
                                      uint8_t *sel_out, int32_t *coef_q_out, size_t *n_coef, void *unpred_out,
                                      size_t *n_unpred, void *decoded) {
    if (dtype == 0)
        return run_lorenzo<float>(*c, eb, static_cast<const float *>(data), quant_out, sel_out, coef_q_out, n_coef,
                                  static_cast<float *>(unpred_out), n_unpred, static_cast<float *>(decoded));
    return run_lorenzo<double>(*c, eb, static_cast<const double *>(data), quant_out, sel_out, coef_q_out, n_coef,
                               static_cast<double *>(unpred_out), n_unpred, static_cast<double *>(decoded));
}
