// sz3_b200/csrc/interp_plan.hpp -- host-side plan of one interpolation decomposition: level list, per-level error
// bounds and the table of traversal positions at which every level block starts emitting.
//
// Restates the control flow of InterpolationDecomposition::init / compress (reference
// include/SZ3/decomposition/InterpolationDecomposition.hpp:79-147,176-213); the per-point work is in the kernels.
#pragma once
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/sz3b.h"
#include "core.cuh"

namespace sz3b {

struct LevelPlan {
    int level;
    uint32_t s;
    uint32_t nb[kMaxDim];
    uint64_t nblocks;
    double eb;
    uint64_t table_off;   // offset of this level's block_base table inside InterpPlan::table
};

struct InterpPlan {
    InterpShape sh;
    uint64_t num = 0;             // elements of one array
    uint32_t anchor_stride = 0;   // after init(): 0 when no anchor grid is used
    uint64_t n_first = 0;         // indices emitted before the first level (anchors, or the single first element)
    std::vector<LevelPlan> levels;
    std::vector<uint64_t> table;
    uint32_t dims2[kMaxDim];
    uint64_t stride2[kMaxDim];
    uint64_t num2 = 0;
    bool tile = false;            // tile schedule (N == 3) or generic per-pass schedule
    bool lean = false;            // per-pass schedule through the row-mapped kernel (interp_lean.cuh, N >= 3)
    bool box_required = false;    // schedule 6: fail instead of falling back
    bool box = false;             // finest level through the box schedule (interp_box.cuh) where it applies; schedule 6 insists
    int interp_id = 1, direction = 0;
    double alpha = 1.25, beta = 2.0;
    double eb = 0;
};

inline int ceil_log2_u64(uint64_t x) {
    int k = 0;
    while ((1ull << k) < x) k++;
    return k;
}

// Returns nullptr on success, else an error message (invalid argument).
inline const char *build_interp_plan(const sz3b_config &c, double eb, int schedule, InterpPlan &pl) {
    const int N = c.N;
    if (N < 1 || N > 4) return "Data dimension higher than 4 is not supported.";
    InterpShape &sh = pl.sh;
    sh.N = N;
    pl.num = 1;
    for (int d = 0; d < kMaxDim; d++) {
        sh.dims[d] = d < N ? static_cast<uint32_t>(c.dims[d]) : 1;
        sh.stride[d] = 0;
        sh.perm[d] = d;
    }
    for (int d = 0; d < N; d++) {
        if (c.dims[d] == 0 || c.dims[d] > 0xffffffffull) return "dimension out of range";
        pl.num *= c.dims[d];
    }
    uint64_t acc = 1;
    for (int d = N - 1; d >= 0; d--) {
        sh.stride[d] = acc;
        acc *= sh.dims[d];
    }
    // dim_sequences[direction]: direction-th lexicographic permutation (:205-212)
    int fact = 1;
    for (int i = 2; i <= N; i++) fact *= i;
    if (c.interpDirection < 0 || c.interpDirection >= fact) return "interpDirection out of range";
    {
        int perm[kMaxDim] = {0, 1, 2, 3};
        for (int k = 0; k < c.interpDirection; k++) std::next_permutation(perm, perm + N);
        for (int d = 0; d < N; d++) sh.perm[d] = perm[d];
    }
    if (c.interpAlgo != SZ3B_INTERP_LINEAR && c.interpAlgo != SZ3B_INTERP_CUBIC) return "unknown interpAlgo";
    sh.cubic = c.interpAlgo == SZ3B_INTERP_CUBIC;
    sh.old_api = N <= 2;
    pl.interp_id = c.interpAlgo;
    pl.direction = c.interpDirection;
    pl.alpha = c.interpAlpha;
    pl.beta = c.interpBeta;
    pl.eb = eb;

    int64_t astride = c.interpAnchorStride;
    if (astride < 0) {
        static const int def[4] = {4096, 128, 32, 16};  // SZAlgoInterp.hpp:21-24
        astride = def[N - 1];
    }
    if (astride & (astride - 1)) return "Anchor stride should be 0 or 2's exponentials";
    int level = -1;
    bool use_anchor = false;
    for (int d = 0; d < N; d++) {
        level = std::max(level, ceil_log2_u64(sh.dims[d]));
        if (sh.dims[d] > static_cast<uint64_t>(astride)) use_anchor = true;
    }
    if (!use_anchor) astride = 0;
    if (astride > 0) {
        int maxl = ceil_log2_u64(static_cast<uint64_t>(astride)) + 1;  // log2 of a power of two
        if (maxl <= level) level = maxl;
    }
    pl.anchor_stride = static_cast<uint32_t>(astride);
    if (astride == 0) {
        pl.n_first = 1;
    } else {
        pl.n_first = 1;
        for (int d = 0; d < N; d++) pl.n_first *= (sh.dims[d] - 1) / astride + 1;
        level--;
    }

    pl.num2 = 1;
    acc = 1;
    for (int d = kMaxDim - 1; d >= 0; d--) {
        pl.dims2[d] = d < N ? (sh.dims[d] - 1) / 2 + 1 : 1;
        pl.stride2[d] = d < N ? acc : 0;
        if (d < N) acc *= pl.dims2[d];
    }
    pl.num2 = acc;
    if (schedule == 2 || schedule == 3) return "tile schedules 2 and 3 (first tile kernels) were retired: use 4 (line walker) or 0";
    pl.tile = (N == 3) && schedule != 1 && schedule != 5;
    // the tile kernels keep tile-relative element offsets in 32 bits: larger arrays take the row-mapped per-pass kernels
    if (pl.tile && pl.num >= (1ull << 32)) {
        if (schedule == 4 || schedule == 6) return "tile schedules need fewer than 2^32 elements";
        pl.tile = false;
    }
    pl.lean = N >= 3 && (schedule == 5 || (schedule == 0 && !pl.tile));
    pl.box = pl.tile && (schedule == 0 || schedule == 6);
    pl.box_required = schedule == 6;
    if ((schedule == 4 || schedule == 6) && N != 3) return "tile schedule needs N == 3";
    if (schedule == 5 && N < 3) return "row-mapped per-pass schedule needs N >= 3";
    if (schedule < 0 || schedule > 6) return "unknown schedule";

    pl.levels.clear();
    pl.table.clear();
    uint64_t pos = pl.n_first;
    for (int lv = level; lv > 0 && lv <= level; lv--) {
        LevelPlan L;
        L.level = lv;
        if (lv - 1 >= 31) return "interpolation level too deep";
        L.s = 1u << (lv - 1);
        double cur = eb;
        if (c.interpAlpha < 0) {
            cur = lv >= 3 ? eb * 0.5 : eb;   // eb_ratio = 0.5 (:468)
        } else if (c.interpAlpha >= 1) {
            double ratio = pow(c.interpAlpha, lv - 1);
            if (ratio > c.interpBeta) ratio = c.interpBeta;
            cur = eb / ratio;
        }
        L.eb = cur;
        const uint64_t B = static_cast<uint64_t>(kInterpBlock) * L.s;
        L.nblocks = 1;
        for (int d = 0; d < kMaxDim; d++) {
            L.nb[d] = d < N ? static_cast<uint32_t>((sh.dims[d] - 1) / B + 1) : 1;
            L.nblocks *= L.nb[d];
        }
        L.table_off = pl.table.size();
        pl.table.resize(pl.table.size() + L.nblocks);
        uint32_t bidx[kMaxDim] = {0, 0, 0, 0};
        for (uint64_t b = 0; b < L.nblocks; b++) {
            pl.table[L.table_off + b] = pos;
            BlockGeom g;
            // dims beyond N have nb == 1 and are ignored by block_geom (loops to sh.N)
            block_geom(sh, L.s, bidx, g);
            PassGeom pg;
            for (int p = 0; p < N; p++) {
                pass_geom(sh, L.s, g, p, pg);
                pos += pg.size;
            }
            for (int d = N - 1; d >= 0; d--) {
                if (++bidx[d] < L.nb[d]) break;
                bidx[d] = 0;
            }
        }
        pl.levels.push_back(L);
    }
    if (pos != pl.num) return "internal error: traversal does not cover the array";
    return nullptr;
}

}  // namespace sz3b
