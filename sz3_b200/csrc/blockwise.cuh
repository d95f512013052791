// sz3_b200/csrc/blockwise.cuh -- geometry of the block walk of BlockwiseDecomposition (reference
// include/SZ3/utils/BlockwiseIterator.hpp:48-56,62-70,103-141): blocks of blockSize^N in row-major block order,
// clipped at the upper faces, elements row-major inside a block.
#pragma once
#include "core.cuh"

namespace sz3b {

struct BlockShape {
    int N;
    uint32_t B;                 // blockSize
    uint32_t dims[kMaxDim];
    uint64_t stride[kMaxDim];   // element strides of the array
    uint32_t nb[kMaxDim];       // blocks per dimension
    uint64_t nblocks;
    uint64_t num;
};

SZ_HD void block_shape_init(BlockShape &bs, int N, const uint64_t *dims, uint32_t B) {
    bs.N = N;
    bs.B = B;
    bs.nblocks = 1;
    bs.num = 1;
    for (int d = 0; d < kMaxDim; d++) {
        bs.dims[d] = d < N ? static_cast<uint32_t>(dims[d]) : 1;
        bs.nb[d] = d < N ? (bs.dims[d] + B - 1) / B : 1;
        bs.stride[d] = 0;
    }
    uint64_t acc = 1;
    for (int d = N - 1; d >= 0; d--) {
        bs.stride[d] = acc;
        acc *= bs.dims[d];
        bs.nblocks *= bs.nb[d];
    }
    bs.num = acc;
}

// block b (row-major over the block grid) -> block index, clipped extents, element offset of its first element
SZ_HD void block_decode(const BlockShape &bs, uint64_t b, uint32_t bi[kMaxDim], uint32_t ext[kMaxDim], uint64_t *off0) {
    uint64_t off = 0;
    for (int d = kMaxDim - 1; d >= 0; d--) {
        if (d >= bs.N) {
            bi[d] = 0;
            ext[d] = 1;
            continue;
        }
        bi[d] = static_cast<uint32_t>(b % bs.nb[d]);
        b /= bs.nb[d];
        const uint32_t lo = bi[d] * bs.B;
        ext[d] = bs.dims[d] - lo < bs.B ? bs.dims[d] - lo : bs.B;
        off += static_cast<uint64_t>(lo) * bs.stride[d];
    }
    *off0 = off;
}

// RegressionPredictor::precompress (RegressionPredictor.hpp:28-55) for block b.  Returns false when the block has an
// extent <= 1 (the reference then falls back to Lorenzo).  Sequential row-major sums on purpose: for T = double the
// accumulation order decides the coefficient bits.
// index[i] * (*c) of the fit (RegressionPredictor.hpp:44): a size_t times a T -- a product in T for floating T; for
// integer T the value is converted to size_t first, so the product is an unsigned 64-bit one (negative values wrap;
// reproduced as is), then widened to double for the sum.
template <class T>
SZ_HD double reg_index_times(uint32_t i, T v) {
    if constexpr (std::is_integral<T>::value)
        return static_cast<double>(static_cast<uint64_t>(i) * static_cast<uint64_t>(static_cast<int64_t>(v)));
    else
        return static_cast<double>(static_cast<T>(i) * v);
}

template <class T>
SZ_HD bool reg_fit_block(const T *data, const BlockShape &bs, uint64_t b, T coef[kMaxDim + 1]) {
    uint32_t bi[kMaxDim], ext[kMaxDim];
    uint64_t off0;
    block_decode(bs, b, bi, ext, &off0);
    for (int d = 0; d < bs.N; d++)
        if (ext[d] <= 1) return false;
    double sum[kMaxDim + 1] = {0, 0, 0, 0, 0};
    const uint32_t e0 = ext[0], e1 = bs.N > 1 ? ext[1] : 1, e2 = bs.N > 2 ? ext[2] : 1, e3 = bs.N > 3 ? ext[3] : 1;
    const uint64_t s0 = bs.stride[0], s1 = bs.N > 1 ? bs.stride[1] : 0, s2 = bs.N > 2 ? bs.stride[2] : 0,
                   s3 = bs.N > 3 ? bs.stride[3] : 0;
    for (uint32_t i0 = 0; i0 < e0; i0++)
        for (uint32_t i1 = 0; i1 < e1; i1++)
            for (uint32_t i2 = 0; i2 < e2; i2++)
                for (uint32_t i3 = 0; i3 < e3; i3++) {
                    const T v = data[off0 + i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3];
                    sum[0] += reg_index_times<T>(i0, v);   // index * value in T, accumulated in double
                    if (bs.N > 1) sum[1] += reg_index_times<T>(i1, v);
                    if (bs.N > 2) sum[2] += reg_index_times<T>(i2, v);
                    if (bs.N > 3) sum[3] += reg_index_times<T>(i3, v);
                    sum[bs.N] += static_cast<double>(v);
                }
    double num = 1;
    for (int d = 0; d < bs.N; d++) num *= static_cast<double>(ext[d]);
    coef[bs.N] = from_double<T>(sum[bs.N] / num);
    for (int d = 0; d < bs.N; d++) {
        const double dd = static_cast<double>(ext[d]);
        coef[d] = from_double<T>((2 * sum[d] / (dd - 1) - sum[bs.N]) * 6 / num / (dd + 1));
        coef[bs.N] = from_double<T>(static_cast<double>(coef[bs.N]) - (dd - 1) * static_cast<double>(coef[d]) / 2);
    }
    return true;
}

// Element gid (memory order): its block, in-block index and block-major traversal position.
SZ_HD void reg_locate(const BlockShape &bs, uint64_t gid, uint64_t *blin_out, uint32_t li[kMaxDim], uint64_t *pos_out) {
    uint32_t x[kMaxDim] = {0, 0, 0, 0};
    uint64_t r = gid;
    for (int d = bs.N - 1; d >= 0; d--) {
        x[d] = static_cast<uint32_t>(r % bs.dims[d]);
        r /= bs.dims[d];
    }
    // pos = sum_d (prod_{e<d} ext_e) * (b_d*B) * (prod_{e>d} dims_e)  +  row-major index inside the block
    uint64_t blin = 0, pos = 0, within = 0, extprod = 1, tail = bs.num;
    for (int d = 0; d < bs.N; d++) {
        const uint32_t bd = x[d] / bs.B;
        const uint32_t lo = bd * bs.B;
        const uint32_t ext = bs.dims[d] - lo < bs.B ? bs.dims[d] - lo : bs.B;
        li[d] = x[d] - lo;
        blin = blin * bs.nb[d] + bd;
        tail /= bs.dims[d];
        pos += extprod * lo * tail;
        within = within * ext + li[d];
        extprod *= ext;
    }
    *blin_out = blin;
    *pos_out = pos + within;
}

// Row-wise form of reg_locate for the predict kernel: everything that depends only on the row (all coordinates but the
// fastest one) is folded once per row; an element then costs one small division.
struct RegRow {
    uint64_t pos_base;     // traversal position contributed by the slower dims (block offsets + in-block row rank)
    uint64_t blin_base;    // block index contributed by the slower dims (times nb of the fastest dim)
    uint64_t extprod;      // product of this row's block extents over the slower dims
    uint64_t within_row;   // row-major rank of the row inside its block (over the slower dims)
    uint32_t li[kMaxDim];  // in-block indices of the slower dims
};

// x[0..N-2]: coordinates of the row in the slower dims
SZ_HD void reg_row_setup(const BlockShape &bs, const uint32_t x[kMaxDim], RegRow &rr) {
    uint64_t blin = 0, pos = 0, within = 0, extprod = 1, tail = bs.num;
    for (int d = 0; d < bs.N - 1; d++) {
        const uint32_t bd = x[d] / bs.B;
        const uint32_t lo = bd * bs.B;
        const uint32_t ext = bs.dims[d] - lo < bs.B ? bs.dims[d] - lo : bs.B;
        rr.li[d] = x[d] - lo;
        blin = blin * bs.nb[d] + bd;
        tail /= bs.dims[d];
        pos += extprod * lo * tail;
        within = within * ext + rr.li[d];
        extprod *= ext;
    }
    rr.pos_base = pos;
    rr.blin_base = blin * bs.nb[bs.N - 1];
    rr.extprod = extprod;
    rr.within_row = within;
}

// element x of the row: block index, in-block index along the fastest dim, traversal position
SZ_HD void reg_row_locate(const BlockShape &bs, const RegRow &rr, uint32_t x, uint32_t mgB, uint64_t *blin,
                          uint32_t *lx, uint64_t *pos) {
    const int L = bs.N - 1;
    // x / B: multiply-high by ceil(2^32 / B) when the row is short enough for that to be exact, else a division
    const uint32_t bx = mgB ? static_cast<uint32_t>((static_cast<uint64_t>(x) * mgB) >> 32) : x / bs.B;
    const uint32_t lo = bx * bs.B;
    const uint32_t ext = bs.dims[L] - lo < bs.B ? bs.dims[L] - lo : bs.B;
    *lx = x - lo;
    *blin = rr.blin_base + bx;
    *pos = rr.pos_base + rr.extprod * lo + rr.within_row * ext + (x - lo);
}

// RegressionPredictor::predict (RegressionPredictor.hpp:77-91), T arithmetic, left to right
template <class T>
SZ_HD T reg_predict(int N, const T *c, const uint32_t li[kMaxDim]) {
    using U = typename Arith<T>::U;   // (integer T: size_t arithmetic in the reference, i.e. wrapping)
    const U c0 = static_cast<U>(c[0]), c1 = static_cast<U>(c[1]);
    if (N == 1) return static_cast<T>(c0 * static_cast<U>(li[0]) + c1);
    const U c2 = static_cast<U>(c[2]);
    if (N == 2) return static_cast<T>(c0 * static_cast<U>(li[0]) + c1 * static_cast<U>(li[1]) + c2);
    const U c3 = static_cast<U>(c[3]);
    if (N == 3) return static_cast<T>(c0 * static_cast<U>(li[0]) + c1 * static_cast<U>(li[1]) + c2 * static_cast<U>(li[2]) + c3);
    return static_cast<T>(c0 * static_cast<U>(li[0]) + c1 * static_cast<U>(li[1]) + c2 * static_cast<U>(li[2]) +
                          c3 * static_cast<U>(li[3]) + static_cast<U>(c[4]));
}

}  // namespace sz3b
