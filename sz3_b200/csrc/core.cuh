// sz3_b200/csrc/core.cuh -- arithmetic and traversal geometry shared by every predict+quantize kernel.
//
// Everything in this header is plain inline code usable from __device__ and from host C++ (the host tail uses the
// same geometry to size buffers and to build the per-level tile tables; tests/emul compiles the kernel bodies with
// g++ to check the indexing logic on a machine without a GPU).  Nothing here is a CPU code path of the product.
//
// Bit-parity rules: no FMA contraction anywhere (nvcc -fmad=false, g++ -ffp-contract=off),
// left-to-right evaluation in T, double only where the reference uses double.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <math.h>
#include <limits>
#include <type_traits>

#if defined(__CUDACC__)
#define SZ_HD __host__ __device__ __forceinline__
#define SZ_D __device__ __forceinline__
#define SZ_NOINLINE __host__ __device__ __noinline__
#else
#define SZ_HD inline
#define SZ_D inline
#define SZ_NOINLINE __attribute__((noinline))
#endif

namespace sz3b {

constexpr int kMaxDim = 4;
constexpr int kInterpBlock = 32;  // reference: InterpolationDecomposition.hpp:85 (blocksize = 32)

// ---------------------------------------------------------------------------------------------------------------------
// LinearQuantizer (reference include/SZ3/quantizer/LinearQuantizer.hpp:43-71, 74-86)
// ---------------------------------------------------------------------------------------------------------------------
struct QuantParams {
    double eb;      // error_bound
    double ebr;     // error_bound_reciprocal = 1.0 / eb (computed on the host exactly like set_eb, :34-37)
    int radius;     // quantbinCnt / 2
    double vmax;    // 2*radius - 1 as double: trunc(v)+1 < 2*radius  <=>  v < 2*radius-1
    float ebf;      // largest float <= eb: for a float x, (double)x <= eb  <=>  x <= ebf
};

SZ_HD QuantParams make_quant(double eb, int radius) {
    QuantParams q;
    q.eb = eb;
    q.ebr = 1.0 / eb;
    q.radius = radius;
    q.vmax = static_cast<double>(2 * static_cast<long long>(radius) - 1);
    float f = static_cast<float>(eb);
    if (static_cast<double>(f) > eb) f = nextafterf(f, -INFINITY);
    q.ebf = f;
    return q;
}

// Exact int <-> double conversions on the FP64 pipe instead of the (16x slower) conversion unit.
SZ_HD int trunc_to_int(double v) {   // 0 <= v < 2^31
#if defined(__CUDA_ARCH__)
    return __double2loint(__dadd_rz(v, 4503599627370496.0));   // 2^52: the integer part lands in the low word
#else
    return v < 2147483647.0 ? static_cast<int>(v) : 0;   // out of range / NaN: value unused by the caller
#endif
}
SZ_HD double int_to_double(int q) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(0x43300000, static_cast<int>(0x80000000u ^ static_cast<unsigned>(q))) - 4503601774854144.0;
#else
    return static_cast<double>(q);
#endif
}

// double -> T as the reference's compiled code does it.  Floating T: a plain conversion.  Integer T (int32_t /
// int64_t, tools/sz3/sz3.cpp:458-461): truncation, and x86's "integer indefinite" (the most negative value) where
// the value does not fit -- CUDA's own conversion would saturate instead.
template <class T>
SZ_HD T from_double(double v) {
    if constexpr (std::is_integral<T>::value) {
        constexpr double lim = sizeof(T) == 4 ? 2147483648.0 : 9223372036854775808.0;
        if (!(v > -lim - (sizeof(T) == 4 ? 1.0 : 0.0) && v < lim)) {
            if (!(sizeof(T) == 8 && v == -lim)) return std::numeric_limits<T>::min();
        }
        return static_cast<T>(v);
    } else {
        return static_cast<T>(v);
    }
}
// wrapping integer arithmetic (what the reference's signed overflow does in practice), plain arithmetic otherwise
template <class T>
struct Arith {
    using U = T;
};
template <>
struct Arith<int32_t> {
    using U = uint32_t;
};
template <>
struct Arith<int64_t> {
    using U = uint64_t;
};

// quantize_and_overwrite.  Returns the shifted index (0 = unpredictable); recon receives the value the reference
// leaves in the working array (the reconstruction, or the untouched original when unpredictable).
template <class T>
SZ_HD int quantize(T data, T pred, const QuantParams &qp, T &recon) {
    // Straight-line form (selects instead of the reference's nested ifs): unpredictable points are rare, so nothing
    // is wasted, and the hot loops keep no divergent region.  NaN / overflow make `inrange` false (comparisons with
    // NaN are false) and whatever the arithmetic below produced is discarded.
    using U = typename Arith<T>::U;
    const T diff = static_cast<T>(static_cast<U>(data) - static_cast<U>(pred));
    const double v = fabs(static_cast<double>(diff)) * qp.ebr;
    const bool inrange = v < qp.vmax;
    const int qi = trunc_to_int(v) + 1;   // garbage when !inrange (discarded)
    const int half = qi >> 1;
    const bool neg = diff < 0;
    const int q2 = neg ? -(half << 1) : (half << 1);
    const int shifted = neg ? qp.radius - half : qp.radius + half;
    const T dec = from_double<T>(static_cast<double>(pred) + int_to_double(q2) * qp.eb);
    // fabs(decompressed_data - data), stored back into a T (LinearQuantizer.hpp:60-61)
    const T err = from_double<T>(fabs(static_cast<double>(static_cast<T>(static_cast<U>(dec) - static_cast<U>(data)))));
    bool ok;
    if (std::is_same<T, float>::value)
        ok = static_cast<float>(err) <= qp.ebf;
    else
        ok = static_cast<double>(err) <= qp.eb;
    ok = ok && inrange;
    if constexpr (std::is_integral<T>::value) {
        // |diff| / eb >= 2^63 (64-bit data a long way from its prediction, e.g. the wrapped sums of a regression fit):
        // the reference's static_cast<int64_t> yields x86's indefinite value and its code carries on with it
        // (LinearQuantizer.hpp:45-61): quant_index = INT64_MIN + 1 -> half_index 0, quant_index -2^63 either sign,
        // and the point is accepted with index `radius` when the (wrapped) error test happens to pass.
        if (!(v < 9223372036854775808.0)) {
            const T dec2 = from_double<T>(static_cast<double>(pred) + -9223372036854775808.0 * qp.eb);
            const T err2 = from_double<T>(fabs(static_cast<double>(static_cast<T>(static_cast<U>(dec2) - static_cast<U>(data)))));
            const bool ok2 = static_cast<double>(err2) <= qp.eb;
            recon = ok2 ? dec2 : data;
            return ok2 ? qp.radius : 0;
        }
    }
    recon = ok ? dec : data;
    return ok ? shifted : 0;
}

// quantize<float> with the integer conversions folded into the FP64 adds (same results bit for bit, four
// instructions shorter; device only -- the host build forwards to quantize<float>):
//   * RZ(v + (2^52 + 1)) has trunc(v) + 1 = qi in its low mantissa word (v >= 0; ulp is 1 in [2^52, 2^53));
//   * clearing bit 0 of that word and subtracting 2^52 gives (double)(qi & ~1) without an int -> double conversion;
//   * -(a * eb) == a * (-eb): the sign of diff goes onto eb with one logic operation, off the dependent chain.
// v >= 2^32 wraps qi; then |dec - data| >= (2^32 - 2^31) eb and the final test rejects the point, as `inrange` does.
SZ_HD int quantize_f32(float data, float pred, const QuantParams &qp, float &recon) {
#if defined(__CUDA_ARCH__)
    const float diff = data - pred;
    const double v = fabs(static_cast<double>(diff)) * qp.ebr;
    const bool inrange = v < qp.vmax;
    const double t = __dadd_rz(v, 4503599627370497.0);
    const int qi = __double2loint(t);
    const double dq = __hiloint2double(__double2hiint(t), qi & ~1) - 4503599627370496.0;
    const double ebs = __hiloint2double(__double2hiint(qp.eb) ^ (__float_as_int(diff) & static_cast<int>(0x80000000u)),
                                        __double2loint(qp.eb));
    const float dec = static_cast<float>(static_cast<double>(pred) + dq * ebs);
    const int half = qi >> 1;
    const int shifted = diff < 0 ? qp.radius - half : qp.radius + half;
    const bool ok = fabsf(dec - data) <= qp.ebf && inrange;
    recon = ok ? dec : data;
    return ok ? shifted : 0;
#else
    return quantize<float>(data, pred, qp, recon);
#endif
}

// recover (LinearQuantizer.hpp:74-86) for a predictable index.
template <class T>
SZ_HD T recover_pred(T pred, int q, const QuantParams &qp) {
    return from_double<T>(static_cast<double>(pred) + static_cast<double>(2 * (q - qp.radius)) * qp.eb);
}

// ---------------------------------------------------------------------------------------------------------------------
// Interpolators (reference include/SZ3/utils/Interpolators.hpp:12-39); T arithmetic, left to right.
// Divisions by 2/8/16 are exact scalings, written as multiplications by the exact reciprocal.
// ---------------------------------------------------------------------------------------------------------------------
// Integer T: the same expressions in (wrapping) integer arithmetic, the divisions truncating as in C++.
template <class T>
SZ_HD T interp_linear(T a, T b) {
    if constexpr (std::is_integral<T>::value) {
        using U = typename Arith<T>::U;
        return static_cast<T>(static_cast<U>(a) + static_cast<U>(b)) / 2;
    } else {
        return (a + b) * static_cast<T>(0.5);
    }
}
template <class T>
SZ_HD T interp_linear1(T a, T b) {  // double arithmetic in the reference (-0.5 and 1.5 are double literals)
    return from_double<T>(-0.5 * static_cast<double>(a) + 1.5 * static_cast<double>(b));
}
template <class T>
SZ_HD T interp_quad_1(T a, T b, T c) {
    if constexpr (std::is_integral<T>::value) {
        using U = typename Arith<T>::U;
        return static_cast<T>(U(3) * U(a) + U(6) * U(b) - U(c)) / 8;
    } else {
        return (static_cast<T>(3) * a + static_cast<T>(6) * b - c) * static_cast<T>(0.125);
    }
}
template <class T>
SZ_HD T interp_quad_2(T a, T b, T c) {
    if constexpr (std::is_integral<T>::value) {
        using U = typename Arith<T>::U;
        return static_cast<T>(U(0) - U(a) + U(6) * U(b) + U(3) * U(c)) / 8;
    } else {
        return (-a + static_cast<T>(6) * b + static_cast<T>(3) * c) * static_cast<T>(0.125);
    }
}
template <class T>
SZ_HD T interp_quad_3(T a, T b, T c) {
    if constexpr (std::is_integral<T>::value) {
        using U = typename Arith<T>::U;
        return static_cast<T>(U(3) * U(a) - U(10) * U(b) + U(15) * U(c)) / 8;
    } else {
        return (static_cast<T>(3) * a - static_cast<T>(10) * b + static_cast<T>(15) * c) * static_cast<T>(0.125);
    }
}
template <class T>
SZ_HD T interp_cubic(T a, T b, T c, T d) {
    if constexpr (std::is_integral<T>::value) {
        using U = typename Arith<T>::U;
        return static_cast<T>(U(0) - U(a) + U(9) * U(b) + U(9) * U(c) - U(d)) / 16;
    } else {
        return (-a + static_cast<T>(9) * b + static_cast<T>(9) * c - d) * static_cast<T>(0.0625);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Traversal geometry of InterpolationDecomposition (reference InterpolationDecomposition.hpp:79-147,309-454;
// SURVEY.md Appendix B).  A "level" has stride s; its blocks are 32*s wide and closed on both faces; a block emits,
// pass by pass, the points it owns (coordinate > begin, or begin == 0).
// ---------------------------------------------------------------------------------------------------------------------
struct InterpShape {
    int N;                    // 1..4
    uint32_t dims[kMaxDim];   // natural order, dims[N-1] fastest
    uint64_t stride[kMaxDim]; // element strides of the full-resolution array
    int perm[kMaxDim];        // dim_sequences[direction]: perm[p] is the dimension interpolated in pass p
    int cubic;                // 1 = cubic, 0 = linear
    int old_api;              // 1 = N<=2 line-by-line variant (InterpolationDecomposition.hpp:247-293,408-428)
};

// Geometry of one block of one level.
struct BlockGeom {
    uint32_t begin[kMaxDim];
    uint32_t end[kMaxDim];   // inclusive
    uint32_t n[kMaxDim];     // points per dim at spacing s: (end-begin)/s + 1
    uint32_t c1[kMaxDim];    // owned lattice points at step s
    uint32_t c2[kMaxDim];    // owned lattice points at step 2s
};

SZ_HD void block_geom(const InterpShape &sh, uint32_t s, const uint32_t bidx[kMaxDim], BlockGeom &g) {
    const uint32_t B = kInterpBlock * s;
    for (int d = 0; d < sh.N; d++) {
        uint32_t b = bidx[d] * B;
        uint32_t e = b + B;
        if (e > sh.dims[d] - 1) e = sh.dims[d] - 1;
        g.begin[d] = b;
        g.end[d] = e;
        g.n[d] = (e - b) / s + 1;
        g.c1[d] = b ? (e - b) / s : e / s + 1;
        g.c2[d] = b ? (e - b) / (2 * s) : e / (2 * s) + 1;
    }
}

// Per (block, pass) emission context.
struct PassGeom {
    int D;                    // dimension interpolated in this pass
    uint32_t n;               // points along D
    uint32_t lo[kMaxDim];     // first owned coordinate in every other dim
    uint32_t step[kMaxDim];   // lattice step in every other dim (s if already passed, 2s otherwise)
    uint32_t cnt[kMaxDim];    // owned lattice points in every other dim (cnt[D] unused)
    uint64_t other;           // product of cnt over d != D
    uint32_t main_cnt;        // indices along D in the main sub-phase
    uint32_t nbnd;            // boundary sub-phases
    uint32_t bnd[3];          // their local indices along D
    uint64_t size;            // points emitted by this pass
};

SZ_HD void pass_geom(const InterpShape &sh, uint32_t s, const BlockGeom &g, int p, PassGeom &pg) {
    const int D = sh.perm[p];
    pg.D = D;
    pg.n = g.n[D];
    uint64_t other = 1;
    for (int d = 0; d < sh.N; d++) {
        pg.lo[d] = 0;
        pg.step[d] = s;
        pg.cnt[d] = 1;
    }
    for (int q = 0; q < sh.N; q++) {
        int d = sh.perm[q];
        if (q == p) continue;
        uint32_t st = q < p ? s : 2 * s;
        pg.step[d] = st;
        pg.lo[d] = g.begin[d] ? g.begin[d] + st : 0;
        pg.cnt[d] = q < p ? g.c1[d] : g.c2[d];
        other *= pg.cnt[d];
    }
    pg.other = other;
    const uint32_t n = pg.n;
    pg.nbnd = 0;
    pg.main_cnt = 0;
    if (n <= 1) {
        pg.size = 0;
        pg.other = 0;
        return;
    }
    if (sh.cubic) {
        pg.main_cnt = n >= 7 ? (n - 7) / 2 + 1 : 0;  // odd i with 3 <= i <= n-4
        pg.bnd[pg.nbnd++] = 1;
        if ((n & 1) && n > 3) pg.bnd[pg.nbnd++] = n - 2;
        if (!(n & 1) && n > 4) pg.bnd[pg.nbnd++] = n - 3;
        if (!(n & 1) && n > 2) pg.bnd[pg.nbnd++] = n - 1;
    } else {
        pg.main_cnt = (n - 1) / 2;  // odd i with i <= n-2
        if (!(n & 1)) pg.bnd[pg.nbnd++] = n - 1;
    }
    pg.size = static_cast<uint64_t>(pg.main_cnt + pg.nbnd) * other;
}

// Offset (inside the pass) at which the point with global coordinates x (local index i along D) is emitted, or
// ~0ull when this block does not own the point (it lies on a low face shared with the previous block).
SZ_HD uint64_t pass_offset(const InterpShape &sh, const PassGeom &pg, const uint32_t x[kMaxDim], uint32_t i) {
    uint64_t base;
    uint32_t idxD, extD;
    const uint32_t n = pg.n;
    bool in_main = sh.cubic ? (i >= 3 && i + 3 < n) : (i + 1 < n);
    if (in_main) {
        base = 0;
        idxD = sh.cubic ? (i - 3) >> 1 : (i - 1) >> 1;
        extD = pg.main_cnt;
    } else {
        uint32_t k = 0;
        while (k < pg.nbnd && pg.bnd[k] != i) k++;
        base = static_cast<uint64_t>(pg.main_cnt + k) * pg.other;
        idxD = 0;
        extD = 1;
    }
    uint64_t rank = 0;
    for (int d = 0; d < sh.N; d++) {
        if (d == pg.D) {
            rank = rank * extD + idxD;
        } else {
            if (x[d] < pg.lo[d]) return ~0ull;
            rank = rank * pg.cnt[d] + (x[d] - pg.lo[d]) / pg.step[d];
        }
    }
    return base + rank;
}

// Prediction of local index i (odd) on a line of n points; v(k) returns the current value at local index k.
// New (N>=3) API: InterpolationDecomposition.hpp:334-400.  The linear-mode tail (i == n-1, n even, n >= 4) needs the
// *reconstruction* of i-2, which the caller passes as recon_im2.
template <class T, class F>
SZ_HD T predict_line(int cubic, uint32_t i, uint32_t n, F &&v, T recon_im2) {
    if (cubic) {
        if (i >= 3) {
            if (i + 3 < n) return interp_cubic<T>(v(i - 3), v(i - 1), v(i + 1), v(i + 3));
            if (i + 1 < n) return interp_quad_2<T>(v(i - 3), v(i - 1), v(i + 1));
            return interp_linear1<T>(v(i - 3), v(i - 1));
        }
        if (i + 3 < n) return interp_quad_1<T>(v(i - 1), v(i + 1), v(i + 3));
        if (i + 1 < n) return interp_linear<T>(v(i - 1), v(i + 1));
        return v(i - 1);
    }
    if (i + 1 < n) return interp_linear<T>(v(i - 1), v(i + 1));
    if (n < 3) return v(i - 1);
    return interp_linear1<T>(recon_im2, v(i - 1));
}

// Old (N<=2) API line predictor: InterpolationDecomposition.hpp:247-293.
template <class T, class F>
SZ_HD T predict_line_old(int cubic, uint32_t i, uint32_t n, F &&v) {
    if (!cubic || n < 5) {
        if (i + 1 < n) return interp_linear<T>(v(i - 1), v(i + 1));
        if (n < 4) return v(i - 1);
        return interp_linear1<T>(v(i - 3), v(i - 1));
    }
    if (i >= 3 && i + 3 < n) return interp_cubic<T>(v(i - 3), v(i - 1), v(i + 1), v(i + 3));
    if (i == 1) return interp_quad_1<T>(v(i - 1), v(i + 1), v(i + 3));
    if (i + 1 < n) return interp_quad_2<T>(v(i - 3), v(i - 1), v(i + 1));
    return interp_quad_3<T>(v(i - 5), v(i - 3), v(i - 1));
}

// Emission offset inside a 1-D line of the old API: main (3,5,..), then i=1, then the quad_2 point, then the tail.
SZ_HD uint32_t line_offset_old(int cubic, uint32_t i, uint32_t n) {
    if (!cubic || n < 5) return (i - 1) >> 1;  // natural order 1,3,5,...
    uint32_t main_cnt = n >= 7 ? (n - 7) / 2 + 1 : 0;
    if (i >= 3 && i + 3 < n) return (i - 3) >> 1;
    if (i == 1) return main_cnt;
    if (i + 1 < n) return main_cnt + 1;
    return main_cnt + 2;
}

}  // namespace sz3b
