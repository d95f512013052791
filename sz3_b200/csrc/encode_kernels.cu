// sz3_b200/csrc/encode_kernels.cu -- GPU replacement of HuffmanEncoder's data-proportional loops
// (reference include/SZ3/encoder/HuffmanEncoder.hpp): the frequency count of init() (:516-527) and the serial bit
// concatenation of encode() (:140-218).  The tree itself (a few thousand nodes) is built on the host
// (huffman_host.cpp) from the histogram these kernels produce.
//
//   k_histogram      warp-aggregated atomics into a shared-memory window, spill to a 64-bit global histogram
//                    (stand-alone variant for side streams; the main stream's histogram is fused into the
//                    predict+quantize kernels, see device_ctx.cuh)
//   k_pack_count     per-chunk code-length sums and unpredictable counts
//   k_pack_scan      exclusive scan of the chunk sums (single CTA)
//   k_pack_write     warp-prefix-scan bit packer: every thread concatenates its symbols in registers, the CTA
//                    assembles its bit range in shared memory and stores it coalesced, MSB-first as the reference
//                    decoder expects (:239-243); the same pass compacts the unpredictable values in stream order.
#include <cuda_runtime.h>

#include "launch.hpp"

namespace sz3b {

constexpr int kPackThreads = 256;
constexpr int kPackPerThread = 16;
constexpr int kPackChunk = kPackThreads * kPackPerThread;  // 4096 symbols per CTA

// ---------------------------------------------------------------------------------------------------------------------
template <class QT>
__global__ void __launch_bounds__(256) k_histogram(const QT *__restrict__ q, uint64_t n, int sym_min, int nbins,
                                                   int center, unsigned long long *__restrict__ ghist) {
    constexpr int W = 2048;
    __shared__ unsigned sh[W];
    for (int i = threadIdx.x; i < W; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const int lo = center - W / 2;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const uint64_t nround = (n + stride - 1) / stride;
    uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (uint64_t r = 0; r < nround; r++, i += stride) {
        bool active = i < n;
        int sym = active ? static_cast<int>(q[i]) - sym_min : -1;
        // warp-aggregated: one atomic per distinct symbol per warp
        unsigned peers = __match_any_sync(0xffffffffu, sym);
        int leader = __ffs(peers) - 1;
        if (active && static_cast<int>(threadIdx.x & 31) == leader) {
            unsigned c = __popc(peers);
            unsigned k = static_cast<unsigned>(sym - lo);
            if (k < static_cast<unsigned>(W))
                atomicAdd(&sh[k], c);
            else if (sym >= 0 && sym < nbins)
                atomicAdd(&ghist[sym], static_cast<unsigned long long>(c));
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < W; k += blockDim.x) {
        unsigned v = sh[k];
        int sym = lo + k;
        if (v && sym >= 0 && sym < nbins) atomicAdd(&ghist[sym], static_cast<unsigned long long>(v));
    }
}

// Histogram of a stretch of 16-bit quantization indices (the streams the box schedule writes, interp_box.cu: counting
// inside that kernel cost a quarter of its time -- 14 instructions per point on its dependent path -- while a pass
// over the indices afterwards is bound by reading them, mostly from L2).  A thread takes eight indices per 16-byte
// load; the sixteen bins around the radius are counted in two registers of 4-bit fields (a shift and an add per index,
// an index outside the window shifts its increment out), widened to 8-bit fields once per load, reduced across the
// warp at the end; everything else (rare) goes to a 2048-bin shared-memory window or to the global histogram.
constexpr int kHistChunk = 256 * 24 * 8;   // indices per CTA: 24 loads per thread (8-bit fields hold 24 * 8 = 192)
__device__ __forceinline__ unsigned shl_clamp(unsigned v, unsigned sh) {
    unsigned r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(sh));   // 0 for sh >= 32
    return r;
}
__global__ void __launch_bounds__(256) k_hist_u16(const uint16_t *__restrict__ q, uint64_t n, int radius, int nbins,
                                                  unsigned long long *__restrict__ ghist) {
    constexpr int W = 2048;
    __shared__ unsigned sh[W];
    for (int i = threadIdx.x; i < W; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const int lo = radius - W / 2, lo16 = radius - 8;
    auto rare = [&](int sym) {
        const unsigned k = static_cast<unsigned>(sym - lo);
        if (k < static_cast<unsigned>(W))
            atomicAdd(&sh[k], 1u);
        else if (sym >= 0 && sym < nbins)
            atomicAdd(&ghist[sym], 1ull);
    };
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kHistChunk;
    const uint64_t end = base + kHistChunk < n ? base + kHistChunk : n;
    // 16-byte loads need an aligned start: the indices before it (and a ragged tail) go one by one
    const uint64_t a0 = (reinterpret_cast<uintptr_t>(q + base) & 15u) ? base + (16 - (reinterpret_cast<uintptr_t>(q + base) & 15u)) / 2 : base;
    const uint64_t astart = a0 < end ? a0 : end;
    const uint64_t nvec = (end - astart) / 8;
    for (uint64_t i = base + threadIdx.x; i < astart; i += blockDim.x) rare(q[i]);
    for (uint64_t i = astart + nvec * 8 + threadIdx.x; i < end; i += blockDim.x) rare(q[i]);
    unsigned acc[4] = {0, 0, 0, 0};   // 8-bit fields: bins 0,2,4,6 | 1,3,5,7 | 8,10,12,14 | 9,11,13,15 of the window
    const uint4 *v = reinterpret_cast<const uint4 *>(q + astart);
    for (uint64_t j = threadIdx.x; j < nvec; j += blockDim.x) {
        const uint4 w = v[j];
        const unsigned words[4] = {w.x, w.y, w.z, w.w};
        unsigned nl = 0, nh = 0, any = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const unsigned sym = (k & 1) ? words[k >> 1] >> 16 : words[k >> 1] & 0xffffu;
            const unsigned k16 = sym - static_cast<unsigned>(lo16);
            nl += shl_clamp(1u, k16 * 4u);
            nh += shl_clamp(1u, k16 * 4u - 32u);
            any |= k16;
        }
        acc[0] += nl & 0x0f0f0f0fu;
        acc[1] += (nl >> 4) & 0x0f0f0f0fu;
        acc[2] += nh & 0x0f0f0f0fu;
        acc[3] += (nh >> 4) & 0x0f0f0f0fu;
        if (any > 15u) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const unsigned sym = (k & 1) ? words[k >> 1] >> 16 : words[k >> 1] & 0xffffu;
                if (sym - static_cast<unsigned>(lo16) > 15u) rare(static_cast<int>(sym));
            }
        }
    }
    // warp reduction of the sixteen counters: 8-bit fields widened to 16 bits (a warp's sum is at most 32 * 192)
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const unsigned e = __reduce_add_sync(0xffffffffu, acc[r] & 0x00ff00ffu);          // fields 0 | 2 of this register
        const unsigned o = __reduce_add_sync(0xffffffffu, (acc[r] >> 8) & 0x00ff00ffu);   // fields 1 | 3
        if (lane < 4) {
            const unsigned cnt = (lane & 1u) ? ((lane & 2u) ? o >> 16 : o & 0xffffu) : ((lane & 2u) ? e >> 16 : e & 0xffffu);
            // field f of register r counts window bin 8 * (r / 2) + 2 * f + (r & 1)
            const unsigned bin = 8u * (r >> 1) + 2u * lane + (r & 1u);
            if (cnt) atomicAdd(&sh[lo16 - lo + static_cast<int>(bin)], cnt);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < W; k += blockDim.x) {
        const unsigned c = sh[k];
        const int sym = lo + k;
        if (c && sym >= 0 && sym < nbins) atomicAdd(&ghist[sym], static_cast<unsigned long long>(c));
    }
}

void launch_hist_u16(const uint16_t *q, uint64_t n, int radius, int nbins, unsigned long long *ghist, cudaStream_t st) {
    if (n == 0) return;
    const uint64_t blocks = (n + kHistChunk - 1) / kHistChunk;
    k_hist_u16<<<static_cast<unsigned>(blocks), 256, 0, st>>>(q, n, radius, nbins, ghist);
}

template <class QT>
__global__ void __launch_bounds__(256) k_minmax_int(const QT *__restrict__ q, uint64_t n, int *__restrict__ mm) {
    int lo = 0x7fffffff, hi = static_cast<int>(0x80000000);
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        int v = static_cast<int>(q[i]);
        lo = min(lo, v);
        hi = max(hi, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[0], lo);
        atomicMax(&mm[1], hi);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// block-wide exclusive scan of one 32-bit value per thread (256 threads); returns the exclusive prefix, total in *tot
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned *warp_sums, unsigned *tot) {
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= static_cast<unsigned>(o)) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned w = lane < (kPackThreads / 32) ? warp_sums[lane] : 0;
        unsigned winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= static_cast<unsigned>(o)) winc += t;
        }
        if (lane < (kPackThreads / 32)) warp_sums[lane] = winc - w;
        if (lane == (kPackThreads / 32) - 1) *tot = winc;
    }
    __syncthreads();
    unsigned r = warp_sums[wid] + inc - v;
    return r;
}

// The packer works on chunks of kPackChunk symbols; a CTA walks over chunks (a few CTAs per SM), a thread takes
// kPackPerThread consecutive symbols of the chunk.
//
// Code table: index streams crowd around one symbol, so a window of kPackWin states around the most frequent one sits
// in shared memory -- entry = len << 24 | code for codes of 1..24 bits, 0 for everything else (absent states, longer
// codes, the unpredictable marker `zero_sym`, states outside the window: index kPackWin).  The hot loops are straight
// line code over the window; a thread that met a 0 entry (rare) then visits just those symbols again with the tables
// in global memory.

// Symbols of one thread of a full chunk.  16-bit indices of a 16-byte aligned stream come in by two 16-byte loads (a
// warp then touches every sector once); anything else element by element.
template <class QT>
__device__ __forceinline__ void pack_load_full(const QT *__restrict__ q, uint64_t base, bool vec, int (&v)[kPackPerThread]) {
    if (sizeof(QT) == 2 && vec) {
        const uint4 *p = reinterpret_cast<const uint4 *>(q + base);
        const uint4 a = p[0], b = p[1];
        const unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            v[2 * k] = static_cast<int>(w[k] & 0xffffu);
            v[2 * k + 1] = static_cast<int>(w[k] >> 16);
        }
    } else {
#pragma unroll
        for (int k = 0; k < kPackPerThread; k++) v[k] = static_cast<int>(q[base + k]);
    }
}

constexpr int kPackWin = 256;
__device__ __forceinline__ void pack_window(unsigned *swin, int wlo, unsigned nstates, int zero_state, const uint8_t *__restrict__ len,
                                            const unsigned long long *__restrict__ code) {
    for (int j = threadIdx.x; j <= kPackWin; j += blockDim.x) {
        const int s = wlo + j;
        unsigned e = 0;
        if (j < kPackWin && s >= 0 && static_cast<unsigned>(s) < nstates && s != zero_state) {
            const unsigned l = len[s];
            if (l && l <= 24) e = (l << 24) | static_cast<unsigned>(code[s]);
        }
        swin[j] = e;
    }
}
__device__ __forceinline__ unsigned pack_lookup(const unsigned *swin, int v, int woff) {
    const unsigned u = static_cast<unsigned>(v - woff);   // woff = sym_min + wlo
    return swin[u < static_cast<unsigned>(kPackWin) ? u : static_cast<unsigned>(kPackWin)];
}

template <class QT>
__global__ void __launch_bounds__(kPackThreads) k_pack_count(const QT *__restrict__ q, uint64_t n, int sym_min,
                                                             int zero_sym, const uint8_t *__restrict__ len,
                                                             const unsigned long long *__restrict__ code, int wlo,
                                                             unsigned nstates, bool vec, uint64_t nchunks,
                                                             unsigned *__restrict__ chunk_bits,
                                                             unsigned *__restrict__ chunk_zeros) {
    __shared__ unsigned swin[kPackWin + 1];
    __shared__ unsigned ws[2][kPackThreads / 32];
    __shared__ unsigned wz[2][kPackThreads / 32];
    pack_window(swin, wlo, nstates, zero_sym - sym_min, len, code);
    __syncthreads();
    const int woff = sym_min + wlo;
    unsigned par = 0;
    for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x, par ^= 1u) {
        const uint64_t base = c * kPackChunk + static_cast<uint64_t>(threadIdx.x) * kPackPerThread;
        unsigned bits = 0, zeros = 0;
        if ((c + 1) * kPackChunk <= n) {
            int v[kPackPerThread];
            pack_load_full(q, base, vec, v);
            unsigned lmin = 0xffu;
#pragma unroll
            for (int k = 0; k < kPackPerThread; k++) {
                const unsigned l = pack_lookup(swin, v[k], woff) >> 24;
                bits += l;
                lmin = l < lmin ? l : lmin;
            }
            if (lmin == 0) {
#pragma unroll
                for (int k = 0; k < kPackPerThread; k++) {
                    if (pack_lookup(swin, v[k], woff) == 0) {
                        bits += len[v[k] - sym_min];
                        zeros += (v[k] == zero_sym);
                    }
                }
            }
        } else {
            for (int k = 0; k < kPackPerThread; k++) {
                if (base + k < n) {
                    const int vk = static_cast<int>(q[base + k]);
                    bits += len[vk - sym_min];
                    zeros += (vk == zero_sym);
                }
            }
        }
        bits = __reduce_add_sync(0xffffffffu, bits);
        zeros = __reduce_add_sync(0xffffffffu, zeros);
        if ((threadIdx.x & 31) == 0) {
            ws[par][threadIdx.x >> 5] = bits;
            wz[par][threadIdx.x >> 5] = zeros;
        }
        __syncthreads();   // (the two parities keep a fast warp of the next chunk off this chunk's sums)
        if (threadIdx.x < 32) {
            unsigned b = threadIdx.x < kPackThreads / 32 ? ws[par][threadIdx.x] : 0u;
            unsigned z = threadIdx.x < kPackThreads / 32 ? wz[par][threadIdx.x] : 0u;
            b = __reduce_add_sync(0xffffffffu, b);
            z = __reduce_add_sync(0xffffffffu, z);
            if (threadIdx.x == 0) {
                chunk_bits[c] = b;
                chunk_zeros[c] = z;
            }
        }
    }
}

// single CTA; off arrays get nchunks+1 entries (last = total).  Rounds of 4096 chunks: a thread takes four consecutive
// ones (coalesced across the CTA), the 1024 partial sums are scanned through shuffles and shared memory.
__global__ void __launch_bounds__(1024) k_pack_scan(const unsigned *__restrict__ chunk_bits,
                                                    const unsigned *__restrict__ chunk_zeros, uint64_t nchunks,
                                                    unsigned long long *__restrict__ bit_off,
                                                    unsigned long long *__restrict__ zero_off) {
    __shared__ unsigned long long wsum[2][2][32];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long carry0 = 0, carry1 = 0;
    unsigned par = 0;
    for (uint64_t base = 0; base < nchunks; base += 4096, par ^= 1u) {
        const uint64_t i = base + 4ull * threadIdx.x;
        unsigned b[4], z[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            b[k] = i + k < nchunks ? chunk_bits[i + k] : 0u;
            z[k] = i + k < nchunks ? chunk_zeros[i + k] : 0u;
        }
        const unsigned long long v0 = static_cast<unsigned long long>(b[0]) + b[1] + b[2] + b[3];
        const unsigned long long v1 = static_cast<unsigned long long>(z[0]) + z[1] + z[2] + z[3];
        unsigned long long inc0 = v0, inc1 = v1;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t0 = __shfl_up_sync(0xffffffffu, inc0, o);
            unsigned long long t1 = __shfl_up_sync(0xffffffffu, inc1, o);
            if (lane >= static_cast<unsigned>(o)) {
                inc0 += t0;
                inc1 += t1;
            }
        }
        if (lane == 31) {
            wsum[par][0][wid] = inc0;
            wsum[par][1][wid] = inc1;
        }
        __syncthreads();
        // every warp scans the 32 warp totals itself (no second barrier; the two parities keep rounds apart)
        unsigned long long w0 = wsum[par][0][lane], w1 = wsum[par][1][lane];
        unsigned long long a0 = w0, a1 = w1;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t0 = __shfl_up_sync(0xffffffffu, a0, o);
            unsigned long long t1 = __shfl_up_sync(0xffffffffu, a1, o);
            if (lane >= static_cast<unsigned>(o)) {
                a0 += t0;
                a1 += t1;
            }
        }
        const unsigned long long tot0 = __shfl_sync(0xffffffffu, a0, 31), tot1 = __shfl_sync(0xffffffffu, a1, 31);
        const unsigned long long before0 = __shfl_sync(0xffffffffu, a0 - w0, wid), before1 = __shfl_sync(0xffffffffu, a1 - w1, wid);
        unsigned long long e0 = carry0 + before0 + inc0 - v0;
        unsigned long long e1 = carry1 + before1 + inc1 - v1;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i + k < nchunks) {
                bit_off[i + k] = e0;
                zero_off[i + k] = e1;
            }
            e0 += b[k];
            e1 += z[k];
        }
        carry0 += tot0;
        carry1 += tot1;
    }
    if (threadIdx.x == 0) {
        bit_off[nchunks] = carry0;
        zero_off[nchunks] = carry1;
    }
}

// Bit accumulator: `acc` holds `nacc` (< 32 between appends) pending bits, left-aligned in a 64-bit register.
struct BitAcc {
    unsigned long long acc;
    unsigned nacc;
    unsigned w;       // next word index in the CTA buffer
    bool first;       // the next flushed word may be shared with the previous thread
};

__device__ __forceinline__ void acc_flush(BitAcc &a, unsigned *sbits) {
    if (a.nacc >= 32) {
        unsigned word = static_cast<unsigned>(a.acc >> 32);
        if (a.first) {
            atomicOr(&sbits[a.w], word);
            a.first = false;
        } else {
            sbits[a.w] = word;
        }
        a.w++;
        a.acc <<= 32;
        a.nacc -= 32;
    }
}
__device__ __forceinline__ void acc_put(BitAcc &a, unsigned bits, unsigned l, unsigned *sbits) {  // l <= 32
    if (l == 0) return;
    a.acc |= static_cast<unsigned long long>(bits) << (64 - a.nacc - l);
    a.nacc += l;
    acc_flush(a, sbits);
}
__device__ __forceinline__ void acc_put_long(BitAcc &a, unsigned long long cw, unsigned l, unsigned *sbits) {  // l <= 64
    if (l > 32) {
        acc_put(a, static_cast<unsigned>(cw >> 32), l - 32, sbits);
        acc_put(a, static_cast<unsigned>(cw), 32, sbits);
    } else {
        acc_put(a, static_cast<unsigned>(cw), l, sbits);
    }
}

template <class QT, class T>
__global__ void __launch_bounds__(kPackThreads, 4) k_pack_write(const QT *__restrict__ q, uint64_t n, int sym_min,
                                                             int zero_sym, const uint8_t *__restrict__ len,
                                                             const unsigned long long *__restrict__ code, int wlo,
                                                             unsigned nstates, bool vec, uint64_t nchunks,
                                                             const unsigned long long *__restrict__ bit_off,
                                                             const unsigned long long *__restrict__ zero_off,
                                                             unsigned *__restrict__ out_words,
                                                             const T *__restrict__ unpred_tmp,
                                                             T *__restrict__ unpred_out) {
    // worst case 4096 symbols x 64 bits + 31 leading bits
    __shared__ unsigned sbits[kPackChunk * 2 + 2];
    __shared__ unsigned swin[kPackWin + 1];
    __shared__ unsigned warp_sums[kPackThreads / 32];
    __shared__ unsigned tot_bits, tot_zeros;
    pack_window(swin, wlo, nstates, zero_sym - sym_min, len, code);
    __syncthreads();
    const int woff = sym_min + wlo;
    constexpr unsigned kLong = 0xffffffffu;   // entry of a symbol whose code does not fit an entry: tables in global memory
    for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const uint64_t base = c * kPackChunk + static_cast<uint64_t>(threadIdx.x) * kPackPerThread;
        const bool full = (c + 1) * kPackChunk <= n;
        int v[kPackPerThread];
        unsigned ent[kPackPerThread];   // len << 24 | code; 0 = no symbol (ragged tail); kLong
        unsigned bits = 0, zeros = 0;
        bool has_long = false;
        if (full) {
            pack_load_full(q, base, vec, v);
            unsigned emin = kLong;
#pragma unroll
            for (int k = 0; k < kPackPerThread; k++) {
                ent[k] = pack_lookup(swin, v[k], woff);
                bits += ent[k] >> 24;
                emin = ent[k] < emin ? ent[k] : emin;
            }
            if (emin == 0) {
#pragma unroll
                for (int k = 0; k < kPackPerThread; k++) {
                    if (ent[k] == 0) {
                        const unsigned l = len[v[k] - sym_min];
                        bits += l;
                        zeros += (v[k] == zero_sym);
                        ent[k] = (l && l <= 24) ? (l << 24) | static_cast<unsigned>(code[v[k] - sym_min]) : kLong;
                        has_long = has_long || ent[k] == kLong;
                    }
                }
            }
        } else {
            for (int k = 0; k < kPackPerThread; k++) {
                ent[k] = 0;
                v[k] = zero_sym + 1;
                if (base + k < n) {
                    v[k] = static_cast<int>(q[base + k]);
                    bits += len[v[k] - sym_min];
                    zeros += (v[k] == zero_sym);
                    ent[k] = kLong;
                    has_long = true;
                }
            }
        }
        const unsigned long long B0 = bit_off[c];
        const unsigned lead = static_cast<unsigned>(B0 & 31);
        unsigned my_bit = block_excl_scan(bits, warp_sums, &tot_bits);
        __syncthreads();
        const unsigned nwords = (lead + tot_bits + 31) >> 5;
        for (unsigned i = threadIdx.x; i < nwords; i += kPackThreads) sbits[i] = 0;
        __syncthreads();
        const unsigned b = lead + my_bit;
        if (full && !has_long && bits <= 64) {
            // the usual case: the thread's sixteen codes fit one 64-bit register -- concatenated without any flush
            // test, left-aligned, then dropped into (at most) three words of the CTA buffer
            unsigned long long t = 0;
#pragma unroll
            for (int k = 0; k < kPackPerThread; k++) t = (t << (ent[k] >> 24)) | (ent[k] & 0x00ffffffu);
            t <<= 64 - bits;   // bits >= 16 here (every symbol has a code of at least one bit)
            const unsigned hi = static_cast<unsigned>(t >> 32), lo = static_cast<unsigned>(t), o = b & 31u, w = b >> 5;
            const unsigned w0 = hi >> o, w1 = __funnelshift_r(lo, hi, o), w2 = __funnelshift_r(0u, lo, o);
            atomicOr(&sbits[w], w0);
            if (w1) atomicOr(&sbits[w + 1], w1);
            if (w2) atomicOr(&sbits[w + 2], w2);
        } else if (bits) {
            BitAcc a;
            a.acc = 0;
            a.nacc = b & 31;
            a.w = b >> 5;
            a.first = true;
#pragma unroll
            for (int k = 0; k < kPackPerThread; k++) {
                const unsigned e = ent[k];
                if (e == kLong) {
                    acc_put_long(a, code[v[k] - sym_min], len[v[k] - sym_min], sbits);
                } else if (full) {   // 1 <= len <= 24
                    const unsigned l = e >> 24;
                    a.acc |= static_cast<unsigned long long>(e & 0x00ffffffu) << (64 - a.nacc - l);
                    a.nacc += l;
                    acc_flush(a, sbits);
                }
            }
            if (a.nacc) atomicOr(&sbits[a.w], static_cast<unsigned>(a.acc >> 32));
        }
        __syncthreads();
        // store: stream bit 0 of a word is its MSB -> byte-swap to memory order
        const unsigned long long W0 = B0 >> 5;
        for (unsigned i = threadIdx.x; i < nwords; i += kPackThreads) {
            unsigned w = __byte_perm(sbits[i], 0, 0x0123);
            if (i == 0 || i == nwords - 1) {
                if (w) atomicOr(&out_words[W0 + i], w);
            } else {
                out_words[W0 + i] = w;
            }
        }
        // ordered compaction of the unpredictable values (LinearQuantizer::unpred, reference LinearQuantizer.hpp:63,68)
        if (unpred_out != nullptr) {
            unsigned my_zero = block_excl_scan(zeros, warp_sums, &tot_zeros);
            if (zeros) {
                unsigned long long zo = zero_off[c] + my_zero;
#pragma unroll
                for (int k = 0; k < kPackPerThread; k++) {
                    if (v[k] == zero_sym) unpred_out[zo++] = unpred_tmp[base + k];
                }
            }
        }
        __syncthreads();   // sbits / warp_sums are reused by the next chunk
    }
}

// ---------------------------------------------------------------------------------------------------------------------
template <class QT>
void launch_histogram(const QT *q, uint64_t n, int sym_min, int nbins, int center, unsigned long long *ghist,
                      cudaStream_t st) {
    if (n == 0) return;
    uint64_t blocks = (n + 256 * 16 - 1) / (256 * 16);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_histogram<QT><<<static_cast<unsigned>(blocks), 256, 0, st>>>(q, n, sym_min, nbins, center, ghist);
}

template <class QT>
void launch_minmax_int(const QT *q, uint64_t n, int *mm, cudaStream_t st) {
    if (n == 0) return;
    uint64_t blocks = (n + 256 * 16 - 1) / (256 * 16);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_minmax_int<QT><<<static_cast<unsigned>(blocks), 256, 0, st>>>(q, n, mm);
}

// Long inputs (the Huffman decoder's 10^5 .. 10^6 subsequence counts): three launches instead of one CTA walking over
// everything -- sums of 4096-element tiles, their scan (k_pack_scan over the tile sums), then every tile scans itself
// from its own offset.
constexpr int kScanTile = 4096;
__global__ void __launch_bounds__(1024) k_scan_tile_sums(const unsigned *__restrict__ a, const unsigned *__restrict__ b, uint64_t n,
                                                         unsigned *__restrict__ sa, unsigned *__restrict__ sb) {
    __shared__ unsigned w[2][32];
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kScanTile + 4ull * threadIdx.x;
    unsigned va = 0, vb = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (i + k < n) {
            va += a[i + k];
            vb += b[i + k];
        }
    }
    va = __reduce_add_sync(0xffffffffu, va);
    vb = __reduce_add_sync(0xffffffffu, vb);
    if ((threadIdx.x & 31) == 0) {
        w[0][threadIdx.x >> 5] = va;
        w[1][threadIdx.x >> 5] = vb;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        va = __reduce_add_sync(0xffffffffu, w[0][threadIdx.x]);
        vb = __reduce_add_sync(0xffffffffu, w[1][threadIdx.x]);
        if (threadIdx.x == 0) {
            sa[blockIdx.x] = va;   // (a tile's sum fits 32 bits wherever its elements' sum does: chunk bit counts
            sb[blockIdx.x] = vb;   //  are < 2^18 each, symbol counts < 2^11)
        }
    }
}
__global__ void __launch_bounds__(1024) k_scan_tiles(const unsigned *__restrict__ a, const unsigned *__restrict__ b, uint64_t n,
                                                     const unsigned long long *__restrict__ ta,
                                                     const unsigned long long *__restrict__ tb, uint64_t ntiles,
                                                     unsigned long long *__restrict__ oa, unsigned long long *__restrict__ ob) {
    __shared__ unsigned long long wsum[2][32];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kScanTile + 4ull * threadIdx.x;
    unsigned x[4], y[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        x[k] = i + k < n ? a[i + k] : 0u;
        y[k] = i + k < n ? b[i + k] : 0u;
    }
    const unsigned long long v0 = static_cast<unsigned long long>(x[0]) + x[1] + x[2] + x[3];
    const unsigned long long v1 = static_cast<unsigned long long>(y[0]) + y[1] + y[2] + y[3];
    unsigned long long inc0 = v0, inc1 = v1;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t0 = __shfl_up_sync(0xffffffffu, inc0, o);
        unsigned long long t1 = __shfl_up_sync(0xffffffffu, inc1, o);
        if (lane >= static_cast<unsigned>(o)) {
            inc0 += t0;
            inc1 += t1;
        }
    }
    if (lane == 31) {
        wsum[0][wid] = inc0;
        wsum[1][wid] = inc1;
    }
    __syncthreads();
    unsigned long long w0 = wsum[0][lane], w1 = wsum[1][lane];
    unsigned long long a0 = w0, a1 = w1;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t0 = __shfl_up_sync(0xffffffffu, a0, o);
        unsigned long long t1 = __shfl_up_sync(0xffffffffu, a1, o);
        if (lane >= static_cast<unsigned>(o)) {
            a0 += t0;
            a1 += t1;
        }
    }
    unsigned long long e0 = ta[blockIdx.x] + __shfl_sync(0xffffffffu, a0 - w0, wid) + inc0 - v0;
    unsigned long long e1 = tb[blockIdx.x] + __shfl_sync(0xffffffffu, a1 - w1, wid) + inc1 - v1;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (i + k < n) {
            oa[i + k] = e0;
            ob[i + k] = e1;
        }
        e0 += x[k];
        e1 += y[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        oa[n] = ta[ntiles];
        ob[n] = tb[ntiles];
    }
}

// 64-bit words of scratch launch_scan_chunks wants for `nchunks` elements (0: the single-CTA scan does it)
size_t scan_scratch_words(uint64_t nchunks) {
    if (nchunks <= 4 * static_cast<uint64_t>(kScanTile)) return 0;
    const uint64_t ntiles = (nchunks + kScanTile - 1) / kScanTile;
    return static_cast<size_t>(2 * (ntiles + 1) + ntiles + 1);
}

void launch_scan_chunks(const unsigned *chunk_bits, const unsigned *chunk_zeros, uint64_t nchunks, unsigned long long *bit_off,
                        unsigned long long *zero_off, cudaStream_t st, unsigned long long *scratch) {
    if (scratch == nullptr || scan_scratch_words(nchunks) == 0) {
        k_pack_scan<<<1, 1024, 0, st>>>(chunk_bits, chunk_zeros, nchunks, bit_off, zero_off);
        return;
    }
    const uint64_t ntiles = (nchunks + kScanTile - 1) / kScanTile;
    // scratch: offsets of the tiles (2 x u64, one extra entry each for the totals), then their sums (2 x u32)
    unsigned long long *ta = scratch, *tb = ta + ntiles + 1;
    unsigned *sa = reinterpret_cast<unsigned *>(tb + ntiles + 1), *sb = sa + ntiles;
    k_scan_tile_sums<<<static_cast<unsigned>(ntiles), 1024, 0, st>>>(chunk_bits, chunk_zeros, nchunks, sa, sb);
    k_pack_scan<<<1, 1024, 0, st>>>(sa, sb, ntiles, ta, tb);
    k_scan_tiles<<<static_cast<unsigned>(ntiles), 1024, 0, st>>>(chunk_bits, chunk_zeros, nchunks, ta, tb, ntiles, bit_off, zero_off);
}

uint64_t pack_num_chunks(uint64_t n) { return (n + kPackChunk - 1) / kPackChunk; }

template <class QT, class T>
void launch_pack(const QT *q, uint64_t n, int sym_min, int zero_sym, const uint8_t *len,
                 const unsigned long long *code, unsigned nstates, int center_state, unsigned *chunk_bits,
                 unsigned *chunk_zeros, unsigned long long *bit_off, unsigned long long *zero_off, unsigned *out_words,
                 const T *unpred_tmp, T *unpred_out, cudaStream_t st, cudaEvent_t after_scan, unsigned long long *scan_scratch) {
    const uint64_t nchunks = pack_num_chunks(n);
    if (nchunks == 0) return;
    const bool vec = (reinterpret_cast<uintptr_t>(q) & 15u) == 0;
    const int wlo = center_state - kPackWin / 2;
    // CTAs walk over the chunks: 8 (count) / 4 (write: 34 KB of shared memory, 64 registers) resident CTAs per SM
    const unsigned g1 = static_cast<unsigned>(nchunks < 148u * 8u ? nchunks : 148u * 8u);
    const unsigned g2 = static_cast<unsigned>(nchunks < 148u * 4u ? nchunks : 148u * 4u);
    k_pack_count<QT><<<g1, kPackThreads, 0, st>>>(q, n, sym_min, zero_sym, len, code, wlo, nstates, vec, nchunks, chunk_bits,
                                                  chunk_zeros);
    launch_scan_chunks(chunk_bits, chunk_zeros, nchunks, bit_off, zero_off, st, scan_scratch);
    if (after_scan) cudaEventRecord(after_scan, st);
    k_pack_write<QT, T><<<g2, kPackThreads, 0, st>>>(q, n, sym_min, zero_sym, len, code, wlo, nstates, vec, nchunks, bit_off,
                                                     zero_off, out_words, unpred_tmp, unpred_out);
}

#define SZ3B_INST_Q(QT)                                                                                              \
    template void launch_histogram<QT>(const QT *, uint64_t, int, int, int, unsigned long long *, cudaStream_t);    \
    template void launch_minmax_int<QT>(const QT *, uint64_t, int *, cudaStream_t);
SZ3B_INST_Q(uint16_t)
SZ3B_INST_Q(uint32_t)
SZ3B_INST_Q(int32_t)
#define SZ3B_INST_P(QT, T)                                                                                           \
    template void launch_pack<QT, T>(const QT *, uint64_t, int, int, const uint8_t *, const unsigned long long *,    \
                                     unsigned, int, unsigned *, unsigned *, unsigned long long *,                   \
                                     unsigned long long *, unsigned *, const T *, T *, cudaStream_t, cudaEvent_t,   \
                                     unsigned long long *);
SZ3B_INST_P(uint16_t, float)
SZ3B_INST_P(uint16_t, double)
SZ3B_INST_P(uint32_t, float)
SZ3B_INST_P(uint32_t, double)
SZ3B_INST_P(int32_t, float)
SZ3B_INST_P(int32_t, double)
SZ3B_INST_P(uint16_t, int32_t)
SZ3B_INST_P(uint16_t, int64_t)
SZ3B_INST_P(uint32_t, int32_t)
SZ3B_INST_P(uint32_t, int64_t)
SZ3B_INST_P(int32_t, int32_t)
SZ3B_INST_P(int32_t, int64_t)

}  // namespace sz3b
