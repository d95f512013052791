// sz3_b200/csrc/huffman_decode.cu -- GPU decoder for the reference's Huffman bitstream (HuffmanEncoder::decode,
// reference include/SZ3/encoder/HuffmanEncoder.hpp:225-255): MSB-first concatenated codes, no restart markers.
//
// The stream is cut into subsequences of kSubBits bits, one thread each.  Where a subsequence's first codeword starts
// is unknown, but prefix codes self-synchronise: a decoder started at a wrong bit offset falls onto true codeword
// boundaries after a few symbols.  So (after Weissenberger & Schmidt, ICPP 2018, restated for this format):
//
//   k_hd_sync   thread i decodes from (i*kSubBits + over[i-1]) to the first codeword boundary at or past
//               (i+1)*kSubBits and publishes that overshoot and its symbol count.  Iterated until no overshoot
//               changes: then every start is the true boundary (thread 0's start is exact,
//               and a thread whose start is exact publishes an exact overshoot -- induction over i; the loop ends
//               only at a fixed point, which is therefore the true one).  2-3 iterations in practice; a stretch where
//               the wrong phase happens to decode consistently (e.g. a run of one 2-bit codeword whose shifted reading
//               is another codeword) needs one round per subsequence of the stretch, so after the first round only the
//               subsequences whose start moved are decoded again (a work list: the launch has one thread per moved
//               predecessor) and the loop runs until nothing moves.
//   (scan)      exclusive prefix sum of the symbol counts -> output offsets (encode_kernels.cu: k_pack_scan)
//   k_hd_write  same decode, symbols written at their offsets, clipped to the stream's symbol count.
//
// Codes are resolved with a first-level table of kLutBits bits in shared memory (entry = length << 24 | node); longer
// codes continue down the tree arrays in global memory.
#include <cuda_runtime.h>

#include "launch.hpp"

namespace sz3b {

constexpr int kSubBits = 1024;
constexpr int kHdThreads = 256;
constexpr int kHdLutBits = 12;

struct HdTables {
    const uint32_t *lut;     // 1 << kHdLutBits entries (HuffmanDecoder::dlut)
    const uint32_t *lut2;    // second level (HuffmanDecoder::lut2)
    const uint32_t *L, *R;   // tree links (node count entries)
    const int *C;            // symbol (state) per node
    const uint8_t *leaf;
    int offset;              // HuffmanEncoder::offset
};

// MSB-first bit reader over big-endian 32-bit words.  `words` is 4-byte aligned and the stream starts `shift` bits
// (0, 8, 16 or 24) into it -- the stream sits wherever the assembled file put it; readable zero padding follows it.
struct BitReader {
    const uint32_t *words;
    uint64_t acc;      // next bits, left aligned
    int have;          // valid bits in acc
    uint64_t next_w;   // next word index to load
    __device__ __forceinline__ void init(const uint32_t *w, uint64_t bitpos) {
        words = w;
        next_w = bitpos >> 5;
        const unsigned sh = static_cast<unsigned>(bitpos & 31);
        const uint64_t w0 = __byte_perm(words[next_w], 0, 0x0123), w1 = __byte_perm(words[next_w + 1], 0, 0x0123);
        acc = ((w0 << 32) | w1) << sh;
        have = 64 - static_cast<int>(sh);
        next_w += 2;
    }
    __device__ __forceinline__ void refill() {
        if (have <= 32) {
            const uint64_t w = __byte_perm(words[next_w++], 0, 0x0123);
            acc |= w << (32 - have);
            have += 32;
        }
    }
    __device__ __forceinline__ void skip(int n) {
        acc <<= n;
        have -= n;
    }
};

// The rest of a code longer than the first-level table, the reader standing after its first kHdLutBits bits: one
// lookup in the second-level table (entry e of the first level: bit 31, S = bits 24..30, base = (bits 0..23) << 4)
// resolves codes of up to kHdLutBits + S bits; anything longer, and prefixes without a sub-table, walk the tree bit
// by bit from where the tables leave off.  Returns the leaf, *len_out = total code length.
__device__ __forceinline__ uint32_t hd_long_code(BitReader &br, uint32_t e, const HdTables &t, uint32_t *len_out) {
    uint32_t node = e & 0xffffffu, len = kHdLutBits;
    bool walk = true;
    if (e & 0x80000000u) {
        const uint32_t S = (e >> 24) & 0x7fu;
        br.refill();
        const uint32_t e2 = t.lut2[(static_cast<size_t>(e & 0xffffffu) << 4) + static_cast<uint32_t>(br.acc >> (64 - S))];
        const uint32_t len2 = e2 >> 24;
        node = e2 & 0xffffffu;
        br.skip(static_cast<int>(len2 ? len2 : S));
        len += len2 ? len2 : S;
        walk = len2 == 0;
    }
    while (walk) {
        br.refill();
        node = (br.acc >> 63) ? t.R[node] : t.L[node];
        br.skip(1);
        len++;
        if (t.leaf[node] || len >= 96) break;
    }
    *len_out = len;
    return node;
}

// decodes one symbol; returns its node (leaf) and advances the reader; *len_out = code length
__device__ __forceinline__ uint32_t hd_symbol(BitReader &br, const uint32_t *slut, const HdTables &t, int *len_out) {
    br.refill();
    const uint32_t e = slut[br.acc >> (64 - kHdLutBits)];
    uint32_t len = e >> 24, node = e & 0xffffffu;
    if (len - 1u < static_cast<uint32_t>(kHdLutBits)) {
        br.skip(static_cast<int>(len));
    } else {   // code longer than the table
        br.skip(kHdLutBits);
        node = hd_long_code(br, e, t, &len);
    }
    *len_out = static_cast<int>(len);
    return node;
}

// One synchronisation round.  Round 1 (list_in == nullptr) decodes every subsequence; later rounds decode the
// successors of the subsequences whose overshoot moved in the previous round (list_in, n_in entries).  `over` is updated
// in place: a reader may see its predecessor's overshoot of this or of the previous round, but a predecessor that moves
// is listed, so its successor runs again in the next round and the loop can only stop at the fixed point.
__global__ void __launch_bounds__(kHdThreads) k_hd_sync(const uint32_t *__restrict__ words, unsigned shift, uint64_t total_bits,
                                                       uint64_t nsub, HdTables t, uint8_t *over, const uint32_t *__restrict__ list_in,
                                                       uint64_t n_in, uint32_t *__restrict__ list_out,
                                                       unsigned *__restrict__ counts, unsigned long long *__restrict__ n_out) {
    __shared__ uint32_t slut[1 << kHdLutBits];
    for (int j = threadIdx.x; j < (1 << kHdLutBits); j += blockDim.x) slut[j] = t.lut[j];
    __syncthreads();
    const uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    uint64_t i;
    if (list_in) {
        if (k >= n_in) return;
        i = static_cast<uint64_t>(list_in[k]) + 1;
    } else {
        i = k;
    }
    if (i >= nsub) return;
    uint64_t pos = i * kSubBits + (i ? over[i - 1] : 0);
    uint64_t limit = (i + 1) * kSubBits;
    if (limit > total_bits) limit = total_bits;
    unsigned cnt = 0;
    if (pos < limit) {
        BitReader br;
        br.init(words, pos + shift);
        while (pos < limit) {
            int len;
            hd_symbol(br, slut, t, &len);
            pos += len;
            cnt++;
        }
    }
    const uint64_t ov = pos > (i + 1) * kSubBits ? pos - (i + 1) * kSubBits : 0;
    const uint8_t o = static_cast<uint8_t>(ov > 255 ? 255 : ov);
    counts[i] = cnt;
    if (o != over[i]) {
        over[i] = o;
        list_out[atomicAdd(n_out, 1ull)] = static_cast<uint32_t>(i);
    }
}

// Final pass: every subsequence is decoded once more from its exact start and its symbols are written.  A thread's
// symbols are contiguous in the output (offs[i] onwards), so they leave in groups of 16 bytes (8 or 4 symbols gathered
// in two 64-bit registers, aligned stores; scalar stores up to the first aligned position and for the tail): one
// 16-byte store per eight symbols instead of eight 2-byte stores that each dirty a sector of their own.  Codes inside
// the first-level table take their symbol from a table in shared memory as well (no node -> symbol load from global).
template <class QT>
__global__ void __launch_bounds__(kHdThreads) k_hd_write(const uint32_t *__restrict__ words, unsigned shift, uint64_t total_bits,
                                                        uint64_t nsub, HdTables t, const uint8_t *__restrict__ over,
                                                        const unsigned long long *__restrict__ offs, uint64_t n,
                                                        QT *__restrict__ out) {
    __shared__ uint32_t slut[1 << kHdLutBits];
    __shared__ QT ssym[1 << kHdLutBits];
    for (int k = threadIdx.x; k < (1 << kHdLutBits); k += blockDim.x) {
        const uint32_t e = t.lut[k];
        slut[k] = e;
        ssym[k] = ((e >> 24) - 1u < static_cast<uint32_t>(kHdLutBits)) ? static_cast<QT>(t.C[e & 0xffffffu] + t.offset) : static_cast<QT>(0);
    }
    __syncthreads();
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nsub) return;
    uint64_t pos = i * kSubBits + (i ? over[i - 1] : 0);
    uint64_t limit = (i + 1) * kSubBits;
    if (limit > total_bits) limit = total_bits;
    uint64_t o = offs[i];
    if (pos >= limit) return;
    constexpr int kSymBits = 8 * static_cast<int>(sizeof(QT));
    constexpr int kPerWord = 64 / kSymBits;   // symbols per 64-bit register: 4 or 2
    constexpr int kGroup = 2 * kPerWord;      // symbols per 16-byte store
    unsigned long long lo = 0, hi = 0;
    int nb = 0;   // symbols gathered; the group starts at out[o]
    BitReader br;
    br.init(words, pos + shift);
    while (pos < limit && o + nb < n) {
        br.refill();
        const uint32_t idx = static_cast<uint32_t>(br.acc >> (64 - kHdLutBits));
        const uint32_t e = slut[idx];
        uint32_t len = e >> 24;
        unsigned long long sym;
        if (len - 1u < static_cast<uint32_t>(kHdLutBits)) {
            sym = ssym[idx];
            br.skip(static_cast<int>(len));
        } else {   // code longer than the table
            br.skip(kHdLutBits);
            const uint32_t node = hd_long_code(br, e, t, &len);
            sym = static_cast<QT>(t.C[node] + t.offset);
        }
        pos += len;
        if (nb == 0 && (o & (kGroup - 1)) != 0) {   // not yet at a 16-byte boundary of the output
            out[o++] = static_cast<QT>(sym);
            continue;
        }
        if (nb < kPerWord)
            lo |= sym << (kSymBits * nb);
        else
            hi |= sym << (kSymBits * (nb - kPerWord));
        if (++nb == kGroup) {
            *reinterpret_cast<ulonglong2 *>(out + o) = make_ulonglong2(lo, hi);
            o += kGroup;
            nb = 0;
            lo = hi = 0;
        }
    }
    for (int k = 0; k < nb; k++) {
        const unsigned long long w = k < kPerWord ? lo >> (kSymBits * k) : hi >> (kSymBits * (k - kPerWord));
        out[o + k] = static_cast<QT>(w);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
uint64_t hd_num_sub(uint64_t total_bits) { return (total_bits + kSubBits - 1) / kSubBits; }

void launch_hd_sync(const uint32_t *words, unsigned shift, uint64_t total_bits, const HdDeviceTables &tb, uint8_t *over, const uint32_t *list_in,
                    uint64_t n_in, uint32_t *list_out, unsigned *counts, unsigned long long *n_out, cudaStream_t st) {
    const uint64_t nsub = hd_num_sub(total_bits);
    const uint64_t nthreads = list_in ? n_in : nsub;
    if (nthreads == 0) return;
    HdTables t{tb.lut, tb.lut2, tb.L, tb.R, tb.C, tb.leaf, tb.offset};
    k_hd_sync<<<static_cast<unsigned>((nthreads + kHdThreads - 1) / kHdThreads), kHdThreads, 0, st>>>(words, shift, total_bits, nsub, t,
                                                                                                 over, list_in, n_in, list_out, counts, n_out);
}

template <class QT>
void launch_hd_write(const uint32_t *words, unsigned shift, uint64_t total_bits, const HdDeviceTables &tb, const uint8_t *over,
                     const unsigned long long *offs, uint64_t n, QT *out, cudaStream_t st) {
    const uint64_t nsub = hd_num_sub(total_bits);
    HdTables t{tb.lut, tb.lut2, tb.L, tb.R, tb.C, tb.leaf, tb.offset};
    k_hd_write<QT><<<static_cast<unsigned>((nsub + kHdThreads - 1) / kHdThreads), kHdThreads, 0, st>>>(words, shift, total_bits, nsub, t,
                                                                                                  over, offs, n, out);
}
template void launch_hd_write<uint16_t>(const uint32_t *, unsigned, uint64_t, const HdDeviceTables &, const uint8_t *,
                                        const unsigned long long *, uint64_t, uint16_t *, cudaStream_t);
template void launch_hd_write<uint32_t>(const uint32_t *, unsigned, uint64_t, const HdDeviceTables &, const uint8_t *,
                                        const unsigned long long *, uint64_t, uint32_t *, cudaStream_t);

}  // namespace sz3b
