// sz3_b200/csrc/zhuf.cuh -- the GPU lossless stage ("zhuf"): standard zstd frames whose blocks hold only
// Huffman-coded literals (no sequences), so that the unmodified reference decoder (Lossless_zstd::decompress ->
// ZSTD_decompress, reference include/SZ3/lossless/Lossless_zstd.hpp:39-45) reads them like any other zstd stream.
//
// Why: the bytes that reach the lossless stage are the output of HuffmanEncoder.  zstd finds no matches in them; its
// whole gain (about 1 %) is the order-0 entropy coding of the literals, and that costs 7 ms of host time per 46 MB at
// level 3 on 16 cores.  Here every 128 KiB block gets its own byte histogram, length-limited Huffman code, tree
// description and 4-stream bit packing on the GPU; the host only learns the total size.  Per-block tables follow the
// local statistics of the index stream (coarse levels first, then the fine level), which gains more than zstd -3 does.
//
// Format facts used (zstd compression format, sections Frame_Header, Block_Header, Literals_Section,
// Huffman_Tree_Description, FSE_Table_Description), checked against libzstd's decoder by tests/test_zhuf.py:
//   * frame   = magic 0xFD2FB528 | FHD 0xA0 (single segment, 4-byte content size) | content size | blocks
//   * block   = 3-byte header (last | type << 1 | size << 3) | payload;  type 0 raw, 2 compressed
//   * literals-only compressed block = literals section | 0x00 (zero sequences)
//   * literals section (type 2, 4 streams, size format 3) = 5-byte header | tree description | 6-byte jump table |
//     4 streams; a stream holds its symbols LAST symbol first, codes appended LSB-first, closed by a single 1 bit
//   * tree description = FSE-compressed weights (accuracy log 6, two interleaved states) or 4-bit direct weights;
//     weight = max_bits + 1 - length, the last present symbol's weight is implied
//   * code of a symbol: canonical, longest codes first, symbols ascending inside a length
//
// Plain inline code shared by the kernels (zhuf_kernels.cu) and the test-only sequential encoder (tests/emul).
#pragma once
#include "core.cuh"

namespace sz3b {

constexpr int kZhufMaxBits = 11;               // Huffman literals: Max_Number_of_Bits
constexpr uint32_t kZhufBlock = 128u << 10;    // Block_Maximum_Size
constexpr uint32_t kZhufFrame = 1u << 20;      // source bytes per frame (8 blocks): frames decode in parallel
constexpr uint32_t kZhufBlocksPerFrame = kZhufFrame / kZhufBlock;
constexpr uint32_t kZhufMinCoded = 1024;       // shorter blocks are stored raw
constexpr uint32_t kZhufFrameHeader = 9;
constexpr uint32_t kZhufDescCap = 136;

struct ZhufBlockInfo {
    uint32_t sym[256];          // len << 16 | code, 0 for absent symbols
    uint8_t desc[kZhufDescCap]; // Huffman_Tree_Description
    uint32_t desc_len;
    uint32_t ok;                // a usable code exists
    uint32_t sb[4];             // bytes of the 4 streams (closing bit included)
    uint32_t coded;             // decided by the scan: 1 = compressed block, 0 = raw block
    uint32_t pad_;
    unsigned long long off;     // offset of the block header in the output
};

// worst-case output size for `len` source bytes (every block raw + headers)
SZ_HD size_t zhuf_bound(size_t len) {
    const size_t frames = (len + kZhufFrame - 1) / kZhufFrame + 1;
    const size_t blocks = (len + kZhufBlock - 1) / kZhufBlock + frames;
    return len + frames * kZhufFrameHeader + blocks * 3 + 64;
}
SZ_HD uint64_t zhuf_num_blocks(uint64_t len) { return (len + kZhufBlock - 1) / kZhufBlock; }
SZ_HD uint32_t zhuf_block_len(uint64_t len, uint64_t g) {
    const uint64_t a = g * kZhufBlock;
    return static_cast<uint32_t>(len - a < kZhufBlock ? len - a : kZhufBlock);
}
// source range of stream s (0..3) of block g
SZ_HD void zhuf_stream_range(uint64_t len, uint64_t g, int s, uint64_t *a, uint64_t *b) {
    const uint32_t bl = zhuf_block_len(len, g);
    const uint32_t seg = (bl + 3) / 4;
    const uint32_t lo = static_cast<uint32_t>(s) * seg < bl ? static_cast<uint32_t>(s) * seg : bl;
    const uint32_t hi = s == 3 ? bl : (lo + seg < bl ? lo + seg : bl);
    *a = g * kZhufBlock + lo;
    *b = g * kZhufBlock + hi;
}
// payload bytes of block g and whether it is worth coding
SZ_HD uint32_t zhuf_block_payload(uint32_t bl, const ZhufBlockInfo &bi, bool *coded) {
    const uint64_t lit = static_cast<uint64_t>(bi.desc_len) + 6 + bi.sb[0] + bi.sb[1] + bi.sb[2] + bi.sb[3];
    const uint64_t payload = 5 + lit + 1;
    *coded = bi.ok && bl >= kZhufMinCoded && payload < bl && lit < (1u << 18);
    return *coded ? static_cast<uint32_t>(payload) : bl;
}

SZ_HD void zhuf_put_le(uint8_t *p, uint64_t v, int nbytes) {
    for (int i = 0; i < nbytes; i++) p[i] = static_cast<uint8_t>(v >> (8 * i));
}

// Headers of block g (block header at out + bi.off): frame header (first block of a frame), block header, literals
// header, tree description, jump table, closing "0 sequences" byte.  Returns the offset of stream 0 / the raw bytes.
SZ_HD uint64_t zhuf_write_headers(uint8_t *out, uint64_t len, uint64_t g, const ZhufBlockInfo &bi) {
    const uint64_t nblocks = zhuf_num_blocks(len);
    const uint64_t f = g / kZhufBlocksPerFrame, g0 = f * kZhufBlocksPerFrame;
    const bool last = g + 1 == nblocks || g + 1 == g0 + kZhufBlocksPerFrame;
    const uint32_t bl = zhuf_block_len(len, g);
    const uint64_t bo = bi.off;
    if (g == g0) {
        const uint64_t fa = f * kZhufFrame;
        const uint64_t flen = len - fa < kZhufFrame ? len - fa : kZhufFrame;
        uint8_t *h = out + bo - kZhufFrameHeader;
        zhuf_put_le(h, 0xFD2FB528u, 4);
        h[4] = 0xA0;   // Frame_Content_Size_flag 2 (4 bytes), Single_Segment_flag
        zhuf_put_le(h + 5, flen, 4);
    }
    if (!bi.coded) {
        zhuf_put_le(out + bo, (last ? 1u : 0u) | (0u << 1) | (static_cast<uint64_t>(bl) << 3), 3);
        return bo + 3;
    }
    const uint64_t lit = static_cast<uint64_t>(bi.desc_len) + 6 + bi.sb[0] + bi.sb[1] + bi.sb[2] + bi.sb[3];
    const uint64_t payload = 5 + lit + 1;
    zhuf_put_le(out + bo, (last ? 1u : 0u) | (2u << 1) | (payload << 3), 3);
    uint8_t *p = out + bo + 3;
    // Literals_Section_Header, size format 3: type(2) | 3(2) | regenerated size(18) | compressed size(18)
    zhuf_put_le(p, 2u | (3u << 2) | (static_cast<uint64_t>(bl) << 4) | (lit << 22), 5);
    p += 5;
    for (uint32_t i = 0; i < bi.desc_len; i++) p[i] = bi.desc[i];
    p += bi.desc_len;
    zhuf_put_le(p, bi.sb[0], 2);
    zhuf_put_le(p + 2, bi.sb[1], 2);
    zhuf_put_le(p + 4, bi.sb[2], 2);
    out[bo + 3 + payload - 1] = 0;   // Sequences_Section: 0 sequences
    return bo + 3 + 5 + bi.desc_len + 6;
}

#if defined(__CUDA_ARCH__)
#define SZ_CTA_SYNC() __syncthreads()
#define SZ_SMEM_INC(p) atomicAdd((p), 1u)
#else
#define SZ_CTA_SYNC() ((void)0)
#define SZ_SMEM_INC(p) (++*(p))
#endif

// ---------------------------------------------------------------------------------------------------------------------
// Huffman code construction (one block = at most 131072 symbols over 256 byte values)
// ---------------------------------------------------------------------------------------------------------------------
struct ZhufMkScratch {   // shared memory of the cooperative length computation
    uint16_t p[2][256], d[2][256];
    uint32_t cnt[256], cum[256];
    int maxdep;
};

// Minimum-redundancy code lengths (Moffat & Katajainen, in place): A[0..n) ascending frequencies in, code lengths out
// (A[0] = the rarest symbol's = the longest), n >= 2.  Called by all nt threads of the CTA.
//   phase 1 (thread 0): the two-queue merge; the heads of both queues are kept in registers, so an iteration costs
//            one or two shared-memory loads instead of four.  Leaves A[i], i < n - 2, then hold parent indices.
//   phase 2 (all): depth of every internal node by pointer jumping (8 rounds cover any tree over 256 leaves).
//   phase 3 (all): internal nodes per depth -> leaves per depth (the deepest leaves go to the rarest symbols).
SZ_HD void zhuf_mk_lengths(uint32_t *A, int n, ZhufMkScratch &M, int tid, int nt) {
    if (tid == 0) {
        A[0] += A[1];
        int root = 0, leaf = 2;
        uint32_t vroot = A[0], vleaf = n > 2 ? A[2] : 0u;   // heads of the internal-node and the leaf queue
        for (int next = 1; next < n - 1; next++) {
            uint32_t sum;
            // first item of the pair
            if (leaf >= n || vroot < vleaf) {
                sum = vroot;
                A[root++] = static_cast<uint32_t>(next);
                vroot = A[root];   // nodes root..next-1 hold weights (next itself is written below, root < next here
                                   // unless the queue is empty, in which case the value is not used before it is set)
            } else {
                sum = vleaf;
                leaf++;
                vleaf = leaf < n ? A[leaf] : 0u;
            }
            // second item
            if (leaf >= n || (root < next && vroot < vleaf)) {
                sum += vroot;
                A[root++] = static_cast<uint32_t>(next);
                vroot = root < next ? A[root] : 0u;
            } else {
                sum += vleaf;
                leaf++;
                vleaf = leaf < n ? A[leaf] : 0u;
            }
            A[next] = sum;
            if (root == next) vroot = sum;   // the queue of internal nodes was empty: the new node is its head
        }
    }
    SZ_CTA_SYNC();
    // internal nodes 0 .. n-2, root = n-2
    const int ni = n - 1;
    for (int i = tid; i < ni; i += nt) {
        const bool is_root = i == ni - 1;
        M.p[0][i] = static_cast<uint16_t>(is_root ? i : A[i]);
        M.d[0][i] = is_root ? 0 : 1;
    }
    for (int i = tid; i < 256; i += nt) M.cnt[i] = 0;
    SZ_CTA_SYNC();
    int cur = 0;
    for (int r = 0; r < 8; r++) {
        for (int i = tid; i < ni; i += nt) {
            const int pi = M.p[cur][i];
            M.d[cur ^ 1][i] = static_cast<uint16_t>(M.d[cur][i] + M.d[cur][pi]);
            M.p[cur ^ 1][i] = M.p[cur][pi];
        }
        cur ^= 1;
        SZ_CTA_SYNC();
    }
    for (int i = tid; i < ni; i += nt) SZ_SMEM_INC(&M.cnt[M.d[cur][i]]);
    SZ_CTA_SYNC();
    if (tid == 0) {
        uint32_t avbl = 1, cum = 0;
        int dep = 0;
        for (; dep < 256; dep++) {
            const uint32_t used = M.cnt[dep];
            cum += avbl - used;   // leaves at this depth
            M.cum[dep] = cum;
            avbl = 2 * used;
            if (avbl == 0) break;
        }
        M.maxdep = dep < 256 ? dep : 255;
    }
    SZ_CTA_SYNC();
    for (int i = tid; i < n; i += nt) {
        const uint32_t pos = static_cast<uint32_t>(n - 1 - i);   // rank from the most frequent leaf
        int dep = 0;
        while (dep < M.maxdep && M.cum[dep] <= pos) dep++;
        A[i] = static_cast<uint32_t>(dep);
    }
    SZ_CTA_SYNC();
}

struct ZhufBits {   // forward bit writer, LSB first
    uint8_t *p;
    uint32_t cap, nbytes;
    unsigned long long acc;
    int nacc;
    bool overflow;
};
SZ_HD void zhuf_bits_init(ZhufBits &w, uint8_t *buf, uint32_t cap) {
    w.p = buf;
    w.cap = cap;
    w.nbytes = 0;
    w.acc = 0;
    w.nacc = 0;
    w.overflow = false;
}
SZ_HD void zhuf_bits_put(ZhufBits &w, uint32_t v, int n) {
    w.acc |= static_cast<unsigned long long>(v & ((1u << n) - 1u)) << w.nacc;
    w.nacc += n;
    while (w.nacc >= 8) {
        if (w.nbytes < w.cap)
            w.p[w.nbytes++] = static_cast<uint8_t>(w.acc);
        else
            w.overflow = true;
        w.acc >>= 8;
        w.nacc -= 8;
    }
}
// bytes written once the last partial byte is flushed (zero padded)
SZ_HD uint32_t zhuf_bits_close(ZhufBits &w) {
    if (w.nacc > 0) {
        if (w.nbytes < w.cap)
            w.p[w.nbytes++] = static_cast<uint8_t>(w.acc);
        else
            w.overflow = true;
        w.acc = 0;
        w.nacc = 0;
    }
    return w.nbytes;
}

struct ZhufFseScratch {   // tables of the weight coder (shared memory on the device: one thread uses them)
    int count[16], norm[16];
    int delta_nb[16], delta_find[16];   // FSE symbol transforms: nbBits = (state + delta_nb) >> 16, next = tab[(state >> nbBits) + delta_find]
    uint8_t table_sym[64], state_tab[64];
};

// FSE-compressed weights: FSE_Table_Description (normalized counts, accuracy log 6) followed by the two-state backward
// bitstream.  F.count[] holds the histogram of the n weights on entry.  Returns the bytes written, 0 when this
// representation is not possible.
SZ_HD uint32_t zhuf_fse_weights(const uint8_t *w, int n, ZhufFseScratch &F, uint8_t *out, uint32_t cap) {
    constexpr int kLog = 6, kSize = 1 << kLog;
    if (n < 2) return 0;
    int max_sym = 0;
    for (int s = 0; s < 16; s++) {
        F.norm[s] = 0;
        if (F.count[s]) max_sym = s;
    }
    // normalized counts: every present symbol at least 1, total kSize, no symbol owning the whole table
    int total = 0, present = 0;
    for (int s = 0; s <= max_sym; s++)
        if (F.count[s]) {
            const int v = (F.count[s] * kSize + n / 2) / n;
            F.norm[s] = v < 1 ? 1 : v;
            total += F.norm[s];
            present++;
        }
    if (present < 2) return 0;   // one repeated weight has no FSE form
    while (total != kSize) {
        int best = -1;
        for (int s = 0; s <= max_sym; s++)
            if (F.norm[s] > (total > kSize ? 1 : 0) && (best < 0 || F.norm[s] > F.norm[best])) best = s;
        if (best < 0) return 0;
        if (total > kSize) {
            F.norm[best]--;
            total--;
        } else {
            F.norm[best]++;
            total++;
        }
    }
    for (int s = 0; s <= max_sym; s++)
        if (F.norm[s] >= kSize) return 0;
    // ---- FSE_Table_Description
    ZhufBits hw;
    zhuf_bits_init(hw, out, cap);
    zhuf_bits_put(hw, kLog - 5, 4);
    {
        int remaining = kSize + 1, threshold = kSize, nbits = kLog + 1;
        int s = 0;
        while (remaining > 1 && s <= max_sym) {
            const int c = F.norm[s++];
            const int max = (2 * threshold - 1) - remaining;
            remaining -= c;
            int v = c + 1;
            if (v >= threshold) v += max;
            zhuf_bits_put(hw, static_cast<uint32_t>(v), nbits - (v < max ? 1 : 0));
            if (c == 0) {   // further zero-probability symbols: 2-bit repeat codes, 3 = "three more and continue"
                int run = 0;
                while (s <= max_sym && F.norm[s] == 0) {
                    run++;
                    s++;
                }
                while (run >= 3) {
                    zhuf_bits_put(hw, 3, 2);
                    run -= 3;
                }
                zhuf_bits_put(hw, static_cast<uint32_t>(run), 2);
            }
            while (remaining < threshold) {
                nbits--;
                threshold >>= 1;
            }
        }
        if (remaining != 1) return 0;
    }
    const uint32_t hdr = zhuf_bits_close(hw);
    if (hw.overflow) return 0;
    // ---- the decoder's table: symbols spread with step (size/2 + size/8 + 3); a symbol's k-th cell in table order is
    //      its sub-state k (decoder: nextState = count + k).  The encoder needs the inverse: state_tab holds the table
    //      indices grouped by symbol in that order, delta_find[s] = (start of the group) - count, so that the cell for
    //      sub-state (state >> nbBits) is state_tab[(state >> nbBits) + delta_find[s]].
    {
        const int step = (kSize >> 1) + (kSize >> 3) + 3;
        int pos = 0;
        for (int s = 0; s <= max_sym; s++)
            for (int i = 0; i < F.norm[s]; i++) {
                F.table_sym[pos] = static_cast<uint8_t>(s);
                pos = (pos + step) & (kSize - 1);
            }
        if (pos != 0) return 0;
        int at = 0;
        int fill[16];
        for (int s = 0; s <= max_sym; s++) {
            const int c = F.norm[s];
            fill[s] = at;
            F.delta_find[s] = at - c;
            if (c) {
                int hb = 0;   // highbit(c - 1), 0 for c == 1
                for (int v = c - 1; v > 1; v >>= 1) hb++;
                const int max_bits = c == 1 ? kLog : kLog - hb;
                F.delta_nb[s] = (max_bits << 16) - (c << max_bits);
            }
            at += c;
        }
        for (int u = 0; u < kSize; u++) F.state_tab[fill[F.table_sym[u]]++] = static_cast<uint8_t>(u);
    }
    // ---- two interleaved states, symbols taken from the last to the first; even positions use state 0.  A state
    //      starts on the smallest sub-state of its symbol: the decoder then reads the most bits for it (at least one,
    //      as no count reaches the table size), which is how it detects the end of the stream on the last two symbols.
    ZhufBits bw;
    zhuf_bits_init(bw, out + hdr, cap - hdr);
    int X0 = 0, X1 = 0;   // states of the even / odd positions
    {
        const int sa = w[n - 1], sb = w[n - 2];
        const int xa = kSize + F.state_tab[F.norm[sa] + F.delta_find[sa]], xb = kSize + F.state_tab[F.norm[sb] + F.delta_find[sb]];
        if ((n - 1) & 1) {
            X1 = xa;
            X0 = xb;
        } else {
            X0 = xa;
            X1 = xb;
        }
    }
    for (int i = n - 3; i >= 0; i--) {
        const int s = w[i];
        const int X = (i & 1) ? X1 : X0;
        const int nb = (X + F.delta_nb[s]) >> 16;
        zhuf_bits_put(bw, static_cast<uint32_t>(X), nb);
        const int nx = kSize + F.state_tab[(X >> nb) + F.delta_find[s]];
        if (i & 1)
            X1 = nx;
        else
            X0 = nx;
    }
    zhuf_bits_put(bw, static_cast<uint32_t>(X1 - kSize), kLog);
    zhuf_bits_put(bw, static_cast<uint32_t>(X0 - kSize), kLog);
    zhuf_bits_put(bw, 1, 1);
    const uint32_t body = zhuf_bits_close(bw);
    if (bw.overflow) return 0;
    return hdr + body;
}

struct ZhufScratch {   // shared memory of the building CTA
    uint32_t work[256];
    uint8_t len[256], w[256];
    uint32_t per_rank[kZhufMaxBits + 2], val[kZhufMaxBits + 2];
    int longest, last, desc_ok;
    ZhufFseScratch fse;
    ZhufMkScratch mk;
};

// Table of one block from its byte histogram, by the tid-th of nt threads of one CTA (all of them call this; the
// tests' sequential twin runs it with nt = 1).  sorted_sym / sorted_freq: the present symbols in ascending
// (frequency, symbol) order, n of them.  The two inherently serial pieces -- the Moffat-Katajainen length computation
// and the FSE state chain over the weights -- run on thread 0; the per-symbol loops are spread over the CTA.
SZ_HD void zhuf_build_table(const uint8_t *sorted_sym, uint32_t *sorted_freq, int n, ZhufScratch &S, ZhufBlockInfo &t, int tid,
                            int nt) {
    for (int s = tid; s < 256; s += nt) {
        t.sym[s] = 0;
        S.len[s] = 0;
    }
    for (int b = tid; b < kZhufMaxBits + 2; b += nt) S.per_rank[b] = 0;
    for (int b = tid; b < 16; b += nt) S.fse.count[b] = 0;
    if (tid == 0) {
        t.desc_len = 0;
        t.ok = 0;
        S.desc_ok = 0;
    }
    SZ_CTA_SYNC();
    if (n < 2) return;
    for (;;) {
        for (int i = tid; i < n; i += nt) S.work[i] = sorted_freq[i];
        SZ_CTA_SYNC();
        zhuf_mk_lengths(S.work, n, S.mk, tid, nt);
        if (tid == 0) S.longest = static_cast<int>(S.work[0]);
        SZ_CTA_SYNC();
        if (S.longest <= kZhufMaxBits) break;
        // flatten the histogram until the code fits; halving keeps the order, the result stays a complete code
        for (int i = tid; i < n; i += nt) sorted_freq[i] = (sorted_freq[i] + 1) / 2;
        SZ_CTA_SYNC();
    }
    for (int i = tid; i < n; i += nt) S.len[sorted_sym[i]] = static_cast<uint8_t>(S.work[i]);
    SZ_CTA_SYNC();
    if (tid == 0) {
        int last = 255;
        while (last >= 0 && S.len[last] == 0) last--;
        S.last = last;
    }
    SZ_CTA_SYNC();
    const int last = S.last, longest = S.longest;
    if (last < 1) return;
    for (int s = tid; s < last; s += nt) {
        const uint8_t wv = S.len[s] ? static_cast<uint8_t>(longest + 1 - S.len[s]) : 0;
        S.w[s] = wv;
#if defined(__CUDA_ARCH__)
        atomicAdd(&S.fse.count[wv], 1);
#else
        S.fse.count[wv]++;
#endif
    }
    for (int s = tid; s < 256; s += nt)
        if (S.len[s]) SZ_SMEM_INC(&S.per_rank[S.len[s]]);
    SZ_CTA_SYNC();
    if (tid == 0) {
        const int nw = last;   // explicit weights
        const uint32_t fse = zhuf_fse_weights(S.w, nw, S.fse, t.desc + 1, kZhufDescCap - 1);
        if (fse > 0 && fse < 128) {
            t.desc[0] = static_cast<uint8_t>(fse);
            t.desc_len = fse + 1;
            S.desc_ok = 1;
        } else if (nw <= 128) {   // direct representation: 4 bits per weight
            t.desc[0] = static_cast<uint8_t>(127 + nw);
            for (int i = 0; i < nw; i += 2)
                t.desc[1 + i / 2] = static_cast<uint8_t>((S.w[i] << 4) | (i + 1 < nw ? S.w[i + 1] : 0));
            t.desc_len = static_cast<uint32_t>(1 + (nw + 1) / 2);
            S.desc_ok = 1;
        }
        // canonical codes: longest codes first, symbols ascending inside a length
        uint32_t min = 0;
        for (int b = longest; b >= 1; b--) {
            S.val[b] = min;
            min = (min + S.per_rank[b]) >> 1;
        }
    }
    SZ_CTA_SYNC();
    if (!S.desc_ok) return;
    for (int s = tid; s < 256; s += nt) {
        const uint32_t l = S.len[s];
        if (!l) continue;
        uint32_t before = 0;   // symbols of the same length below s
        for (int j = 0; j < s; j++) before += S.len[j] == l ? 1u : 0u;
        t.sym[s] = (l << 16) | (S.val[l] + before);
    }
    if (tid == 0) t.ok = 1;
    SZ_CTA_SYNC();
}

}  // namespace sz3b
