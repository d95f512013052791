// sz3_b200/csrc/zhuf_dec.cuh -- decoder of the GPU lossless stage's own frames (zhuf.cuh): standard zstd frames whose
// blocks are raw or hold only Huffman-coded literals in four streams.  Decompression otherwise hands every frame to
// libzstd on the host (Lossless_zstd::decompress, reference include/SZ3/lossless/Lossless_zstd.hpp:39-45: 3.4 ms for
// the 46 MB of a 512^3 stream on 16 threads, the largest stage of a decompression); frames of this shape need no
// match copying, so their blocks decode independently on the GPU -- one thread per stream.  Anything else in a stream
// (a frame written by real zstd: sequences, RLE or tree-less literals, other header forms) makes the walker say "not
// mine" and the host path runs as before.
//
// Format facts (zstd compression format; RFC 8878 sections 3.1.1, 4.2.1, 4.2.2): see the list in zhuf.cuh.  Decoding
// side: the weights of a Huffman_Tree_Description are either 4-bit direct or FSE-compressed (accuracy log <= 6, two
// interleaved states, backward bitstream closed by a 1 bit); Max_Number_of_Bits = highbit(sum 2^(w-1)) + 1 and the last
// symbol's weight fills the sum to the next power of two; a symbol of weight w owns 2^(w-1) consecutive cells of the
// 2^max decode table, weights ascending, symbols ascending inside a weight.
//
// Plain inline code shared by the kernel (zhuf_kernels.cu) and the test-only sequential decoder (tests/emul), which is
// checked against libzstd on the CPU.
#pragma once
#include <string.h>

#include "zhuf.cuh"

namespace sz3b {

struct ZhufDecBlock {
    unsigned long long src;    // offset (in the compressed payload) of the tree description, or of the bytes of a raw block
    unsigned long long dst;    // offset of the block's first byte in the decoded stream
    uint32_t regen;            // decoded bytes
    uint32_t coded;            // 0 = raw block
    uint32_t lit;              // compressed block: bytes of tree description + jump table + four streams
    uint32_t pad_;
};

SZ_HD uint32_t zhuf_get_le(const uint8_t *p, int nbytes) {
    uint32_t v = 0;
    for (int i = 0; i < nbytes; i++) v |= static_cast<uint32_t>(p[i]) << (8 * i);
    return v;
}

// Walks the frames of payload[0, size): fills blocks[0, *nblocks) and *raw_total.  Returns false when the payload is
// anything but a concatenation of frames of the shape zhuf writes (the caller then uses libzstd), or holds more than
// `cap` blocks.  Host code (also run by tests/emul).
inline bool zhuf_walk_frames(const uint8_t *p, size_t size, ZhufDecBlock *blocks, size_t cap, size_t *nblocks, uint64_t *raw_total) {
    size_t at = 0, nb = 0;
    uint64_t raw = 0;
    while (at < size) {
        if (size - at < kZhufFrameHeader) return false;
        if (zhuf_get_le(p + at, 4) != 0xFD2FB528u || p[at + 4] != 0xA0) return false;
        const uint64_t fcs = zhuf_get_le(p + at + 5, 4);
        at += kZhufFrameHeader;
        uint64_t got = 0;
        for (;;) {
            if (size - at < 3) return false;
            const uint32_t bh = zhuf_get_le(p + at, 3);
            const uint32_t last = bh & 1u, type = (bh >> 1) & 3u, bsize = bh >> 3;
            at += 3;
            if (nb >= cap) return false;
            ZhufDecBlock &b = blocks[nb];
            b.pad_ = 0;
            if (type == 0) {
                if (bsize > kZhufBlock || size - at < bsize) return false;
                b.src = at;
                b.dst = raw + got;
                b.regen = bsize;
                b.coded = 0;
                b.lit = 0;
            } else if (type == 2) {
                if (bsize < 5 + 1 + 6 + 1 || size - at < bsize) return false;
                // Literals_Section_Header: compressed literals (2), size format 3 (4 streams, 18 + 18 bits)
                const uint8_t *h = p + at;
                if ((h[0] & 3u) != 2u || ((h[0] >> 2) & 3u) != 3u) return false;
                const uint64_t v = static_cast<uint64_t>(zhuf_get_le(h, 4)) | (static_cast<uint64_t>(h[4]) << 32);
                const uint32_t regen = static_cast<uint32_t>((v >> 4) & 0x3ffffu), lit = static_cast<uint32_t>((v >> 22) & 0x3ffffu);
                if (regen == 0 || regen > kZhufBlock || 5 + static_cast<uint64_t>(lit) + 1 != bsize) return false;
                if (h[5 + lit] != 0) return false;   // Sequences_Section: anything but "0 sequences" needs real zstd
                b.src = at + 5;
                b.dst = raw + got;
                b.regen = regen;
                b.coded = 1;
                b.lit = lit;
            } else {
                return false;   // RLE / reserved
            }
            got += b.regen;
            at += bsize;
            nb++;
            if (last) break;
        }
        if (got != fcs) return false;
        raw += got;
    }
    *nblocks = nb;
    *raw_total = raw;
    return true;
}

// backward bit reader over bytes [p, p + len): the highest set bit of the last byte closes the stream
struct ZhufBackBits {
    const uint8_t *p;
    long long pos;    // bits left below the read position
    bool over;        // a read reached below bit 0 (zeros are supplied, as libzstd does)
};
SZ_HD bool zhuf_back_init(ZhufBackBits &r, const uint8_t *p, uint32_t len) {
    r.p = p;
    r.over = false;
    r.pos = 0;
    if (len == 0 || p[len - 1] == 0) return false;
    int hb = 7;
    while (!((p[len - 1] >> hb) & 1)) hb--;
    r.pos = static_cast<long long>(len - 1) * 8 + hb;
    return true;
}
SZ_HD uint32_t zhuf_back_read(ZhufBackBits &r, int n) {   // n <= 16; most significant bit first
    uint32_t v = 0;
    for (int i = 0; i < n; i++) {
        r.pos--;
        uint32_t bit = 0;
        if (r.pos >= 0)
            bit = (r.p[r.pos >> 3] >> (r.pos & 7)) & 1u;
        else
            r.over = true;
        v = (v << 1) | bit;
    }
    return v;
}

struct ZhufDecScratch {   // shared memory of the decoding CTA
    uint8_t w[256];       // weights of the symbols (0 = absent)
    int nsym, maxbits, ok;
    uint32_t desc_len;
    uint32_t start[kZhufMaxBits + 2];   // first table cell of each weight
    // FSE decoding table of the weights (accuracy log <= 6)
    int norm[16];
    uint8_t fsym[64], fnb[64];
    uint16_t fbase[64];
};

// Huffman_Tree_Description at d[0, avail) -> S.w[0, S.nsym), S.maxbits, S.desc_len.  One thread (a few hundred serial
// steps).  Returns false for descriptions this decoder does not take (the stream is then reported as corrupt).
SZ_HD bool zhuf_read_weights(const uint8_t *d, uint32_t avail, ZhufDecScratch &S) {
    if (avail < 1) return false;
    const uint32_t hb = d[0];
    int nw = 0;
    for (int i = 0; i < 256; i++) S.w[i] = 0;
    if (hb >= 128) {   // direct: 4 bits per weight
        nw = static_cast<int>(hb) - 127;
        const uint32_t nbytes = static_cast<uint32_t>(nw + 1) / 2;
        if (1 + nbytes > avail) return false;
        for (int i = 0; i < nw; i++) S.w[i] = (i & 1) ? (d[1 + i / 2] & 15u) : (d[1 + i / 2] >> 4);
        S.desc_len = 1 + nbytes;
    } else {   // FSE-compressed
        if (hb == 0 || 1 + hb > avail) return false;
        const uint8_t *f = d + 1;
        // ---- FSE_Table_Description: forward bit stream, least significant bit first
        uint32_t bitpos = 0;
        const uint32_t nbits_total = hb * 8;
        auto rd = [&](int n) -> uint32_t {
            uint32_t v = 0;
            for (int i = 0; i < n; i++, bitpos++)
                if (bitpos < nbits_total) v |= static_cast<uint32_t>((f[bitpos >> 3] >> (bitpos & 7)) & 1u) << i;
            return v;
        };
        const int log = static_cast<int>(rd(4)) + 5;
        if (log > 6) return false;
        const int size = 1 << log;
        int remaining = size + 1, threshold = size, nbits = log + 1, s = 0;
        for (int i = 0; i < 16; i++) S.norm[i] = 0;
        while (remaining > 1 && s < 16) {
            const int max = (2 * threshold - 1) - remaining;
            // small values take nbits - 1 bits: peek nbits - 1, then decide
            const uint32_t save = bitpos;
            int v = static_cast<int>(rd(nbits - 1));
            if (v >= max) {
                bitpos = save;
                v = static_cast<int>(rd(nbits));
                if (v >= threshold) v -= max;
            }
            const int c = v - 1;   // -1 would be a "less than one" probability: zhuf never writes it
            if (c < 0) return false;
            S.norm[s++] = c;
            remaining -= c;
            if (c == 0) {
                for (;;) {
                    const uint32_t rep = rd(2);
                    s += static_cast<int>(rep);
                    if (rep != 3) break;
                }
                if (s > 16) return false;
            }
            while (remaining < threshold && threshold > 1) {
                nbits--;
                threshold >>= 1;
            }
        }
        if (remaining != 1 || bitpos > nbits_total) return false;
        const int max_sym = s - 1;
        const uint32_t hdr = (bitpos + 7) / 8;
        // ---- decoding table: symbols spread with step size/2 + size/8 + 3; cell u of symbol x: the k-th cell of x in
        //      table order has nextState counter norm[x] + k
        {
            const int step = (size >> 1) + (size >> 3) + 3;
            int pos = 0;
            for (int x = 0; x <= max_sym; x++)
                for (int i = 0; i < S.norm[x]; i++) {
                    S.fsym[pos] = static_cast<uint8_t>(x);
                    pos = (pos + step) & (size - 1);
                }
            if (pos != 0) return false;
            int next[16];
            for (int x = 0; x < 16; x++) next[x] = S.norm[x];
            for (int u = 0; u < size; u++) {
                const int x = S.fsym[u];
                const int ns = next[x]++;
                int hbit = 0;
                for (int t = ns; t > 1; t >>= 1) hbit++;
                const int nb = log - hbit;
                S.fnb[u] = static_cast<uint8_t>(nb);
                S.fbase[u] = static_cast<uint16_t>((ns << nb) - size);
            }
        }
        // ---- the weights: two interleaved states over the backward bitstream
        ZhufBackBits br;
        if (hb <= hdr || !zhuf_back_init(br, f + hdr, hb - hdr)) return false;
        uint32_t s1 = zhuf_back_read(br, log), s2 = zhuf_back_read(br, log);
        if (br.over) return false;
        for (;;) {
            if (nw > 253) return false;
            S.w[nw++] = S.fsym[s1];
            s1 = S.fbase[s1] + zhuf_back_read(br, S.fnb[s1]);
            if (br.over) {
                S.w[nw++] = S.fsym[s2];
                break;
            }
            S.w[nw++] = S.fsym[s2];
            s2 = S.fbase[s2] + zhuf_back_read(br, S.fnb[s2]);
            if (br.over) {
                S.w[nw++] = S.fsym[s1];
                break;
            }
        }
        S.desc_len = 1 + hb;
    }
    // ---- the implied last weight
    uint32_t total = 0;
    for (int i = 0; i < nw; i++) {
        if (S.w[i] > kZhufMaxBits) return false;
        if (S.w[i]) total += 1u << (S.w[i] - 1);
    }
    if (total == 0 || nw > 255) return false;
    int maxbits = 0;
    while ((1u << maxbits) <= total) maxbits++;   // highbit(total) + 1
    if (maxbits > kZhufMaxBits) return false;
    const uint32_t rest = (1u << maxbits) - total;
    if (rest & (rest - 1)) return false;   // must be a power of two
    int lw = 0;
    while ((1u << lw) < rest) lw++;
    S.w[nw] = static_cast<uint8_t>(lw + 1);
    S.nsym = nw + 1;
    S.maxbits = maxbits;
    return true;
}

// Decode table of 2^maxbits cells (nbits << 8 | symbol) from S.w, by the tid-th of nt threads (all of them call this).
SZ_HD void zhuf_dec_table(ZhufDecScratch &S, uint16_t *tab, int tid, int nt) {
    if (tid == 0) {
        uint32_t cnt[kZhufMaxBits + 2];
        for (int b = 0; b < kZhufMaxBits + 2; b++) cnt[b] = 0;
        for (int i = 0; i < S.nsym; i++) cnt[S.w[i]]++;
        uint32_t at = 0;
        for (int b = 1; b <= kZhufMaxBits + 1; b++) {
            S.start[b] = at;
            at += cnt[b] << (b - 1);
        }
    }
    SZ_CTA_SYNC();
    for (int s = tid; s < S.nsym; s += nt) {
        const uint32_t w = S.w[s];
        if (!w) continue;
        uint32_t before = 0;   // symbols of the same weight below s
        for (int j = 0; j < s; j++) before += S.w[j] == w ? 1u : 0u;
        const uint32_t n = 1u << (w - 1);
        const uint32_t first = S.start[w] + before * n;
        const uint16_t e = static_cast<uint16_t>(((S.maxbits + 1 - w) << 8) | static_cast<uint32_t>(s));
        for (uint32_t k = 0; k < n; k++) tab[first + k] = e;
    }
    SZ_CTA_SYNC();
}

SZ_HD uint32_t zhuf_load32(const uint8_t *p) {   // p is 4-byte aligned
#if defined(__CUDA_ARCH__)
    return *reinterpret_cast<const uint32_t *>(p);
#else
    uint32_t v;
    memcpy(&v, p, 4);
    return v;
#endif
}

SZ_HD void zhuf_store32(uint8_t *p, uint32_t v) {   // p is 4-byte aligned
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint32_t *>(p) = v;
#else
    memcpy(p, &v, 4);
#endif
}

// One Huffman stream src[0, len) -> n symbols at dst.  A 64-bit window holds the bits below the read position, most
// significant first (bit 63 = the next bit).  A code is at most 11 bits, so after a top-up to more than 32 valid bits
// two symbols decode without a look at the fill level: four symbols per round -- two top-up tests, four table
// lookups, one 4-byte store -- with the aligned 4-byte load of a top-up issued one top-up ahead of its use.  The one
// thread that runs this has nothing to hide latencies behind: the round is written for a short dependent chain
// (lookup -> length -> shift).  Returns false when the stream does not end exactly on its first bit (reads below it
// see zeros, as in libzstd).
SZ_HD bool zhuf_dec_stream(const uint8_t *src, uint32_t len, const uint16_t *tab, int maxbits, uint8_t *dst, uint32_t n) {
    if (len == 0 || src[len - 1] == 0) return false;
    int hb = 7;
    while (!((src[len - 1] >> hb) & 1)) hb--;
    const uint32_t total = (len - 1) * 8 + static_cast<uint32_t>(hb);   // bits below the closing bit
    unsigned long long win = hb ? static_cast<unsigned long long>(src[len - 1] & ((1u << hb) - 1u)) << (64 - hb) : 0ull;
    int have = hb;             // valid bits in the window
    uint32_t ptr = len - 1;    // bytes below src + ptr are not in the window yet
    while (ptr > 0 && (reinterpret_cast<uintptr_t>(src + ptr) & 3u)) {   // down to an aligned address: at most 3 bytes
        ptr--;
        win |= static_cast<unsigned long long>(src[ptr]) << (56 - have);
        have += 8;
    }
    uint32_t pre = ptr >= 4 ? zhuf_load32(src + ptr - 4) : 0u;
    const int shift = 64 - maxbits;
    uint32_t used = 0;
#define ZHUF_TOPUP()                                                                  \
    if (have <= 32) {                                                                 \
        if (have < 0) return false;                                                   \
        if (ptr >= 4) {                                                               \
            win |= static_cast<unsigned long long>(pre) << (32 - have);               \
            have += 32;                                                               \
            ptr -= 4;                                                                 \
            if (ptr >= 4) pre = zhuf_load32(src + ptr - 4);                           \
        } else {                                                                      \
            while (have <= 56 && ptr > 0) {                                           \
                ptr--;                                                                \
                win |= static_cast<unsigned long long>(src[ptr]) << (56 - have);      \
                have += 8;                                                            \
            }                                                                         \
        }                                                                             \
    }
#define ZHUF_SYM(out)                                                 \
    {                                                                 \
        const uint32_t e_ = tab[win >> shift];                        \
        const uint32_t nb_ = e_ >> 8;                                 \
        out = e_ & 0xffu;                                             \
        win <<= nb_;                                                  \
        have -= static_cast<int>(nb_);                                \
        used += nb_;                                                  \
    }
    uint32_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 3u)) {
        uint32_t b;
        ZHUF_TOPUP();
        ZHUF_SYM(b);
        dst[i++] = static_cast<uint8_t>(b);
    }
    while (i + 4 <= n) {
        uint32_t b0, b1, b2, b3;
        ZHUF_TOPUP();
        ZHUF_SYM(b0);
        ZHUF_SYM(b1);
        ZHUF_TOPUP();
        ZHUF_SYM(b2);
        ZHUF_SYM(b3);
        zhuf_store32(dst + i, b0 | (b1 << 8) | (b2 << 16) | (b3 << 24));
        i += 4;
    }
    while (i < n) {
        uint32_t b;
        ZHUF_TOPUP();
        ZHUF_SYM(b);
        dst[i++] = static_cast<uint8_t>(b);
    }
#undef ZHUF_TOPUP
#undef ZHUF_SYM
    return used == total;
}

// Sizes of the four streams of a compressed block whose tree description took desc_len of its `lit` bytes:
// d = start of the tree description.  Returns false on inconsistent sizes.
SZ_HD bool zhuf_stream_sizes(const uint8_t *d, uint32_t lit, uint32_t desc_len, uint32_t regen, uint32_t sb[4], uint32_t sn[4]) {
    if (desc_len + 6 > lit) return false;
    const uint8_t *j = d + desc_len;
    sb[0] = zhuf_get_le(j, 2);
    sb[1] = zhuf_get_le(j + 2, 2);
    sb[2] = zhuf_get_le(j + 4, 2);
    const uint32_t rest = lit - desc_len - 6;
    if (sb[0] + sb[1] + sb[2] >= rest) return false;
    sb[3] = rest - sb[0] - sb[1] - sb[2];
    const uint32_t seg = (regen + 3) / 4;
    if (3 * seg >= regen) return false;
    sn[0] = sn[1] = sn[2] = seg;
    sn[3] = regen - 3 * seg;
    return true;
}

}  // namespace sz3b
