// tests/emul/zhuf_emul.cpp -- TEST INFRASTRUCTURE: sequential encoder of the "zhuf" frames built from the same table
// builder and layout code (sz3_b200/csrc/zhuf.cuh) the CUDA kernels use, so that the format can be checked against
// libzstd's decoder on a machine without a GPU.  Never linked into the product library.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../sz3_b200/csrc/zhuf.cuh"

using namespace sz3b;

extern "C" long long zhuf_emul_compress(const uint8_t *src, size_t len, uint8_t *out, size_t cap, int *coded_blocks) {
    const uint64_t nblocks = zhuf_num_blocks(len);
    std::vector<ZhufBlockInfo> info(nblocks);
    // k_zhuf_build: histogram, rank sort, table, stream sizes
    for (uint64_t g = 0; g < nblocks; g++) {
        const uint32_t bl = zhuf_block_len(len, g);
        uint32_t hist[256] = {0};
        for (uint32_t i = 0; i < bl; i++) hist[src[g * kZhufBlock + i]]++;
        uint8_t ss[256];
        uint32_t sf[256];
        ZhufScratch scratch;
        int n = 0;
        for (int s = 0; s < 256; s++) {   // rank sort, as the warp does it
            if (!hist[s]) continue;
            int rank = 0;
            for (int j = 0; j < 256; j++)
                if (hist[j] && (hist[j] < hist[s] || (hist[j] == hist[s] && j < s))) rank++;
            ss[rank] = static_cast<uint8_t>(s);
            sf[rank] = hist[s];
            n++;
        }
        zhuf_build_table(ss, sf, n, scratch, info[g], 0, 1);
        for (int s = 0; s < 4; s++) {
            uint64_t a, b, bits = 0;
            zhuf_stream_range(len, g, s, &a, &b);
            for (uint64_t i = a; i < b; i++) bits += info[g].sym[src[i]] >> 16;
            info[g].sb[s] = static_cast<uint32_t>(bits / 8 + 1);
        }
    }
    // k_zhuf_scan
    uint64_t at = 0;
    *coded_blocks = 0;
    for (uint64_t g = 0; g < nblocks; g++) {
        if (g % kZhufBlocksPerFrame == 0) at += kZhufFrameHeader;
        bool coded;
        const uint32_t payload = zhuf_block_payload(zhuf_block_len(len, g), info[g], &coded);
        info[g].coded = coded;
        info[g].off = at;
        at += 3 + payload;
        *coded_blocks += coded;
    }
    if (at > cap) return -1;
    // k_zhuf_encode
    for (uint64_t g = 0; g < nblocks; g++) {
        const uint32_t bl = zhuf_block_len(len, g);
        uint64_t p = zhuf_write_headers(out, len, g, info[g]);
        if (!info[g].coded) {
            memcpy(out + p, src + g * kZhufBlock, bl);
            continue;
        }
        for (int s = 0; s < 4; s++) {
            uint64_t a, b;
            zhuf_stream_range(len, g, s, &a, &b);
            uint8_t *dst = out + p;
            memset(dst, 0, info[g].sb[s]);
            uint64_t bit = 0;
            for (uint64_t i = b; i-- > a;) {   // last symbol first
                const uint32_t e = info[g].sym[src[i]];
                const uint32_t code = e & 0xffffu, nb = e >> 16;
                for (uint32_t k = 0; k < nb; k++, bit++)
                    if ((code >> k) & 1u) dst[bit >> 3] |= static_cast<uint8_t>(1u << (bit & 7));
            }
            dst[bit >> 3] |= static_cast<uint8_t>(1u << (bit & 7));   // end mark
            if (bit / 8 + 1 != info[g].sb[s]) return -3;
            p += info[g].sb[s];
        }
    }
    return static_cast<long long>(at);
}

// Sequential twin of k_zhuf_decode (sz3_b200/csrc/zhuf_dec.cuh): frames -> bytes.  Returns the decoded size, -1 when the
// payload is not of the shape zhuf writes, -2 when a block does not decode.
#include "../../sz3_b200/csrc/zhuf_dec.cuh"

extern "C" long long zhuf_emul_decompress(const uint8_t *cmp, size_t size, uint8_t *out, size_t cap, int misalign) {
    std::vector<ZhufDecBlock> blocks(size / 3 + 16);
    size_t nb = 0;
    uint64_t raw = 0;
    if (!zhuf_walk_frames(cmp, size, blocks.data(), blocks.size(), &nb, &raw)) return -1;
    if (raw > cap) return -3;
    // (the kernel reads the payload from a device copy whose alignment differs from the host's: try a few)
    std::vector<uint8_t> shifted(size + 16);
    memcpy(shifted.data() + misalign, cmp, size);
    const uint8_t *base = shifted.data() + misalign;
    static ZhufDecScratch S;
    static uint16_t tab[1 << kZhufMaxBits];
    for (size_t g = 0; g < nb; g++) {
        const ZhufDecBlock &b = blocks[g];
        if (!b.coded) {
            memcpy(out + b.dst, base + b.src, b.regen);
            continue;
        }
        const uint8_t *d = base + b.src;
        if (!zhuf_read_weights(d, b.lit, S)) return -2;
        for (auto &t : tab) t = 0xffff;
        zhuf_dec_table(S, tab, 0, 1);
        uint32_t sb[4], sn[4];
        if (!zhuf_stream_sizes(d, b.lit, S.desc_len, b.regen, sb, sn)) return -2;
        const uint8_t *sp = d + S.desc_len + 6;
        uint8_t *dp = out + b.dst;
        for (int s = 0; s < 4; s++) {
            if (!zhuf_dec_stream(sp, sb[s], tab, S.maxbits, dp, sn[s])) return -2;
            sp += sb[s];
            dp += sn[s];
        }
    }
    return static_cast<long long>(raw);
}
