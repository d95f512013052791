// tests/emul/huffman_host_emul.cpp -- TEST INFRASTRUCTURE (never part of the product library).
// The product's host half of the Huffman stage (sz3_b200/csrc/huffman_host.cpp: tree, tree blob, code table) driven
// without a GPU: the histogram and the bit packing, which the product does in CUDA kernels, are done here by plain
// loops, so that the CPU-only suite can pin the tree builder against HuffmanEncoder::save / ::encode of the reference
// (include/SZ3/encoder/HuffmanEncoder.hpp:108-125, :140-218).
#include "../../sz3_b200/csrc/huffman_host.cpp"

extern "C" long long emul_huffman_encode(const int *q, size_t n, int nbins, unsigned char *out, size_t *tree_len) {
    std::vector<unsigned long long> hist(static_cast<size_t>(nbins), 0);
    for (size_t i = 0; i < n; i++) {
        if (q[i] < 0 || q[i] >= nbins) return -2;
        hist[static_cast<size_t>(q[i])]++;
    }
    sz3b::HuffmanBook book;
    const char *err = nullptr;
    if (!sz3b::huffman_build(hist.data(), hist.size(), 0, book, &err)) return -1;
    unsigned char *p = out;
    memcpy(p, book.tree_blob.data(), book.tree_blob.size());
    p += book.tree_blob.size();
    *tree_len = book.tree_blob.size();
    const uint64_t out_size = (book.total_bits + 7) / 8;
    memcpy(p, &out_size, 8);
    p += 8;
    memset(p, 0, out_size);
    uint64_t bit = 0;
    for (size_t i = 0; i < n; i++) {
        const size_t s = static_cast<size_t>(q[i] - book.offset);
        const uint64_t c = book.code[s];
        for (int b = book.len[s] - 1; b >= 0; b--, bit++)
            if ((c >> b) & 1) p[bit >> 3] |= static_cast<unsigned char>(0x80u >> (bit & 7));
    }
    if (bit != book.total_bits) return -3;
    return static_cast<long long>(p - out) + static_cast<long long>(out_size);
}
