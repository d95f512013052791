// sz3_b200/csrc/huffman_host.hpp -- host part of the Huffman stage: tree construction from a histogram, the tree
// blob of HuffmanEncoder::save and the (code, length) table the GPU bit packer consumes.
//
// Restates reference include/SZ3/encoder/HuffmanEncoder.hpp:516-561 (init), :440-470 (priority queue),
// :478-508 (build_code), :563-579 + :601-628 (pad_tree / tree bytes), :108-125 (save), :261-279 (load).
// The tree must be *identical* to the reference's (same heap tie-breaks, leaves inserted in ascending symbol
// order), otherwise the bitstream differs; the data-proportional work (histogram, packing) runs on the GPU.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace sz3b {

struct HuffmanBook {
    int offset = 0;                 // smallest symbol (HuffmanEncoder::offset)
    uint32_t state_num = 0;         // max - offset + 2
    uint32_t node_count = 0;        // 2 * distinct - 1
    std::vector<uint64_t> code;     // per state (symbol - offset): code bits, right-aligned
    std::vector<uint8_t> len;       // per state: code length in bits (0 for absent symbols and for the 1-symbol tree)
    std::vector<uint8_t> tree_blob; // exactly what HuffmanEncoder::save writes
    uint64_t total_bits = 0;        // sum over symbols of freq * len
    int max_len = 0;
};

// hist[k] = frequency of symbol (sym_base + k), k in [0, nbins).  Returns false (with *err set) when the histogram
// is empty or a code is longer than 64 bits (the GPU packer's limit; unreachable below ~1e13 symbols).
bool huffman_build(const unsigned long long *hist, size_t nbins, int sym_base, HuffmanBook &book, const char **err);

// Decoder side (used by decompression): parses the blob written by save and decodes n symbols from a bitstream
// laid out as `size_t outSize | bits`.  Advances *pos past everything it consumed.  Returns false on malformed input.
bool huffman_decode(const uint8_t *&pos, size_t &remaining, size_t n, std::vector<int> &out, const char **err);

// Table-driven decoder for the main index stream: `load` parses the tree blob of HuffmanEncoder::save (:108-125,
// :261-279) and builds a kLutBits-wide first-level table (most codes of a skewed index distribution resolve in one
// lookup), `decode` reads `size_t outSize | bits` (HuffmanEncoder::decode :225-255) and writes n symbols.
struct HuffmanDecoder {
    static constexpr int kLutBits = 12;
    int offset = 0;
    uint32_t nc = 0;
    std::vector<uint32_t> L, R;
    std::vector<int> C;
    std::vector<uint8_t> leaf;
    std::vector<uint32_t> lut;   // (len << 24) | node: len > 0 -> leaf `node` reached after len bits; len == 0 -> continue at `node`
    // Device form of the table, two levels (huffman_decode.cu).  dlut[prefix]: (len << 24) | node as above for codes of
    // at most kLutBits bits; a longer code has bit 31 set, S = bits 24..30 and base = (bits 0..23) << 4: its next S bits
    // index lut2[base ..] -- (len2 << 24) | node for a leaf len2 more bits down, or (0 << 24) | node to keep walking
    // after all S bits.  Prefixes whose sub-table would not fit keep the plain (0 << 24) | node entry.
    std::vector<uint32_t> dlut, lut2;
    bool load(const uint8_t *&pos, size_t &remaining, const char **err);
    template <class Out>
    bool decode(const uint8_t *&pos, size_t &remaining, size_t n, Out *out, const char **err) const;
};

}  // namespace sz3b
