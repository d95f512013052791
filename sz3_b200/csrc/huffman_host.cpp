// sz3_b200/csrc/huffman_host.cpp -- see huffman_host.hpp.
#include "huffman_host.hpp"

#include <string.h>

#include <utility>

namespace sz3b {
namespace {

struct Node {
    int left = -1, right = -1;
    uint64_t freq = 0;
    int sym = 0;      // state for leaves, 0 for internal nodes (the reference's pool is zero-initialised)
    uint8_t leaf = 0;
};

// 1-indexed binary min-heap of node ids keyed by freq, with the reference's exact sift rules
// (HuffmanEncoder.hpp:440-470): ties keep insertion-dependent order, which fixes the tree shape.
struct Heap {
    struct Item {
        uint64_t f;   // the node's frequency, kept next to its id: the sift loops then touch one array only
        int id;
    };
    std::vector<Item> a;
    const std::vector<Node> *nodes;
    explicit Heap(const std::vector<Node> *n, size_t cap) : a(cap + 2), nodes(n) {}
    int end = 1;
    void push(int id) {
        const uint64_t fid = (*nodes)[id].freq;
        int i = end++;
        for (int j; (j = i >> 1) != 0; i = j) {
            if (a[j].f <= fid) break;
            a[i] = a[j];
        }
        a[i] = Item{fid, id};
    }
    int pop() {
        if (end < 2) return -1;
        const int top = a[1].id;
        end--;
        a[1] = a[end];
        int i = 1, l;
        while ((l = i << 1) < end) {
            if (l + 1 < end && a[l + 1].f < a[l].f) l++;
            if (a[i].f > a[l].f) {
                std::swap(a[i], a[l]);
                i = l;
            } else {
                break;
            }
        }
        return top;
    }
};

inline void put_be32(uint8_t *p, uint32_t v) {
    p[0] = static_cast<uint8_t>(v >> 24);
    p[1] = static_cast<uint8_t>(v >> 16);
    p[2] = static_cast<uint8_t>(v >> 8);
    p[3] = static_cast<uint8_t>(v);
}
inline uint32_t get_be32(const uint8_t *p) {
    return (static_cast<uint32_t>(p[0]) << 24) | (static_cast<uint32_t>(p[1]) << 16) |
           (static_cast<uint32_t>(p[2]) << 8) | p[3];
}

template <class W>
void write_links(uint8_t *dst, const std::vector<uint32_t> &v) {
    for (size_t i = 0; i < v.size(); i++) {
        W w = static_cast<W>(v[i]);
        memcpy(dst + i * sizeof(W), &w, sizeof(W));
    }
}

}  // namespace

bool huffman_build(const unsigned long long *hist, size_t nbins, int sym_base, HuffmanBook &book, const char **err) {
    // almost every bin of a 65536-bin histogram is empty: look at 16 bins per test
    size_t lo = nbins, hi = 0;
    std::vector<uint32_t> used;   // bins that occur, ascending
    used.reserve(1024);
    for (size_t k0 = 0; k0 < nbins; k0 += 16) {
        const size_t k1 = k0 + 16 < nbins ? k0 + 16 : nbins;
        unsigned long long any = 0;
        for (size_t k = k0; k < k1; k++) any |= hist[k];
        if (!any) continue;
        for (size_t k = k0; k < k1; k++) {
            if (hist[k]) {
                if (lo == nbins) lo = k;
                hi = k;
                used.push_back(static_cast<uint32_t>(k));
            }
        }
    }
    const size_t present = used.size();
    if (lo == nbins) {
        if (err) *err = "Huffman bins should not be empty";
        return false;
    }
    book.offset = sym_base + static_cast<int>(lo);
    book.state_num = static_cast<uint32_t>(hi - lo + 2);
    const size_t states = book.state_num;
    book.code.assign(states, 0);
    book.len.assign(states, 0);

    // sized by the symbols present, not by the symbol range: a few hundred nodes instead of megabytes of fresh pages
    std::vector<Node> nodes;
    nodes.reserve(2 * present);
    Heap heap(&nodes, 2 * present);
    size_t distinct = 0;
    for (const uint32_t k : used) {
        Node nd;
        nd.freq = hist[k];
        nd.sym = static_cast<int>(k - lo);
        nd.leaf = 1;
        nodes.push_back(nd);
        heap.push(static_cast<int>(nodes.size()) - 1);
        distinct++;
    }
    while (heap.end > 2) {
        int l = heap.pop();
        int r = heap.pop();
        Node nd;
        nd.left = l;
        nd.right = r;
        nd.freq = nodes[l].freq + nodes[r].freq;
        nodes.push_back(nd);
        heap.push(static_cast<int>(nodes.size()) - 1);
    }
    const int root = heap.a[1].id;
    book.node_count = static_cast<uint32_t>(2 * distinct - 1);

    // pre-order walk: assigns the serialisation ids of pad_tree and the codes of build_code in one pass
    const uint32_t nc = book.node_count;
    std::vector<uint32_t> L(nc, 0), R(nc, 0);
    std::vector<int> C(nc, 0);
    std::vector<uint8_t> t(nc, 0);
    uint32_t next_id = 0;
    book.total_bits = 0;
    book.max_len = 0;
    // pad_tree numbers a node when it is first visited in pre-order (the whole left subtree before the right
    // child); an explicit stack with the left child pushed last pops nodes in exactly that order.
    struct Frame {
        int node;
        int depth;
        uint64_t bits;
        int parent_id;   // serialisation id of the parent, -1 for the root
        bool is_right;
    };
    std::vector<Frame> st;
    st.push_back({root, 0, 0, -1, false});
    bool too_long = false;
    while (!st.empty()) {
        Frame fr = st.back();
        st.pop_back();
        uint32_t id = next_id++;
        if (fr.parent_id >= 0) {
            if (fr.is_right)
                R[fr.parent_id] = id;
            else
                L[fr.parent_id] = id;
        }
        const Node &nd = nodes[fr.node];
        C[id] = nd.sym;
        t[id] = nd.leaf;
        if (nd.leaf) {
            if (fr.depth > 64) {
                too_long = true;
            } else {
                book.code[nd.sym] = fr.bits;
                book.len[nd.sym] = static_cast<uint8_t>(fr.depth);
                book.total_bits += nd.freq * static_cast<uint64_t>(fr.depth);
                if (fr.depth > book.max_len) book.max_len = fr.depth;
            }
        } else {
            uint64_t b = fr.depth < 64 ? (fr.bits << 1) : 0;
            // push right first so that the left subtree is numbered first
            st.push_back({nd.right, fr.depth + 1, b | 1, static_cast<int>(id), true});
            st.push_back({nd.left, fr.depth + 1, b, static_cast<int>(id), false});
        }
    }
    if (too_long) {
        if (err) *err = "Huffman code longer than 64 bits is not supported by the GPU packer";
        return false;
    }

    // HuffmanEncoder::save (:108-125)
    const size_t lw = nc <= 256 ? 1 : (nc <= 65536 ? 2 : 4);
    const size_t blob = 4 + 4 + 4 + 1 + 2 * nc * lw + nc * sizeof(int) + nc;
    book.tree_blob.assign(blob, 0);
    uint8_t *p = book.tree_blob.data();
    memcpy(p, &book.offset, 4);
    put_be32(p + 4, nc);
    put_be32(p + 8, book.state_num / 2);
    p += 12;
    *p++ = 0;  // sysEndianType: little endian host
    if (lw == 1) {
        write_links<uint8_t>(p, L);
        write_links<uint8_t>(p + nc, R);
    } else if (lw == 2) {
        write_links<uint16_t>(p, L);
        write_links<uint16_t>(p + 2 * nc, R);
    } else {
        write_links<uint32_t>(p, L);
        write_links<uint32_t>(p + 4 * static_cast<size_t>(nc), R);
    }
    p += 2 * nc * lw;
    memcpy(p, C.data(), nc * sizeof(int));
    p += nc * sizeof(int);
    memcpy(p, t.data(), nc);
    return true;
}

bool huffman_decode(const uint8_t *&pos, size_t &remaining, size_t n, std::vector<int> &out, const char **err) {
    auto fail = [&](const char *m) {
        if (err) *err = m;
        return false;
    };
    if (remaining < 13) return fail("huffman: truncated header");
    int offset;
    memcpy(&offset, pos, 4);
    const uint32_t nc = get_be32(pos + 4);
    const uint8_t *p = pos + 13;
    const size_t lw = nc <= 256 ? 1 : (nc <= 65536 ? 2 : 4);
    const size_t body = 2 * static_cast<size_t>(nc) * lw + static_cast<size_t>(nc) * 5;
    if (nc == 0 || remaining < 13 + body + 8) return fail("huffman: truncated tree");
    std::vector<uint32_t> L(nc), R(nc);
    for (uint32_t i = 0; i < nc; i++) {
        uint32_t l = 0, r = 0;
        memcpy(&l, p + i * lw, lw);
        memcpy(&r, p + (nc + static_cast<size_t>(i)) * lw, lw);
        L[i] = l;
        R[i] = r;
    }
    const uint8_t *pc = p + 2 * static_cast<size_t>(nc) * lw;
    const uint8_t *pt = pc + static_cast<size_t>(nc) * 4;
    std::vector<int> C(nc);
    memcpy(C.data(), pc, static_cast<size_t>(nc) * 4);
    p = pt + nc;
    uint64_t enc_len;
    memcpy(&enc_len, p, 8);
    p += 8;
    size_t used = static_cast<size_t>(p - pos);
    if (remaining < used + enc_len) return fail("huffman: truncated bitstream");
    out.resize(n);
    if (pt[0]) {  // single-symbol tree (:233-237)
        for (size_t i = 0; i < n; i++) out[i] = C[0] + offset;
    } else {
        uint32_t node = 0;
        size_t cnt = 0;
        const uint64_t nbits = enc_len * 8;
        for (uint64_t b = 0; cnt < n; b++) {
            if (b >= nbits) return fail("huffman: bitstream exhausted");
            uint32_t bit = (p[b >> 3] >> (7 - (b & 7))) & 1u;
            node = bit ? R[node] : L[node];
            if (node == 0 || node >= nc) return fail("huffman: bad tree link");
            if (pt[node]) {
                out[cnt++] = C[node] + offset;
                node = 0;
            }
        }
    }
    p += enc_len;
    remaining -= static_cast<size_t>(p - pos);
    pos = p;
    return true;
}

bool HuffmanDecoder::load(const uint8_t *&pos, size_t &remaining, const char **err) {
    auto fail = [&](const char *m) {
        if (err) *err = m;
        return false;
    };
    if (remaining < 13) return fail("huffman: truncated header");
    memcpy(&offset, pos, 4);
    nc = get_be32(pos + 4);
    const uint8_t *p = pos + 13;
    const size_t lw = nc <= 256 ? 1 : (nc <= 65536 ? 2 : 4);
    const size_t body = 2 * static_cast<size_t>(nc) * lw + static_cast<size_t>(nc) * 5;
    if (nc == 0 || remaining < 13 + body) return fail("huffman: truncated tree");
    L.assign(nc, 0);
    R.assign(nc, 0);
    for (uint32_t i = 0; i < nc; i++) {
        uint32_t l = 0, r = 0;
        memcpy(&l, p + i * lw, lw);
        memcpy(&r, p + (nc + static_cast<size_t>(i)) * lw, lw);
        if (l >= nc || r >= nc) return fail("huffman: bad tree link");
        L[i] = l;
        R[i] = r;
    }
    const uint8_t *pc = p + 2 * static_cast<size_t>(nc) * lw;
    C.resize(nc);
    memcpy(C.data(), pc, static_cast<size_t>(nc) * 4);
    leaf.assign(pc + static_cast<size_t>(nc) * 4, pc + static_cast<size_t>(nc) * 5);
    p = pc + static_cast<size_t>(nc) * 5;
    remaining -= static_cast<size_t>(p - pos);
    pos = p;
    // first-level table: walk the tree along every kLutBits-bit prefix
    lut.assign(1u << kLutBits, 0);
    if (!leaf[0]) {
        for (uint32_t pre = 0; pre < (1u << kLutBits); pre++) {
            uint32_t node = 0;
            uint32_t e = 0;
            for (int b = 0; b < kLutBits; b++) {
                node = ((pre >> (kLutBits - 1 - b)) & 1u) ? R[node] : L[node];
                if (node == 0) break;   // link to the root = absent child (malformed path): leave unresolved at root
                if (leaf[node]) {
                    e = (static_cast<uint32_t>(b + 1) << 24) | node;
                    break;
                }
            }
            lut[pre] = e ? e : node;   // len == 0: continue walking from `node` after kLutBits bits
        }
    }
    // second level for the device decoder: one global lookup instead of a bit-by-bit walk for codes up to
    // kLutBits + 12 bits (0.5 % of the symbols of a typical index stream, but every one of them stalled its warp)
    dlut = lut;
    lut2.clear();
    if (!leaf[0]) {
        // height below every node (iterative post-order; links to node 0 are absent children)
        std::vector<uint8_t> height(nc, 0);
        {
            std::vector<uint32_t> order;
            order.reserve(nc);
            std::vector<uint32_t> st{0};
            std::vector<uint8_t> seen(nc, 0);
            while (!st.empty()) {
                const uint32_t v = st.back();
                st.pop_back();
                if (seen[v]) continue;
                seen[v] = 1;
                order.push_back(v);
                if (!leaf[v]) {
                    if (L[v] && !seen[L[v]]) st.push_back(L[v]);
                    if (R[v] && !seen[R[v]]) st.push_back(R[v]);
                }
            }
            for (size_t i = order.size(); i-- > 0;) {
                const uint32_t v = order[i];
                if (leaf[v]) continue;
                const int hl = L[v] ? height[L[v]] + 1 : 0, hr = R[v] ? height[R[v]] + 1 : 0;
                height[v] = static_cast<uint8_t>(std::min(255, std::max(hl, hr)));
            }
        }
        constexpr size_t kMaxSecond = static_cast<size_t>(8) << 20;   // entries
        for (uint32_t pre = 0; pre < (1u << kLutBits); pre++) {
            const uint32_t e = lut[pre];
            if ((e >> 24) != 0 || e == 0) continue;   // resolved, or a malformed path left at the root
            const uint32_t top = e & 0xffffffu;
            const int S = std::min<int>(12, std::max<int>(1, height[top]));
            const size_t base = (lut2.size() + 15) & ~static_cast<size_t>(15);
            if (base + (static_cast<size_t>(1) << S) > kMaxSecond || (base >> 4) >= (1u << 24)) continue;
            lut2.resize(base + (static_cast<size_t>(1) << S), 0);
            for (uint32_t sfx = 0; sfx < (1u << S); sfx++) {
                uint32_t node = top, e2 = 0;
                for (int b = 0; b < S; b++) {
                    node = ((sfx >> (S - 1 - b)) & 1u) ? R[node] : L[node];
                    if (node == 0) break;
                    if (leaf[node]) {
                        e2 = (static_cast<uint32_t>(b + 1) << 24) | node;
                        break;
                    }
                }
                lut2[base + sfx] = e2 ? e2 : node;
            }
            dlut[pre] = 0x80000000u | (static_cast<uint32_t>(S) << 24) | static_cast<uint32_t>(base >> 4);
        }
    }
    if (lut2.empty()) lut2.push_back(0);
    return true;
}

template <class Out>
bool HuffmanDecoder::decode(const uint8_t *&pos, size_t &remaining, size_t n, Out *out, const char **err) const {
    auto fail = [&](const char *m) {
        if (err) *err = m;
        return false;
    };
    if (remaining < 8) return fail("huffman: truncated bitstream header");
    uint64_t enc_len;
    memcpy(&enc_len, pos, 8);
    const uint8_t *p = pos + 8;
    if (remaining - 8 < enc_len) return fail("huffman: truncated bitstream");
    if (leaf[0]) {   // single-symbol tree (:233-237)
        const Out v = static_cast<Out>(C[0] + offset);
        for (size_t i = 0; i < n; i++) out[i] = v;
    } else {
        const uint64_t nbits = enc_len * 8;
        uint64_t bitpos = 0;       // bits consumed
        uint64_t acc = 0;          // next bits, left-aligned
        int have = 0;              // valid bits in acc
        size_t next_byte = 0;
        auto refill = [&]() {
            while (have <= 56 && next_byte < enc_len) {
                acc |= static_cast<uint64_t>(p[next_byte++]) << (56 - have);
                have += 8;
            }
        };
        for (size_t cnt = 0; cnt < n; cnt++) {
            refill();
            uint32_t e = lut[acc >> (64 - kLutBits)];
            uint32_t len = e >> 24, node = e & 0xffffffu;
            if (len == 0) {   // long code: keep walking bit by bit
                uint64_t a = acc << kLutBits;
                int used = kLutBits;
                if (used > have) return fail("huffman: bitstream exhausted");
                for (;;) {
                    if (used >= have) {   // need more bits than buffered: consume what we have and refill
                        acc = used >= 64 ? 0 : acc << used;
                        have -= used;
                        bitpos += used;
                        used = 0;
                        refill();
                        a = acc;
                        if (have == 0) return fail("huffman: bitstream exhausted");
                    }
                    node = (a >> 63) ? R[node] : L[node];
                    a <<= 1;
                    used++;
                    if (node == 0) return fail("huffman: bad tree link");
                    if (leaf[node]) break;
                }
                len = static_cast<uint32_t>(used);
            }
            if (static_cast<int>(len) > have || bitpos + len > nbits) return fail("huffman: bitstream exhausted");
            out[cnt] = static_cast<Out>(C[node] + offset);
            acc <<= len;
            have -= static_cast<int>(len);
            bitpos += len;
        }
    }
    p += enc_len;
    remaining -= static_cast<size_t>(p - pos);
    pos = p;
    return true;
}

template bool HuffmanDecoder::decode<uint16_t>(const uint8_t *&, size_t &, size_t, uint16_t *, const char **) const;
template bool HuffmanDecoder::decode<uint32_t>(const uint8_t *&, size_t &, size_t, uint32_t *, const char **) const;
template bool HuffmanDecoder::decode<int32_t>(const uint8_t *&, size_t &, size_t, int32_t *, const char **) const;

}  // namespace sz3b
