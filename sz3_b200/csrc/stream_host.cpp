// sz3_b200/csrc/stream_host.cpp -- see stream_host.hpp.
#include "stream_host.hpp"

#include <algorithm>
#include <atomic>
#include <thread>

namespace sz3b {

static std::atomic<int> g_host_threads{0};

int host_threads() {
    int n = g_host_threads.load();
    if (n <= 0) {
        n = static_cast<int>(std::thread::hardware_concurrency());
        if (n <= 0) n = 1;
        if (n > 64) n = 64;
    }
    return n;
}
void set_host_threads(int n) { g_host_threads.store(n); }

void config_set_dims(sz3b_config &c, int nd, const uint64_t *dims) {
    int n = 0;
    for (int i = 0; i < nd && n < 4; i++)
        if (dims[i] > 1) c.dims[n++] = dims[i];
    if (n == 0) c.dims[n++] = 1;
    for (int i = n; i < 4; i++) c.dims[i] = 0;
    c.N = n;
    c.predDim = n;
    c.blockSize = n == 1 ? 128 : (n == 2 ? 16 : 6);
}

size_t config_save(const sz3b_config &c, uint8_t *out) {
    uint8_t *p = out + 1;
    put<char>(p, static_cast<char>(c.N));
    // vector_bit_width + vector2bytes (ByteUtil.hpp:195-238): dims packed LSB-first with a common bit width
    uint64_t mx = 0;
    for (int i = 0; i < c.N; i++) mx = std::max<uint64_t>(mx, c.dims[i]);
    uint8_t bw = 0;
    while (mx > 0) {
        mx >>= 1;
        ++bw;
    }
    put<uint8_t>(p, bw);
    const size_t nbits = static_cast<size_t>(bw) * c.N;
    const size_t nbytes = (nbits + 7) / 8;
    memset(p, 0, nbytes);
    size_t bit = 0;
    for (int i = 0; i < c.N; i++)
        for (int j = 0; j < bw; j++, bit++)
            if ((c.dims[i] >> j) & 1) p[bit >> 3] |= static_cast<uint8_t>(1u << (bit & 7));
    p += nbytes;
    put<uint64_t>(p, config_num(c));
    put<uint8_t>(p, static_cast<uint8_t>(c.cmprAlgo));
    put<uint8_t>(p, static_cast<uint8_t>(c.errorBoundMode));
    switch (c.errorBoundMode) {
        case SZ3B_EB_ABS: put<double>(p, c.absErrorBound); break;
        case SZ3B_EB_REL: put<double>(p, c.relErrorBound); break;
        case SZ3B_EB_PSNR: put<double>(p, c.psnrErrorBound); break;
        case SZ3B_EB_L2NORM: put<double>(p, c.l2normErrorBound); break;
        case SZ3B_EB_ABS_OR_REL:
        case SZ3B_EB_ABS_AND_REL:
            put<double>(p, c.absErrorBound);
            put<double>(p, c.relErrorBound);
            break;
        default: break;
    }
    uint8_t bools = static_cast<uint8_t>((c.lorenzo & 1) << 7 | (c.lorenzo2 & 1) << 6 | (c.regression & 1) << 5 |
                                         (c.regression2 & 1) << 4 | ((c.openmp != 0) & 1) << 3);
    put<uint8_t>(p, bools);
    put<uint8_t>(p, static_cast<uint8_t>(c.dataType));
    put<int32_t>(p, c.quantbinCnt);
    put<int32_t>(p, c.blockSize);
    put<uint8_t>(p, static_cast<uint8_t>(c.predDim));
    out[0] = static_cast<uint8_t>(p - out);
    return static_cast<size_t>(p - out);
}

bool config_load(sz3b_config &c, const uint8_t *in, size_t len) {
    if (len < 4) return false;
    const uint8_t *p = in;
    const uint8_t conf_size = get<uint8_t>(p);
    const uint8_t *end = p + conf_size;   // the reference computes c1 after reading the size byte
    if (conf_size + 1u > len + 1u && conf_size > len) return false;
    c.N = get<char>(p);
    if (c.N < 1 || c.N > 4) return false;
    const uint8_t bw = get<uint8_t>(p);
    const size_t nbits = static_cast<size_t>(bw) * c.N;
    for (int i = 0; i < 4; i++) c.dims[i] = 0;
    size_t bit = 0;
    for (int i = 0; i < c.N; i++)
        for (int j = 0; j < bw; j++, bit++)
            c.dims[i] |= static_cast<uint64_t>((p[bit >> 3] >> (bit & 7)) & 1u) << j;
    p += (nbits + 7) / 8;
    (void)get<uint64_t>(p);  // num
    c.cmprAlgo = get<uint8_t>(p);
    c.errorBoundMode = get<uint8_t>(p);
    switch (c.errorBoundMode) {
        case SZ3B_EB_ABS: c.absErrorBound = get<double>(p); break;
        case SZ3B_EB_REL: c.relErrorBound = get<double>(p); break;
        case SZ3B_EB_PSNR: c.psnrErrorBound = get<double>(p); break;
        case SZ3B_EB_L2NORM: c.l2normErrorBound = get<double>(p); break;
        case SZ3B_EB_ABS_OR_REL:
        case SZ3B_EB_ABS_AND_REL:
            c.absErrorBound = get<double>(p);
            c.relErrorBound = get<double>(p);
            break;
        default: break;
    }
    if (p < end) {
        uint8_t b = get<uint8_t>(p);
        c.lorenzo = (b >> 7) & 1;
        c.lorenzo2 = (b >> 6) & 1;
        c.regression = (b >> 5) & 1;
        c.regression2 = (b >> 4) & 1;
        c.openmp = (b >> 3) & 1;
    }
    if (p < end) c.dataType = get<uint8_t>(p);
    if (p < end) c.quantbinCnt = get<int32_t>(p);
    if (p < end) c.blockSize = get<int32_t>(p);
    if (p < end) c.predDim = get<uint8_t>(p);
    return true;
}

size_t zstd_compress_framed(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, int threads,
                            bool *too_small) {
    *too_small = false;
    if (dst_cap < sizeof(uint64_t) || dst_cap - sizeof(uint64_t) < ZSTD_compressBound(src_len)) {
        *too_small = true;
        return 0;
    }
    uint8_t *p = dst;
    put<uint64_t>(p, static_cast<uint64_t>(src_len));
    const size_t kMinChunk = static_cast<size_t>(4) << 20;
    size_t nchunks = 1;
    if (threads > 1 && src_len >= 2 * kMinChunk) {
        nchunks = std::min<size_t>(static_cast<size_t>(threads) * 2, src_len / kMinChunk);
        if (nchunks < 1) nchunks = 1;
    }
    if (nchunks == 1) {
        size_t r = ZSTD_compress(p, dst_cap - 8, src, src_len, 3);
        if (ZSTD_isError(r)) return 0;
        return r + 8;
    }
    const size_t chunk = (src_len + nchunks - 1) / nchunks;
    std::vector<std::vector<uint8_t>> bufs(nchunks);
    std::vector<size_t> sizes(nchunks, 0);
    std::atomic<size_t> next{0};
    std::atomic<bool> failed{false};
    auto worker = [&]() {
        for (;;) {
            size_t k = next.fetch_add(1);
            if (k >= nchunks) break;
            size_t off = k * chunk;
            size_t len = std::min(chunk, src_len - off);
            bufs[k].resize(ZSTD_compressBound(len));
            size_t r = ZSTD_compress(bufs[k].data(), bufs[k].size(), src + off, len, 3);
            if (ZSTD_isError(r)) {
                failed = true;
                break;
            }
            sizes[k] = r;
        }
    };
    int nt = static_cast<int>(std::min<size_t>(threads, nchunks));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    if (failed) return 0;
    size_t total = 0;
    for (size_t k = 0; k < nchunks; k++) total += sizes[k];
    if (total > dst_cap - 8) {
        *too_small = true;
        return 0;
    }
    for (size_t k = 0; k < nchunks; k++) {
        memcpy(p, bufs[k].data(), sizes[k]);
        p += sizes[k];
    }
    return static_cast<size_t>(p - dst);
}

bool zstd_decompress_framed(const uint8_t *src, size_t src_len, std::vector<uint8_t> &out) {
    if (src_len < 8) return false;
    const uint8_t *p = src;
    uint64_t n = get<uint64_t>(p);
    out.resize(n);
    size_t r = ZSTD_decompress(out.data(), n, p, src_len - 8);
    return !ZSTD_isError(r) && r == n;
}

bool zstd_decompress_into(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, size_t *dst_len) {
    if (src_len < 8) return false;
    const uint8_t *p = src;
    uint64_t n = get<uint64_t>(p);
    *dst_len = n;
    if (n > dst_cap) return false;
    size_t r = ZSTD_decompress(dst, n, p, src_len - 8);
    return !ZSTD_isError(r) && r == n;
}

}  // namespace sz3b
