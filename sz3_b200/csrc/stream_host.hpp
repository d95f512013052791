// sz3_b200/csrc/stream_host.hpp -- byte-level stream format helpers of the host tail:
// Config blob, little-endian writers, and the Lossless_zstd framing with a thread-parallel multi-frame zstd pass.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/sz3b.h"
#include "zstd_decl.h"

namespace sz3b {

constexpr uint32_t kMagic = 0xF342F310u;                               // SZ3_MAGIC_NUMBER (version.hpp.in)
constexpr uint32_t kDataVer = (3u << 24) | (3u << 16) | (2u << 8) | 0; // versionInt("3.3.2")

template <class V>
inline void put(uint8_t *&p, const V &v) {
    memcpy(p, &v, sizeof(V));
    p += sizeof(V);
}
template <class V>
inline V get(const uint8_t *&p) {
    V v;
    memcpy(&v, p, sizeof(V));
    p += sizeof(V);
    return v;
}

// Config::setDims (Config.hpp:161-177)
void config_set_dims(sz3b_config &c, int nd, const uint64_t *dims);
inline uint64_t config_num(const sz3b_config &c) {
    uint64_t n = 1;
    for (int i = 0; i < c.N; i++) n *= c.dims[i];
    return n;
}
// Config::save / load (Config.hpp:312-413)
size_t config_save(const sz3b_config &c, uint8_t *out);
bool config_load(sz3b_config &c, const uint8_t *in, size_t len);

// Lossless_zstd::compress (lossless/Lossless_zstd.hpp:29-37): size_t srcLen | zstd frame(s), level 3.
// Returns the bytes written, or 0 with *too_small = true when dstCap - 8 < ZSTD_compressBound(srcLen) (the reference
// throws std::length_error there).  With threads > 1 the source is cut into chunks compressed concurrently and the
// frames are concatenated; ZSTD_decompress (what the reference decoder calls) accepts concatenated frames.
// `ready`, if given, is called by a worker before it touches src[0, upto): it blocks until those bytes have arrived
// (the packed stream is still streaming in from the GPU while the first chunks are already being compressed).
// `scratch` holds the per-chunk frames before they are concatenated; it only grows, so callers keep it across calls.
struct ZstdReady {
    virtual void wait(size_t upto) = 0;
    virtual ~ZstdReady() {}
};
// `allow_raw`: the stream is entropy-coded already (post-Huffman), so the adaptive policy may store chunks as raw
// zstd frames when zstd gains < 1 % on sampled chunks (lossless_policy() == 1, the default; 0 = always compress).
size_t zstd_compress_framed(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, int threads,
                            bool *too_small, ZstdReady *ready = nullptr, std::vector<uint8_t> *scratch = nullptr,
                            bool allow_raw = false);
void set_lossless_policy(int p);
int lossless_policy();
void set_frame_decoder(int m);   // 1 = zhuf-shaped frames decode on the GPU (zhuf_dec.cuh), 0 = libzstd on the host
int frame_decoder();

// Persistent host worker threads of the host tail (zstd chunks, frame concatenation).
void host_parallel(int nworkers, void (*fn)(void *arg, int worker), void *arg);
// Lossless_zstd::decompress (:39-45).  Returns false on a zstd error.
bool zstd_decompress_framed(const uint8_t *src, size_t src_len, std::vector<uint8_t> &out);
bool zstd_decompress_into(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, size_t *dst_len);
// Same stream, frames decompressed concurrently when every frame states its content size (the frames this library
// writes do); falls back to one ZSTD_decompress call otherwise.  *raw_len = the size_t prefix.
// `done` (optional) hears about every stretch of `dst` as soon as it is decoded (from the worker threads)
struct FrameDone {
    virtual void frame(size_t off, size_t len) = 0;
    virtual ~FrameDone() {}
};
bool zstd_decompress_parallel(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, size_t *raw_len, int threads,
                              FrameDone *done = nullptr);
size_t zstd_framed_raw_len(const uint8_t *src, size_t src_len);

int host_threads();
void set_host_threads(int n);

}  // namespace sz3b
