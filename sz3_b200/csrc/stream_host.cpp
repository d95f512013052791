// sz3_b200/csrc/stream_host.cpp -- see stream_host.hpp.
#include "stream_host.hpp"

#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace sz3b {

static std::atomic<int> g_host_threads{0};

int host_threads() {
    int n = g_host_threads.load();
    if (n <= 0) {
        // this process's share of the cores when a launcher says how many ranks share the host (torchrun, Open MPI):
        // a full-size pool per rank oversubscribes the cores and halves the end-to-end rate (tests/gpu_cores.sh).  At
        // most six per rank: every concurrent tuner trial has a thread spinning on its stream, and twelve of them per
        // rank on a 2 x 12-core host stalled the ranks' main threads and the NCCL proxies for up to 145 ms
        // (tests/gpu_2gpu_probe.sh: 6.1 / 14.4 ms per step against 5.1 / 13.0 ms with six)
        static const int share = [] {
            int c = static_cast<int>(std::thread::hardware_concurrency());
            if (c <= 0) c = 1;
            if (c > 64) c = 64;
            int ranks = 1;
            for (const char *name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE"})
                if (const char *e = getenv(name)) ranks = std::max(ranks, atoi(e));
            return ranks > 1 ? std::max(2, std::min(c / ranks, 6)) : c;
        }();
        n = share;
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

// The blob comes from an untrusted stream (sz3b_peek_config, sz3b_decompress, the slab Configs of an OpenMP container):
// every read is checked against min(confSize, len); a bit width beyond 64 or a zero dimension is rejected.
bool config_load(sz3b_config &c, const uint8_t *in, size_t len) {
    if (len < 4) return false;
    const uint8_t conf_size = in[0];   // counts itself (Config::save stores the distance to the start of the blob)
    if (conf_size < 4 || conf_size > len) return false;
    const uint8_t *p = in + 1;
    const uint8_t *const end = in + conf_size;
    auto room = [&](size_t n) { return static_cast<size_t>(end - p) >= n; };
    if (!room(2)) return false;
    c.N = get<char>(p);
    if (c.N < 1 || c.N > 4) return false;
    const uint8_t bw = get<uint8_t>(p);
    if (bw < 1 || bw > 64) return false;
    const size_t nbits = static_cast<size_t>(bw) * c.N;
    if (!room((nbits + 7) / 8)) return false;
    for (int i = 0; i < 4; i++) c.dims[i] = 0;
    size_t bit = 0;
    for (int i = 0; i < c.N; i++)
        for (int j = 0; j < bw; j++, bit++)
            c.dims[i] |= static_cast<uint64_t>((p[bit >> 3] >> (bit & 7)) & 1u) << j;
    for (int i = 0; i < c.N; i++)
        if (c.dims[i] == 0) return false;
    p += (nbits + 7) / 8;
    if (!room(8 + 1 + 1)) return false;
    (void)get<uint64_t>(p);  // num
    c.cmprAlgo = get<uint8_t>(p);
    c.errorBoundMode = get<uint8_t>(p);
    switch (c.errorBoundMode) {
        case SZ3B_EB_ABS:
            if (!room(8)) return false;
            c.absErrorBound = get<double>(p);
            break;
        case SZ3B_EB_REL:
            if (!room(8)) return false;
            c.relErrorBound = get<double>(p);
            break;
        case SZ3B_EB_PSNR:
            if (!room(8)) return false;
            c.psnrErrorBound = get<double>(p);
            break;
        case SZ3B_EB_L2NORM:
            if (!room(8)) return false;
            c.l2normErrorBound = get<double>(p);
            break;
        case SZ3B_EB_ABS_OR_REL:
        case SZ3B_EB_ABS_AND_REL:
            if (!room(16)) return false;
            c.absErrorBound = get<double>(p);
            c.relErrorBound = get<double>(p);
            break;
        default: break;
    }
    if (room(1)) {
        uint8_t b = get<uint8_t>(p);
        c.lorenzo = (b >> 7) & 1;
        c.lorenzo2 = (b >> 6) & 1;
        c.regression = (b >> 5) & 1;
        c.regression2 = (b >> 4) & 1;
        c.openmp = (b >> 3) & 1;
    }
    if (room(1)) c.dataType = get<uint8_t>(p);
    if (room(4)) c.quantbinCnt = get<int32_t>(p);
    if (room(4)) c.blockSize = get<int32_t>(p);
    if (room(1)) c.predDim = get<uint8_t>(p);
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// persistent worker pool
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct HostPool {
    std::mutex run_mu;            // one parallel region at a time (concurrent callers queue up here)
    std::mutex mu;
    std::condition_variable cv_start, cv_done;
    std::vector<std::thread> threads;
    uint64_t generation = 0;
    int active = 0;               // workers (beyond the caller) taking part in the current region
    int pending = 0;
    void (*fn)(void *, int) = nullptr;
    void *arg = nullptr;
    bool stop = false;

    void worker_main(int idx) {
        uint64_t seen = 0;
        for (;;) {
            void (*f)(void *, int);
            void *a;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_start.wait(lk, [&] { return stop || (generation != seen && idx <= active); });
                if (stop) return;
                seen = generation;
                f = fn;
                a = arg;
            }
            f(a, idx);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    void run(int nworkers, void (*f)(void *, int), void *a) {
        std::lock_guard<std::mutex> rl(run_mu);
        if (nworkers < 1) nworkers = 1;
        {
            std::unique_lock<std::mutex> lk(mu);
            while (static_cast<int>(threads.size()) < nworkers - 1) {
                int idx = static_cast<int>(threads.size()) + 1;
                threads.emplace_back([this, idx] { worker_main(idx); });
            }
            fn = f;
            arg = a;
            active = nworkers - 1;
            pending = nworkers - 1;
            generation++;
        }
        cv_start.notify_all();
        f(a, 0);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
        active = 0;
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_start.notify_all();
        for (auto &t : threads) t.join();
    }
};
HostPool &pool() {
    static HostPool *p = new HostPool();   // intentionally leaked: worker threads must not be joined at exit
    return *p;
}
}  // namespace

void host_parallel(int nworkers, void (*fn)(void *arg, int worker), void *arg) { pool().run(nworkers, fn, arg); }

// ---------------------------------------------------------------------------------------------------------------------
namespace {
// A zstd frame that stores its content (Raw_Block, RFC 8878 3.1.1.2): magic | descriptor 0xA0 (single segment, 4-byte
// content size) | size | blocks of <= 128 KiB, each with a 3-byte header (last flag | type 0 << 1 | size << 3).
constexpr size_t kRawBlock = static_cast<size_t>(128) << 10;
size_t raw_frame_size(size_t len) { return 9 + 3 * (len ? (len + kRawBlock - 1) / kRawBlock : 1) + len; }
size_t raw_frame_write(const uint8_t *src, size_t len, uint8_t *dst) {
    uint8_t *p = dst;
    put<uint32_t>(p, 0xFD2FB528u);
    put<uint8_t>(p, 0xA0);
    put<uint32_t>(p, static_cast<uint32_t>(len));
    size_t off = 0;
    do {
        const size_t n = std::min(kRawBlock, len - off);
        const uint32_t hdr = static_cast<uint32_t>(n << 3) | (off + n >= len ? 1u : 0u);
        p[0] = static_cast<uint8_t>(hdr);
        p[1] = static_cast<uint8_t>(hdr >> 8);
        p[2] = static_cast<uint8_t>(hdr >> 16);
        p += 3;
        memcpy(p, src + off, n);
        p += n;
        off += n;
    } while (off < len);
    return static_cast<size_t>(p - dst);
}

struct ZJob {
    const uint8_t *src;
    size_t src_len, chunk, nchunks, slot;
    uint8_t *scratch;
    uint8_t *dst;
    std::vector<size_t> sizes, offs;
    std::vector<uint8_t> raw;     // 1 = chunk stored as a raw frame written straight into dst
    std::atomic<size_t> next{0};
    std::atomic<bool> failed{false};
    ZstdReady *ready;
    int phase = 0;                // 0 = probes, 1 = remaining chunks compressed, 2 = concatenate (+ raw frames)
    size_t probe_stride = 1;
};
thread_local ZSTD_CCtx *t_cctx = nullptr;

bool zjob_compress(ZJob &j, size_t k) {
    const size_t off = k * j.chunk;
    const size_t len = std::min(j.chunk, j.src_len - off);
    if (j.ready) j.ready->wait(off + len);
    if (!t_cctx) t_cctx = ZSTD_createCCtx();
    size_t r = t_cctx ? ZSTD_compressCCtx(t_cctx, j.scratch + k * j.slot, j.slot, j.src + off, len, 3)
                      : ZSTD_compress(j.scratch + k * j.slot, j.slot, j.src + off, len, 3);
    if (ZSTD_isError(r)) return false;
    j.sizes[k] = r;
    return true;
}

void zjob_worker(void *arg, int) {
    ZJob &j = *static_cast<ZJob *>(arg);
    for (;;) {
        size_t k = j.next.fetch_add(1);
        if (k >= j.nchunks) break;
        const bool probe = k % j.probe_stride == 0;
        if (j.phase == 0) {
            if (probe && !zjob_compress(j, k)) j.failed = true;
        } else if (j.phase == 1) {
            if (!probe && !zjob_compress(j, k)) j.failed = true;
        } else {
            const size_t off = k * j.chunk;
            const size_t len = std::min(j.chunk, j.src_len - off);
            if (j.raw[k]) {
                if (j.ready) j.ready->wait(off + len);
                raw_frame_write(j.src + off, len, j.dst + j.offs[k]);
            } else {
                memcpy(j.dst + j.offs[k], j.scratch + k * j.slot, j.sizes[k]);
            }
        }
    }
}

std::atomic<int> g_lossless_policy{2};
}  // namespace

void set_lossless_policy(int p) { g_lossless_policy.store(p); }
int lossless_policy() { return g_lossless_policy.load(); }

namespace {
constexpr int kFrameDecoderDefault = 1;   // 512^3 stream: 6.28 -> 5.55 ms (profiles/r2h_frames.log)
std::atomic<int> g_frame_decoder{[] {
    const char *e = getenv("SZ3B_FRAME_DECODER");
    return e && *e ? (atoi(e) != 0 ? 1 : 0) : kFrameDecoderDefault;
}()};
}  // namespace
void set_frame_decoder(int m) { g_frame_decoder.store(m != 0 ? 1 : 0); }
int frame_decoder() { return g_frame_decoder.load(); }

size_t zstd_compress_framed(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, int threads,
                            bool *too_small, ZstdReady *ready, std::vector<uint8_t> *scratch, bool allow_raw) {
    *too_small = false;
    if (dst_cap < sizeof(uint64_t) || dst_cap - sizeof(uint64_t) < ZSTD_compressBound(src_len)) {
        *too_small = true;
        return 0;
    }
    uint8_t *p = dst;
    put<uint64_t>(p, static_cast<uint64_t>(src_len));
    const size_t kChunk = static_cast<size_t>(1) << 20;
    if (threads <= 1 || src_len < 2 * kChunk) {
        if (ready) ready->wait(src_len);
        // (a context kept per thread: the tuner's trial streams are a few hundred KB each, where creating and
        //  clearing a fresh context costs as much as the compression; same bytes as ZSTD_compress)
        if (!t_cctx) t_cctx = ZSTD_createCCtx();
        size_t r = t_cctx ? ZSTD_compressCCtx(t_cctx, p, dst_cap - 8, src, src_len, 3) : ZSTD_compress(p, dst_cap - 8, src, src_len, 3);
        if (ZSTD_isError(r)) return 0;
        return r + 8;
    }
    ZJob j;
    j.src = src;
    j.src_len = src_len;
    // about 1 MiB per frame, but a whole number of rounds over the worker threads (46 frames on 16 threads would leave
    // the third round two-thirds empty)
    {
        const size_t rounds = (src_len + kChunk * threads - 1) / (kChunk * static_cast<size_t>(threads));
        const size_t want = rounds * static_cast<size_t>(threads);
        j.chunk = std::max<size_t>((src_len + want - 1) / want, static_cast<size_t>(256) << 10);
        j.chunk = (j.chunk + 4095) & ~static_cast<size_t>(4095);
    }
    j.nchunks = (src_len + j.chunk - 1) / j.chunk;
    j.slot = ZSTD_compressBound(j.chunk);
    std::vector<uint8_t> local;
    std::vector<uint8_t> &sc = scratch ? *scratch : local;
    if (sc.size() < j.nchunks * j.slot) sc.resize(j.nchunks * j.slot);
    j.scratch = sc.data();
    j.dst = p;
    j.sizes.assign(j.nchunks, 0);
    j.offs.assign(j.nchunks, 0);
    j.raw.assign(j.nchunks, 0);
    j.ready = ready;
    const int nt = static_cast<int>(std::min<size_t>(threads, j.nchunks));
    // Adaptive policy (post-Huffman streams only): every 8th chunk is a probe.  When zstd gains less than 1 % on the
    // probes -- entropy-coded indices of noisy data -- the other chunks are stored as raw frames: at most 0.875 % of
    // compression ratio traded for 7/8 of the host time.  The stream stays a plain concatenation of zstd frames.
    const bool adaptive = allow_raw && g_lossless_policy.load() == 1 && j.nchunks >= 16 && j.chunk <= 0xfffffff0u;
    j.probe_stride = adaptive ? 8 : 1;
    j.phase = 0;
    host_parallel(nt, zjob_worker, &j);
    if (j.failed) return 0;
    bool store_raw = false;
    if (adaptive) {
        size_t in = 0, out = 0;
        for (size_t k = 0; k < j.nchunks; k += j.probe_stride) {
            in += std::min(j.chunk, src_len - k * j.chunk);
            out += j.sizes[k];
        }
        store_raw = static_cast<double>(out) > 0.99 * static_cast<double>(in);
        if (!store_raw) {
            j.phase = 1;
            j.next = 0;
            host_parallel(nt, zjob_worker, &j);
            if (j.failed) return 0;
        }
    }
    size_t total = 0;
    for (size_t k = 0; k < j.nchunks; k++) {
        if (store_raw && k % j.probe_stride != 0) {
            j.raw[k] = 1;
            j.sizes[k] = raw_frame_size(std::min(j.chunk, src_len - k * j.chunk));
        }
        j.offs[k] = total;
        total += j.sizes[k];
    }
    if (total > dst_cap - 8) {
        *too_small = true;
        return 0;
    }
    j.phase = 2;
    j.next = 0;
    host_parallel(nt, zjob_worker, &j);
    return total + 8;
}

size_t zstd_framed_raw_len(const uint8_t *src, size_t src_len) {
    if (src_len < 8) return 0;
    uint64_t n;
    memcpy(&n, src, 8);
    return static_cast<size_t>(n);
}

namespace {
struct DFrame {
    const uint8_t *src;
    size_t csize, off, dsize;
};
struct DJob {
    std::vector<DFrame> frames;
    uint8_t *dst;
    std::atomic<size_t> next{0};
    std::atomic<bool> failed{false};
    FrameDone *done = nullptr;
};
void djob_worker(void *arg, int) {
    DJob &j = *static_cast<DJob *>(arg);
    for (;;) {
        size_t k = j.next.fetch_add(1);
        if (k >= j.frames.size()) break;
        const DFrame &f = j.frames[k];
        if (!j.done) {
            size_t r = ZSTD_decompress(j.dst + f.off, f.dsize, f.src, f.csize);
            if (ZSTD_isError(r) || r != f.dsize) j.failed = true;
            continue;
        }
        // with a listener: streaming decompression, so that every 256 KiB of the frame is reported (and travels on)
        // while the rest is still being decoded
        thread_local ZSTD_DCtx *dctx = nullptr;
        if (!dctx) dctx = ZSTD_createDCtx();
        if (!dctx) {
            j.failed = true;
            continue;
        }
        ZSTD_outBuffer out{j.dst + f.off, f.dsize, 0};
        size_t in_pos = 0, reported = 0, window = static_cast<size_t>(160) << 10;
        bool ok = true;
        for (;;) {
            // (input is offered a block or so at a time: one call with the whole frame would decode all of it before
            //  returning)
            ZSTD_inBuffer piece{f.src, std::min(f.csize, in_pos + window), in_pos};
            const size_t prev_out = out.pos;
            const size_t r = ZSTD_decompressStream(dctx, &out, &piece);
            if (ZSTD_isError(r)) {
                ok = false;
                break;
            }
            const bool progress = piece.pos > in_pos || out.pos > prev_out;
            in_pos = piece.pos;
            if (out.pos - reported >= (static_cast<size_t>(256) << 10) || r == 0 || out.pos == f.dsize) {
                if (out.pos > reported) j.done->frame(f.off + reported, out.pos - reported);
                reported = out.pos;
            }
            if (r == 0) break;   // frame complete
            if (!progress) {
                if (piece.size >= f.csize) {   // everything was on offer: truncated or oversized frame
                    ok = false;
                    break;
                }
                window *= 2;
            }
        }
        if (!ok || out.pos != f.dsize) {
            j.failed = true;
            ZSTD_freeDCtx(dctx);   // (a context left in mid-frame state is not reused)
            dctx = nullptr;
        }
    }
}
}  // namespace

bool zstd_decompress_parallel(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, size_t *raw_len, int threads,
                              FrameDone *done) {
    if (src_len < 8) return false;
    const size_t n = zstd_framed_raw_len(src, src_len);
    *raw_len = n;
    if (n > dst_cap) return false;
    const uint8_t *p = src + 8;
    size_t rem = src_len - 8;
    DJob j;
    j.dst = dst;
    j.done = done;
    size_t off = 0;
    bool splittable = true;
    while (rem > 0) {
        size_t cs = ZSTD_findFrameCompressedSize(p, rem);
        if (ZSTD_isError(cs) || cs == 0 || cs > rem) {
            splittable = false;
            break;
        }
        unsigned long long ds = ZSTD_getFrameContentSize(p, cs);
        if (ds >= 0xfffffffffffffffeull) {   // unknown / error
            splittable = false;
            break;
        }
        j.frames.push_back({p, cs, off, static_cast<size_t>(ds)});
        off += static_cast<size_t>(ds);
        p += cs;
        rem -= cs;
    }
    if (!splittable || off != n || j.frames.size() < 2 || threads < 2) {
        size_t r = ZSTD_decompress(dst, n, src + 8, src_len - 8);
        if (ZSTD_isError(r) || r != n) return false;
        if (done) done->frame(0, n);
        return true;
    }
    host_parallel(static_cast<int>(std::min<size_t>(threads, j.frames.size())), djob_worker, &j);
    return !j.failed;
}

}  // namespace sz3b
