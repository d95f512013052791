// sz3_b200/csrc/pipeline.cu -- host orchestration of the GPU compression pipelines.
//
// Mirrors, stage by stage, the reference call stack:
//   SZ_compress_dispatcher      include/SZ3/api/impl/SZDispatcher.hpp:13-76
//   SZ_compress_Interp          include/SZ3/api/impl/SZAlgoInterp.hpp:17-30
//   SZ_compress_Interp_lorenzo  include/SZ3/api/impl/SZAlgoInterp.hpp:122-286   (auto-tuner)
//   SZGenericCompressor         include/SZ3/compressor/SZGenericCompressor.hpp:38-63
//   SZ_compress_OMP             include/SZ3/api/impl/SZImplOMP.hpp:16-117
// but every data-proportional loop is a CUDA kernel; the host only builds the Huffman tree from the histogram,
// lays out the byte stream and runs zstd.  There is no CPU implementation of the kernels to fall back to.
#include <cuda_runtime.h>

#include <functional>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <exception>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>

#include "blockwise.cuh"
#include "lorenzo.cuh"
#include "zhuf.cuh"
#include "zhuf_dec.cuh"
#include "huffman_host.hpp"
#include "interp_body.cuh"
#include "interp_box.cuh"
#include "interp_plan.hpp"
#include "launch.hpp"
#include "pipeline.hpp"
#include "stream_host.hpp"

namespace sz3b {

// ---------------------------------------------------------------------------------------------------------------------
// workspace pool
// ---------------------------------------------------------------------------------------------------------------------
static std::mutex g_pool_mu;
static std::vector<Workspace *> g_pool;

Workspace *workspace_acquire() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) throw CudaError{e, "cudaGetDevice (is a CUDA device visible? this library has no CPU path)", __FILE__, __LINE__};
    {
        // last in, first out: a call releases its own workspace after the ones its tuner trials borrowed, so the next
        // call gets the workspace whose buffers already have the size of a whole array (the trial workspaces stay
        // small; handing one of them to a call costs device and pinned allocations of hundreds of megabytes)
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (size_t i = g_pool.size(); i-- > 0;) {
            if (g_pool[i]->device == dev) {
                Workspace *ws = g_pool[i];
                g_pool.erase(g_pool.begin() + i);
                ws->prof_reset();
                return ws;
            }
        }
    }
    Workspace *ws = new Workspace();
    ws->device = dev;
    SZ3B_CUDA(cudaStreamCreateWithFlags(&ws->st, cudaStreamNonBlocking));
    SZ3B_CUDA(cudaStreamCreateWithFlags(&ws->st_copy, cudaStreamNonBlocking));
    SZ3B_CUDA(cudaEventCreateWithFlags(&ws->ev_copy, cudaEventDisableTiming));
    return ws;
}

void workspace_release(Workspace *ws) {
    // a call that failed half-way may still have its background copy / kernels in flight on this workspace's buffers
    if (ws->st_copy) stream_wait(ws->st_copy);
    if (ws->st_low) stream_wait(ws->st_low);
    if (ws->st) stream_wait(ws->st);
    cudaGetLastError();
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool.push_back(ws);
}

// ---------------------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------------------
// Pageable host memory -> device: a cudaMemcpy from pageable memory runs at the speed of the driver's single staging
// thread (a fraction of the link rate).  Here the host worker pool copies 8 MiB pieces into a ring of pinned buffers
// and every piece crosses the link as soon as it is staged: the array travels at min(host memcpy rate, link rate).
namespace {
struct StageCopy {
    uint8_t *dst;
    const uint8_t *src;
    size_t len;
    int parts;
};
void stage_copy_part(void *arg, int worker) {
    StageCopy &j = *static_cast<StageCopy *>(arg);
    const size_t per = ((j.len + j.parts - 1) / j.parts + 63) & ~static_cast<size_t>(63);
    const size_t a = std::min(j.len, per * worker), b = std::min(j.len, a + per);
    if (b > a) memcpy(j.dst + a, j.src + a, b - a);
}
}  // namespace
static void upload_pageable(Workspace &ws, void *d_dst, const void *h_src, size_t bytes) {
    constexpr size_t kPiece = static_cast<size_t>(8) << 20;
    constexpr int kRing = 4;
    uint8_t *ring = static_cast<uint8_t *>(ws.upload_ring.ensure(kPiece * kRing));
    cudaEvent_t ev[kRing];
    for (int i = 0; i < kRing; i++) ev[i] = ws.event();
    const int threads = std::max(1, std::min(host_threads(), 8));
    size_t off = 0;
    for (int k = 0; off < bytes; k++, off += kPiece) {
        const size_t len = std::min(kPiece, bytes - off);
        uint8_t *slot = ring + static_cast<size_t>(k % kRing) * kPiece;
        if (k >= kRing) SZ3B_CUDA(event_wait(ev[k % kRing]));   // the slot's previous piece has left
        StageCopy j{slot, static_cast<const uint8_t *>(h_src) + off, len, threads};
        host_parallel(threads, stage_copy_part, &j);
        SZ3B_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(d_dst) + off, slot, len, cudaMemcpyHostToDevice, ws.st));
        SZ3B_CUDA(cudaEventRecord(ev[k % kRing], ws.st));
    }
    ws.h2d_bytes += bytes;
}

// ... and the way back: device -> pageable host memory through the same ring (the pieces are copied out by the pool
// while the next ones cross the link).  Returns when `h_dst` is complete.
static void download_pageable(Workspace &ws, void *h_dst, const void *d_src, size_t bytes) {
    constexpr size_t kPiece = static_cast<size_t>(8) << 20;
    constexpr int kRing = 4;
    uint8_t *ring = static_cast<uint8_t *>(ws.upload_ring.ensure(kPiece * kRing));
    cudaEvent_t ev[kRing];
    for (int i = 0; i < kRing; i++) ev[i] = ws.event();
    const int threads = std::max(1, std::min(host_threads(), 8));
    const size_t npieces = (bytes + kPiece - 1) / kPiece;
    auto issue = [&](size_t k) {
        const size_t off = k * kPiece, len = std::min(kPiece, bytes - off);
        SZ3B_CUDA(cudaMemcpyAsync(ring + (k % kRing) * kPiece, static_cast<const uint8_t *>(d_src) + off, len, cudaMemcpyDeviceToHost, ws.st));
        SZ3B_CUDA(cudaEventRecord(ev[k % kRing], ws.st));
    };
    for (size_t k = 0; k < npieces && k < static_cast<size_t>(kRing - 1); k++) issue(k);
    for (size_t k = 0; k < npieces; k++) {
        if (k + kRing - 1 < npieces) issue(k + kRing - 1);   // (its slot was emptied in the previous iteration)
        SZ3B_CUDA(event_wait(ev[k % kRing]));
        const size_t off = k * kPiece, len = std::min(kPiece, bytes - off);
        StageCopy j{static_cast<uint8_t *>(h_dst) + off, ring + (k % kRing) * kPiece, len, threads};
        host_parallel(threads, stage_copy_part, &j);
    }
    ws.d2h_bytes += bytes;
}

template <class T>
static const T *to_device(Workspace &ws, const T *data, int loc, size_t num) {
    if (loc == SZ3B_DEVICE) return data;
    T *d = ws.data.as<T>(num);
    size_t h = ws.stage_begin("h2d_input");
    const size_t bytes = num * sizeof(T);
    static const bool plain = getenv("SZ3B_PLAIN_PAGEABLE_H2D") != nullptr;   // diagnostics
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, data) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();   // an unregistered pointer may leave an error behind on old drivers
    if (!pinned && !plain && bytes >= (static_cast<size_t>(32) << 20))
        upload_pageable(ws, d, data, bytes);
    else
        ws.h2d(d, data, bytes);
    ws.stage_end(h, 0);
    return d;
}

// Pinned (page-locked, device-mapped) host input: returns the device alias of `data`, or nullptr.
static const void *mapped_alias(const void *data) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, data) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (attr.type != cudaMemoryTypeHost || attr.devicePointer == nullptr) return nullptr;
    return attr.devicePointer;
}

// Bulk uploads of concurrent callers go up one after the other instead of sharing the link command by command: the
// caller whose array is complete then encodes and downloads while the next caller's array is still on its way (two
// callers interleaved would both see their last plane after two upload times).
namespace {
struct UploadChain {
    std::mutex mu;
    cudaEvent_t ring[4] = {nullptr, nullptr, nullptr, nullptr};
    int head = 0;
    bool any = false;
};
UploadChain &upload_chain(int device) {
    static UploadChain chains[64];
    return chains[device >= 0 && device < 64 ? device : 0];
}
}  // namespace
bool device_upload_busy(int device) {
    UploadChain &c = upload_chain(device);
    if (!c.mu.try_lock()) return true;   // someone is enqueueing a bulk upload right now
    const bool busy = c.any && c.ring[c.head] && cudaEventQuery(c.ring[c.head]) == cudaErrorNotReady;
    cudaGetLastError();
    c.mu.unlock();
    return busy;
}
namespace {
struct UploadTurn {   // holds the chain while one call enqueues its copies on its copy stream
    UploadChain &c;
    cudaStream_t st;
    UploadTurn(Workspace &ws) : c(upload_chain(ws.device)), st(ws.st_copy) {
        c.mu.lock();
        if (c.any) cudaStreamWaitEvent(st, c.ring[c.head], 0);
    }
    ~UploadTurn() {
        const int next = (c.head + 1) & 3;
        if (!c.ring[next]) cudaEventCreateWithFlags(&c.ring[next], cudaEventDisableTiming);
        if (c.ring[next] && cudaEventRecord(c.ring[next], st) == cudaSuccess) {
            c.head = next;
            c.any = true;
        }
        c.mu.unlock();
    }
};
}  // namespace

// Starts the H2D of the whole input on the copy stream and returns at once; ws.st waits for it via join_copy().
template <class T>
static const T *to_device_background(Workspace &ws, const T *data, size_t num) {
    T *d = ws.data.as<T>(num);
    UploadTurn turn(ws);
    // in pieces: the tuner's own small uploads share the H2D copy engine and can only slip in between commands
    const size_t bytes = num * sizeof(T), piece = static_cast<size_t>(8) << 20;
    for (size_t off = 0; off < bytes; off += piece)
        SZ3B_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t *>(d) + off, reinterpret_cast<const uint8_t *>(data) + off,
                                  std::min(piece, bytes - off), cudaMemcpyHostToDevice, ws.st_copy));
    ws.h2d_bytes += bytes;
    SZ3B_CUDA(cudaEventRecord(ws.ev_copy, ws.st_copy));
    ws.bulk_copy_in_flight = true;
    return d;
}
static void join_copy(Workspace &ws) {
    SZ3B_CUDA(cudaStreamWaitEvent(ws.st, ws.ev_copy, 0));
    ws.bulk_copy_in_flight = false;
    ws.copy_plan.active = false;
}

// Same, for a 3-D array that the interpolation tile path will consume: planes of the outermost dimension in the
// order the levels need them (Workspace::CopyPlan).  Coarse levels (stride >= 2) only touch even coordinates, so they
// can run once the even planes are there; the level-1 tiles of block-row b read planes [32 b, 32 b + 32].
template <class T>
static const T *to_device_planes(Workspace &ws, const T *data, const sz3b_config &conf) {
    const size_t num = config_num(conf);
    T *d = ws.data.as<T>(num);
    const size_t nz = conf.dims[0], plane = num / nz * sizeof(T);
    uint8_t *dst = reinterpret_cast<uint8_t *>(d);
    const uint8_t *src = reinterpret_cast<const uint8_t *>(data);
    UploadTurn turn(ws);
    // even planes, 32 per command (measured: 2-D copies of 8 planes lose 1 % of the link rate, of 32 none; the
    // tuner's small uploads do not queue behind them, they go through the SM copy path while this is in flight)
    const size_t n_even = (nz + 1) / 2;
    for (size_t k = 0; k < n_even; k += 32) {
        const size_t cnt = std::min<size_t>(32, n_even - k);
        SZ3B_CUDA(cudaMemcpy2DAsync(dst + 2 * k * plane, 2 * plane, src + 2 * k * plane, 2 * plane, plane, cnt,
                                    cudaMemcpyHostToDevice, ws.st_copy));
    }
    ws.copy_plan.ev_even = ws.event();
    SZ3B_CUDA(cudaEventRecord(ws.copy_plan.ev_even, ws.st_copy));
    // odd planes, one group per block-row of the finest level (blocks of 32, the last one closed at nz - 1)
    const size_t nrows = (nz - 1 + kInterpBlock - 1) / kInterpBlock;
    ws.copy_plan.ev_row.clear();
    for (size_t b = 0; b < nrows; b++) {
        const size_t z0 = b * kInterpBlock + 1;
        const size_t z_end = std::min<size_t>((b + 1) * kInterpBlock, nz - 1);   // last plane the block-row reads
        if (z0 <= z_end) {
            const size_t cnt = (z_end - z0) / 2 + 1;
            SZ3B_CUDA(cudaMemcpy2DAsync(dst + z0 * plane, 2 * plane, src + z0 * plane, 2 * plane, plane, cnt,
                                        cudaMemcpyHostToDevice, ws.st_copy));
        }
        cudaEvent_t e = ws.event();
        SZ3B_CUDA(cudaEventRecord(e, ws.st_copy));
        ws.copy_plan.ev_row.push_back(e);
    }
    ws.h2d_bytes += num * sizeof(T);
    SZ3B_CUDA(cudaEventRecord(ws.ev_copy, ws.st_copy));
    ws.bulk_copy_in_flight = true;
    ws.copy_plan.active = true;
    return d;
}

template <class T>
void minmax_stage(Workspace &ws, const T *data, int loc, size_t num, double *mn, double *mx) {
    const T *d = to_device(ws, data, loc, num);
    T *mm = ws.misc.as<T>(2);
    size_t h = ws.stage_begin("minmax");
    launch_minmax<T>(d, num, mm, ws.st);
    ws.stage_end(h, 2);
    T hmm[2];
    ws.d2h(hmm, mm, sizeof(hmm));
    SZ3B_CUDA(stream_wait(ws.st));
    *mn = static_cast<double>(hmm[0]);
    *mx = static_cast<double>(hmm[1]);
}

// calAbsErrorBound (Statistic.hpp:24-56); with `have_range` the scan is skipped (OMP path: the slabs' min / max were
// reduced by the caller; a range of 0 -- a constant field -- is a legitimate value and leads to the lossless path
// exactly as in the reference).  d_data is device resident.
template <class T>
static void resolve_abs_eb(Workspace &ws, sz3b_config &conf, const T *d_data, T range, bool have_range = false) {
    if (conf.errorBoundMode == SZ3B_EB_ABS) return;
    auto get_range = [&]() -> T {
        if (have_range) return range;
        double mn, mx;
        minmax_stage<T>(ws, d_data, SZ3B_DEVICE, config_num(conf), &mn, &mx);
        return static_cast<T>(static_cast<T>(mx) - static_cast<T>(mn));   // T subtraction as data_range()
    };
    switch (conf.errorBoundMode) {
        case SZ3B_EB_REL:
            conf.absErrorBound = conf.relErrorBound * get_range();
            break;
        case SZ3B_EB_PSNR: {
            double v1 = conf.psnrErrorBound + 10 * log10(1 - 2.0 / 3.0 * 0.99);
            double v2 = v1 / (-20);
            double v3 = pow(10, v2);
            conf.absErrorBound = get_range() * v3;
            break;
        }
        case SZ3B_EB_L2NORM:
            conf.absErrorBound = sqrt(3.0 / config_num(conf)) * conf.l2normErrorBound;
            break;
        case SZ3B_EB_ABS_AND_REL:
            conf.absErrorBound = std::min(conf.absErrorBound, conf.relErrorBound * get_range());
            break;
        case SZ3B_EB_ABS_OR_REL:
            conf.absErrorBound = std::max(conf.absErrorBound, conf.relErrorBound * get_range());
            break;
        default:
            fail(SZ3B_E_INVALID_ARGUMENT, "Error bound mode not supported");
    }
    conf.errorBoundMode = SZ3B_EB_ABS;
}

template <class T>
double abs_eb_stage(Workspace &ws, const sz3b_config &conf, const T *data, int loc) {
    sz3b_config c = conf;
    const T *d = c.errorBoundMode == SZ3B_EB_ABS || c.errorBoundMode == SZ3B_EB_L2NORM
                     ? data
                     : to_device(ws, data, loc, config_num(c));
    resolve_abs_eb<T>(ws, c, d, static_cast<T>(0));
    return c.absErrorBound;
}

// ---------------------------------------------------------------------------------------------------------------------
// interpolation decomposition on the device
// ---------------------------------------------------------------------------------------------------------------------
struct IndexStream {       // device-resident result of a decomposition
    void *q = nullptr;     // QT[n]
    void *unpred_tmp = nullptr;
    unsigned long long *hist = nullptr;
    uint64_t n = 0;
    int radius = 0;
    bool wide = false;     // QT = uint32 (radius > 32768)
};

// Box schedule (interp_box.cu).  BoxPlan says, per level, where the level's input values come from: the input array
// (stride 1 by TMA, stride 2 by per-element copies) or a compact copy of the stride-4 / 8 / 16 lattices made by one
// whole-GPU gather (a single coarse tile would otherwise gather its ~36 k scattered sectors with one SM).
struct BoxPlan {
    bool any = false;
    float *compact[3] = {nullptr, nullptr, nullptr};   // lattices of stride 4, 8, 16
    int ncompact = 0;
};

template <class T, class QT>
static bool box_applicable(const InterpArgs<T, QT> &A) {
    if constexpr (std::is_same<T, float>::value && std::is_same<QT, uint16_t>::value)
        return interp_box_applicable(A);
    else
        return false;
}

template <class T, class QT>
static void box_prepare(Workspace &ws, const InterpPlan &pl, InterpArgs<T, QT> A, uint32_t nbatch, BoxPlan &bp) {
    if constexpr (std::is_same<T, float>::value && std::is_same<QT, uint16_t>::value) {
        if (!(pl.box && pl.tile && nbatch == 1)) return;
        int need = 0;
        for (const LevelPlan &L : pl.levels) {
            A.s = L.s;
            if (!interp_box_applicable(A)) continue;
            bp.any = true;
            for (int k = 0; k < 3; k++)
                if (L.s == (4u << k)) need = std::max(need, k + 1);
        }
        if (need) {
            size_t total = 0, off[3];
            for (int k = 0; k < need; k++) {
                size_t e = 1;
                for (int d = 0; d < 3; d++) e *= (pl.sh.dims[d] - 1) / (4u << k) + 1;
                off[k] = total;
                total += (e + 63) & ~static_cast<size_t>(63);
            }
            float *base = ws.compact.as<float>(total);
            for (int k = 0; k < need; k++) bp.compact[k] = base + off[k];
            bp.ncompact = need;
            interp_launch_compact(A.data, pl.sh.dims, pl.sh.stride, 4, need, bp.compact, ws.st);
        }
    }
}

// Tiles [A.tile0, A.tile0 + ntiles) of the level A.s through the box kernel; false = not this type / shape.
template <class T, class QT>
static bool launch_box(const InterpArgs<T, QT> &A, const BoxPlan &bp, uint64_t ntiles, cudaStream_t st) {
    if constexpr (std::is_same<T, float>::value && std::is_same<QT, uint16_t>::value) {
        if (!bp.any || !interp_box_applicable(A)) return false;
        BoxSrc S;
        uint32_t sdims[3];
        int k = -1;
        for (int j = 0; j < bp.ncompact; j++)
            if (A.s == (4u << j)) k = j;
        if (k >= 0) {
            for (int d = 0; d < 3; d++) sdims[d] = (A.sh.dims[d] - 1) / A.s + 1;
            S.p = bp.compact[k];
            S.st[2] = 1;
            S.st[1] = sdims[2];
            S.st[0] = static_cast<uint64_t>(sdims[1]) * sdims[2];
            for (int d = 0; d < 3; d++) S.ost[d] = S.st[d];
            S.odiv = A.s;
            S.tma = (sdims[2] & 3u) == 0;
        } else {
            for (int d = 0; d < 3; d++) {
                sdims[d] = A.sh.dims[d];
                S.ost[d] = A.sh.stride[d];
                S.st[d] = A.sh.stride[d] * A.s;
            }
            S.p = A.data;
            S.odiv = 1;
            S.tma = A.s == 1 && (sdims[2] & 3u) == 0 && (reinterpret_cast<uintptr_t>(A.data) & 15u) == 0;
        }
        // a level with fewer tiles than the GPU has CTA slots: four (eight) CTAs per tile, so that a tile's latency is
        // two (one) plane rounds instead of five (phase A is repeated by each of them; only the first emits it)
        S.split = ntiles <= 8 ? 8u : (ntiles * 4 <= 2u * 148u ? 4u : 1u);
        if (S.tma && interp_launch_box(A, S, sdims, ntiles, st)) return true;
        S.tma = 0;
        return interp_launch_box(A, S, sdims, ntiles, st);
    } else {
        return false;
    }
}

template <class T, class QT>
static void run_interp(Workspace &ws, const InterpPlan &pl, const T *d_data, uint32_t nbatch, int radius, QT *d_q,
                       T *d_unpred_tmp, unsigned long long *d_hist, DevBuf &recon_buf, int *launches,
                       uint64_t *hist_owed = nullptr) {
    InterpArgs<T, QT> A;
    memset(&A, 0, sizeof(A));
    A.sh = pl.sh;
    A.data = d_data;
    A.data_bstride = pl.num;
    A.q_bstride = pl.num;
    A.recon2_bstride = pl.num2;
    for (int d = 0; d < kMaxDim; d++) {
        A.dims2[d] = pl.dims2[d];
        A.stride2[d] = pl.stride2[d];
    }
    if (pl.tile) {
        A.recon2 = recon_buf.as<T>(pl.num2 * nbatch);
        A.work = nullptr;
    } else {
        A.work = recon_buf.as<T>(pl.num * nbatch);
        A.recon2 = nullptr;
    }
    A.q = d_q;
    A.unpred_tmp = d_unpred_tmp;
    A.hist = d_hist;
    uint64_t *d_table = ws.tables.as<uint64_t>(pl.table.size() + 1);
    if (!pl.table.empty())
        ws.h2d(d_table, pl.table.data(), pl.table.size() * sizeof(uint64_t));
    A.qp = make_quant(pl.eb, radius);
    A.s = 0;
    // input still arriving in plane order (to_device_planes): the line-walker tile path follows the copy, anything
    // else waits for all of it
    bool planes = ws.copy_plan.active && d_data == ws.data.p && nbatch == 1;
    if (planes && !(pl.tile && !pl.levels.empty() && pl.levels.back().s == 1 &&
                    pl.levels.back().nb[0] == ws.copy_plan.ev_row.size())) {
        join_copy(ws);
        planes = false;
    }
    if (planes) SZ3B_CUDA(cudaStreamWaitEvent(ws.st, ws.copy_plan.ev_even, 0));
    BoxPlan bp;
    box_prepare<T, QT>(ws, pl, A, nbatch, bp);
    if (bp.ncompact) (*launches)++;
    SZ3B_CUDA(cudaGetLastError());
    interp_launch_anchors<T, QT>(A, pl.anchor_stride, pl.n_first, nbatch, ws.st);
    (*launches)++;
    SZ3B_CUDA(cudaGetLastError());
    uint64_t hist_from = 0, hist_to = 0;
    // (stretches of consecutive levels are counted by one launch: hist_from .. hist_to is what is still owed)
    auto flush_hist = [&]() {
        if constexpr (std::is_same<QT, uint16_t>::value) {
            if (hist_to > hist_from) {
                launch_hist_u16(d_q + hist_from, hist_to - hist_from, radius, 2 * radius, d_hist, ws.st);
                (*launches)++;
            }
        }
        hist_from = hist_to = 0;
    };
    auto count_box = [&](uint64_t from, uint64_t to) {
        if (hist_to != from) flush_hist();
        if (hist_to == hist_from) hist_from = from;
        hist_to = to;
    };
    for (const LevelPlan &L : pl.levels) {
        A.qp = make_quant(L.eb, radius);
        A.s = L.s;
        for (int d = 0; d < kMaxDim; d++) A.nb[d] = L.nb[d];
        A.block_base = d_table + L.table_off;
        const bool box = bp.any && box_applicable<T, QT>(A);
        if (pl.box_required && L.s == 1 && !box) fail(SZ3B_E_UNSUPPORTED, "box schedule does not apply to this shape / type");
        // the box kernel does not count its indices: a pass over the stretch of the stream it wrote does (k_hist_u16)
        const uint64_t lv_begin = pl.table[L.table_off];
        const uint64_t lv_end = (&L == &pl.levels.back()) ? pl.num : pl.table[(&L + 1)->table_off];
        if (planes && L.s == 1) {
            // the finest level, block-row by block-row as the odd planes arrive
            const uint64_t per_row = static_cast<uint64_t>(L.nb[1]) * L.nb[2];
            for (uint32_t b = 0; b < L.nb[0]; b++) {
                SZ3B_CUDA(cudaStreamWaitEvent(ws.st, ws.copy_plan.ev_row[b], 0));
                A.tile0 = static_cast<uint32_t>(b * per_row);
                if (box && launch_box<T, QT>(A, bp, per_row, ws.st)) {
                    count_box(pl.table[L.table_off + b * per_row], b + 1 < L.nb[0] ? pl.table[L.table_off + (b + 1) * per_row] : lv_end);
                    flush_hist();   // counted while the next planes are still arriving
                } else if constexpr (std::is_floating_point<T>::value)
                    interp_launch_ltiles<T, QT>(A, per_row, nbatch, ws.st);
                SZ3B_CUDA(cudaGetLastError());
                (*launches)++;
            }
            A.tile0 = 0;
            join_copy(ws);
            continue;
        }
        if (pl.tile) {
            // (the tile schedules are floating-point kernels; integer element types run the per-pass kernels)
            if constexpr (std::is_floating_point<T>::value) {
                // the finest level is the dominant launch of the path (7/8 of the points): it gets a record of its
                // own inside the predict_quantize stage (bench.py: roofline.dominant_launch)
                const bool finest = L.s == 1 && nbatch == 1;
                const size_t hf = finest ? ws.stage_begin("predict_quantize_finest_level") : 0;
                if (box && launch_box<T, QT>(A, bp, L.nblocks, ws.st)) {
                    count_box(lv_begin, lv_end);
                } else {
                    interp_launch_ltiles<T, QT>(A, L.nblocks, nbatch, ws.st);
                }
                if (finest) ws.stage_end(hf, 0);
            } else {
                fail(SZ3B_E_RUNTIME, "tile schedule planned for an integer element type");
            }
            SZ3B_CUDA(cudaGetLastError());
            (*launches)++;
        } else {
            for (int p = 0; p < pl.sh.N; p++) {
                if (pass_points(A, p) == 0) continue;
                // the last pass of the finest level is read by nobody: its reconstructions need not be stored
                const bool write_work = !(L.s == 1 && p == pl.sh.N - 1);
                // (tuner batches of small cubes stay on the point-mapped kernel: a CTA of the row-mapped one would
                //  spend its time on the block table)
                bool done = false;
                if constexpr (std::is_floating_point<T>::value)
                    done = pl.lean && nbatch == 1 && interp_launch_lean<T, QT>(A, p, nbatch, write_work, false, nullptr, ws.st);
                if (!done) interp_launch_pass<T, QT>(A, p, nbatch, ws.st);
                (*launches)++;
            }
        }
    }
    // the stretch of the stream the box schedule left uncounted: handed to the caller (who times the histogram of
    // HuffmanEncoder::init as its own stage) or counted here
    if (hist_owed) {
        hist_owed[0] = hist_from;
        hist_owed[1] = hist_to;
    } else {
        flush_hist();
    }
    SZ3B_CUDA(cudaGetLastError());
}

// InterpolationDecomposition::save (:149-159) + LinearQuantizer::save (:95-104) up to (excluding) the unpred values
template <class T>
static size_t interp_save_header(const InterpPlan &pl, int radius, uint64_t n_unpred, uint8_t *out) {
    uint8_t *p = out;
    for (int d = 0; d < pl.sh.N; d++) put<uint64_t>(p, pl.sh.dims[d]);
    put<uint32_t>(p, kInterpBlock);
    put<int32_t>(p, pl.interp_id);
    put<int32_t>(p, pl.direction);
    put<uint64_t>(p, pl.anchor_stride);
    put<double>(p, pl.alpha);
    put<double>(p, pl.beta);
    put<uint8_t>(p, 2);  // LinearQuantizer uid
    put<double>(p, pl.eb);
    put<int32_t>(p, radius);
    put<uint64_t>(p, n_unpred);
    return static_cast<size_t>(p - out);
}

// ---------------------------------------------------------------------------------------------------------------------
// Huffman stage: histogram (already on the device) -> host tree -> GPU bit packer.  Produces, in pinned host memory
// at `dst`:  tree blob | size_t n | size_t outSize | bits ; and the unpredictable values (stream order) at unpred_dst.
// ---------------------------------------------------------------------------------------------------------------------
struct EncodeLayout {
    size_t tree_len = 0;
    size_t out_size = 0;     // bytes of Huffman bits
    uint64_t n_unpred = 0;
};

template <class QT, class T>
static void encode_indices(Workspace &ws, const QT *d_q, uint64_t n, const unsigned long long *d_hist, int nbins,
                           int sym_base, bool has_unpred, const T *d_unpred_tmp, HuffmanBook &book, EncodeLayout &lay) {
    unsigned long long *h_hist = static_cast<unsigned long long *>(ws.hist_host.ensure(sizeof(unsigned long long) * nbins));
    ws.d2h(h_hist, d_hist, sizeof(unsigned long long) * nbins);
    SZ3B_CUDA(stream_wait(ws.st));
    double t0 = now_ms();
    const char *err = nullptr;
    if (!huffman_build(h_hist, nbins, sym_base, book, &err))
        fail(strstr(err, "empty") ? SZ3B_E_INVALID_ARGUMENT : SZ3B_E_UNSUPPORTED, err);
    ws.host_stage("huffman_tree_host", now_ms() - t0);
    lay.tree_len = book.tree_blob.size();
    lay.out_size = (book.total_bits + 7) / 8;
    lay.n_unpred = has_unpred && sym_base == 0 ? h_hist[0] : 0;

    size_t h = ws.stage_begin("huffman_pack");
    const size_t states = book.state_num;
    unsigned long long *d_code = ws.code.as<unsigned long long>(states);
    uint8_t *d_len = ws.len.as<uint8_t>(states);
    ws.h2d(d_code, book.code.data(), states * sizeof(uint64_t));
    ws.h2d(d_len, book.len.data(), states);
    const uint64_t nchunks = pack_num_chunks(n);
    unsigned *d_cb = ws.chunk_bits.as<unsigned>(nchunks + 1);
    unsigned *d_cz = ws.chunk_zeros.as<unsigned>(nchunks + 1);
    unsigned long long *d_bo = ws.bit_off.as<unsigned long long>(nchunks + 2);
    unsigned long long *d_zo = ws.zero_off.as<unsigned long long>(nchunks + 2 + scan_scratch_words(nchunks));
    const size_t nwords = (book.total_bits + 31) / 32 + 2;
    unsigned *d_words = ws.out_words.as<unsigned>(nwords);
    SZ3B_CUDA(cudaMemsetAsync(d_words, 0, nwords * sizeof(unsigned), ws.st));
    T *d_unpred_out = lay.n_unpred ? ws.unpred_out.as<T>(lay.n_unpred) : nullptr;
    // the packer keeps a window of the code table around the most frequent state in shared memory
    size_t top = 0;
    for (size_t k = 0; k < states && book.offset - sym_base + k < static_cast<size_t>(nbins); k++)
        if (h_hist[book.offset - sym_base + k] > h_hist[book.offset - sym_base + top]) top = k;
    launch_pack<QT, T>(d_q, n, book.offset, 0, d_len, d_code, static_cast<unsigned>(states), static_cast<int>(top), d_cb, d_cz,
                       d_bo, d_zo, d_words, d_unpred_tmp, d_unpred_out, ws.st, nullptr,
                       scan_scratch_words(nchunks) ? d_zo + nchunks + 2 : nullptr);
    SZ3B_CUDA(cudaGetLastError());
    ws.stage_end(h, 3);
}

// D2H of the packed bits / unpredictables into an assembled (pinned) buffer laid out as SZGenericCompressor does:
//   decomposition.save | encoder.save | size_t n | size_t outSize | bits          (SZGenericCompressor.hpp:51-56)
// The bits stream in as pieces of kD2HPiece bytes, each followed by an event, so that the host zstd workers start on
// the first chunks while the rest is still crossing PCIe (ArrivalGate).
constexpr size_t kD2HPiece = static_cast<size_t>(4) << 20;

struct ArrivalGate : ZstdReady {
    std::vector<size_t> upto;          // buffer bytes [0, upto[j]) are valid once ev[j] has completed
    std::vector<cudaEvent_t> ev;
    std::atomic<size_t> done{0};       // pieces known to have arrived
    void wait(size_t need) override {
        size_t j = done.load(std::memory_order_acquire);
        while (j < upto.size() && (j == 0 ? 0 : upto[j - 1]) < need) {
            event_wait(ev[j]);
            j++;
        }
        size_t cur = done.load(std::memory_order_relaxed);
        while (cur < j && !done.compare_exchange_weak(cur, j, std::memory_order_release)) {
        }
    }
};

template <class T>
static size_t assemble_stream(Workspace &ws, const uint8_t *decomp_hdr, size_t decomp_hdr_len, const EncodeLayout &lay,
                              const HuffmanBook &book, uint64_t n, PinBuf &dstbuf, uint8_t **dst_out,
                              ArrivalGate &gate) {
    const size_t total = decomp_hdr_len + lay.n_unpred * sizeof(T) + lay.tree_len + 16 + lay.out_size;
    uint8_t *dst = static_cast<uint8_t *>(dstbuf.ensure(total + 16));
    uint8_t *p = dst;
    memcpy(p, decomp_hdr, decomp_hdr_len);
    p += decomp_hdr_len;
    size_t h = ws.stage_begin("d2h_stream");
    if (lay.n_unpred)
        ws.d2h(p, ws.unpred_out.p, lay.n_unpred * sizeof(T));
    p += lay.n_unpred * sizeof(T);
    memcpy(p, book.tree_blob.data(), lay.tree_len);
    p += lay.tree_len;
    put<uint64_t>(p, n);
    put<uint64_t>(p, lay.out_size);
    const size_t bits_off = static_cast<size_t>(p - dst);
    const uint8_t *d_bits = static_cast<const uint8_t *>(ws.out_words.p);
    size_t off = 0;
    while (off < lay.out_size) {
        // pieces end on multiples of kD2HPiece of the assembled buffer
        size_t end_abs = ((bits_off + off) / kD2HPiece + 1) * kD2HPiece;
        size_t len = std::min(lay.out_size - off, end_abs - (bits_off + off));
        ws.d2h(p + off, d_bits + off, len);
        off += len;
        cudaEvent_t e = ws.event();
        SZ3B_CUDA(cudaEventRecord(e, ws.st));
        gate.ev.push_back(e);
        gate.upto.push_back(bits_off + off);
    }
    if (gate.ev.empty() || gate.upto.back() < total) {
        cudaEvent_t e = ws.event();
        SZ3B_CUDA(cudaEventRecord(e, ws.st));
        gate.ev.push_back(e);
        gate.upto.push_back(total);
    }
    ws.stage_end(h, 0);
    *dst_out = dst;
    return total;
}

struct TooSmall {};   // stands in for std::length_error(SZ3_ERROR_COMP_BUFFER_NOT_LARGE_ENOUGH)

static size_t zstd_stage(Workspace &ws, const uint8_t *src, size_t len, uint8_t *dst, size_t cap, int threads,
                         ZstdReady *gate) {
    double t0 = now_ms();
    bool small = false;
    // gate != nullptr <=> the source is a packed (Huffman-coded) stream, not raw data
    size_t r = zstd_compress_framed(src, len, dst, cap, threads, &small, gate, &ws.zscratch, gate != nullptr);
    SZ3B_CUDA(stream_wait(ws.st));
    ws.host_stage("zstd_host", now_ms() - t0);
    if (small) throw TooSmall{};
    if (r == 0) fail(SZ3B_E_RUNTIME, "zstd compression failed");
    return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// Lossless stage on the GPU (lossless policy 2, zhuf.cuh): the stream SZGenericCompressor hands to Lossless_zstd
// (SZGenericCompressor.hpp:51-60) is assembled in device memory, coded into zstd frames of Huffman-only literal blocks
// by zhuf_kernels.cu, and only the compressed frames cross PCIe.  Same layout as Lossless_zstd::compress
// (Lossless_zstd.hpp:29-37): size_t srcLen | zstd frames.
// ---------------------------------------------------------------------------------------------------------------------
constexpr size_t kZhufMinStream = static_cast<size_t>(4) << 20;   // shorter streams keep the reference's own zstd call

template <class T>
static size_t stream_len(size_t decomp_hdr_len, const EncodeLayout &lay) {
    return decomp_hdr_len + lay.n_unpred * sizeof(T) + lay.tree_len + 16 + lay.out_size;
}

struct CopyJob {
    uint8_t *dst;
    const uint8_t *src;
    size_t len;
    int parts;
};
static void copy_part(void *arg, int worker) {
    CopyJob &j = *static_cast<CopyJob *>(arg);
    const size_t per = (j.len + j.parts - 1) / j.parts;
    const size_t a = std::min(j.len, per * worker), b = std::min(j.len, a + per);
    if (b > a) memcpy(j.dst + a, j.src + a, b - a);
}

// device -> caller memory: straight into pinned / registered memory, through the pinned staging buffer and the host
// worker pool otherwise (a pageable cudaMemcpy would run at a fraction of the link rate)
static void deliver_d2h(Workspace &ws, uint8_t *dst, const uint8_t *d_src, size_t bytes) {
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, dst) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();   // an unregistered pointer may leave an error behind on old drivers
    if (pinned) {
        ws.d2h(dst, d_src, bytes);
        SZ3B_CUDA(stream_wait(ws.st));
        return;
    }
    uint8_t *stage = static_cast<uint8_t *>(ws.stage.ensure(bytes + 16));
    ws.d2h(stage, d_src, bytes);
    SZ3B_CUDA(stream_wait(ws.st));
    CopyJob j{dst, stage, bytes, std::max(1, std::min(host_threads(), 8))};
    if (bytes < (1u << 20))
        memcpy(dst, stage, bytes);
    else
        host_parallel(j.parts, copy_part, &j);
}

// d_src[0, len) -> dst (host) as zhuf frames; returns their size.  The tables of all blocks are built by one launch
// (its time is one block's latency); offsets and bit streams then follow in up to four slices of whole frames, all
// queued at once, and while slice j + 1 is being written the frames of slice j cross PCIe on the copy stream (straight
// into `dst` when it is pinned; otherwise one slice and a staged copy).
static size_t zhuf_run(Workspace &ws, const uint8_t *d_src, size_t len, uint8_t *dst, size_t cap) {
    const uint64_t nblocks = zhuf_num_blocks(len);
    ZhufBlockInfo *d_info = ws.zinfo.as<ZhufBlockInfo>(nblocks + 1);
    uint8_t *d_out = ws.zdst.as<uint8_t>(zhuf_bound(len));
    unsigned long long *d_total = ws.counters.as<unsigned long long>(4) + 3;
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, dst) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const uint64_t nframes = (nblocks + kZhufBlocksPerFrame - 1) / kZhufBlocksPerFrame;
    // a payload that will be placed later (ws.placer) stays on the device until its size is known: one slice
    Placer *const placer = ws.placer && !ws.placer->used ? ws.placer : nullptr;
    const int nslices = pinned && nframes >= 16 && !placer ? 4 : 1;
    unsigned long long *h_log = static_cast<unsigned long long *>(ws.hist_host.ensure(64));
    cudaEvent_t ev[4];
    size_t h = ws.stage_begin("lossless_gpu");
    SZ3B_CUDA(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), ws.st));
    launch_zhuf_build(d_src, len, d_info, ws.st);
    for (int j = 0; j < nslices; j++) {
        const uint64_t f0 = nframes * j / nslices, f1 = nframes * (j + 1) / nslices;
        const uint64_t g0 = f0 * kZhufBlocksPerFrame, g1 = std::min<uint64_t>(f1 * kZhufBlocksPerFrame, nblocks);
        // the running total goes straight into pinned (device-mapped) host memory: a copy command would queue behind
        // the previous slice's frames on the D2H engine
        launch_zhuf_emit(d_src, len, g0, g1, d_info, d_out, d_total, h_log + j, ws.st);
        ev[j] = ws.event();
        SZ3B_CUDA(cudaEventRecord(ev[j], ws.st));
    }
    ws.stage_end(h, 1 + 2 * nslices);
    double t0 = now_ms();
    unsigned long long done = 0;
    bool small = false;
    for (int j = 0; j < nslices; j++) {
        SZ3B_CUDA(event_wait(ev[j]));
        const unsigned long long end = h_log[j];
        if (end > zhuf_bound(len) || end < done) fail(SZ3B_E_RUNTIME, "GPU lossless stage produced an invalid size");
        if (end > cap) {
            small = true;
            break;
        }
        if (nslices > 1) {
            SZ3B_CUDA(cudaMemcpyAsync(dst + done, d_out + done, end - done, cudaMemcpyDeviceToHost, ws.st_copy));
            ws.d2h_bytes += end - done;
        }
        done = end;
    }
    SZ3B_CUDA(stream_wait(ws.st));
    if (nslices > 1) SZ3B_CUDA(stream_wait(ws.st_copy));
    SZ3B_CUDA(cudaGetLastError());
    if (small) throw TooSmall{};
    if (placer && done && placer->raw_bytes / static_cast<double>(done + sizeof(uint64_t)) >= 3) {
        // (a lossy result below ratio 3 may still be replaced by the lossless one, SZDispatcher.hpp:62-74: that one is
        //  placed by dispatch_compress after the decision)
        uint8_t *f = placer->place(done + sizeof(uint64_t));
        placer->used = true;
        uint8_t *q = f;
        put<uint64_t>(q, static_cast<uint64_t>(len));
        deliver_d2h(ws, q, d_out, done);
        ws.host_stage("d2h_compressed", now_ms() - t0);
        return done;
    }
    if (nslices == 1 && done) deliver_d2h(ws, dst, d_out, done);
    ws.host_stage("d2h_compressed", now_ms() - t0);
    return done;
}

// stage-level entry (sz3b_lossless_compress): any byte buffer through the GPU lossless stage
size_t lossless_gpu_stage(Workspace &ws, const uint8_t *src, size_t len, int loc, uint8_t *out, size_t cap) {
    if (cap < sizeof(uint64_t) + zhuf_bound(len)) fail(SZ3B_E_INVALID_ARGUMENT, "output buffer too small");
    const uint8_t *d_src = src;
    if (loc == SZ3B_HOST) {
        uint8_t *d = ws.zsrc.as<uint8_t>(len + 64);
        if (len) SZ3B_CUDA(cudaMemcpyAsync(d, src, len, cudaMemcpyHostToDevice, ws.st));
        ws.h2d_bytes += len;
        d_src = d;
    }
    uint8_t *p = out;
    put<uint64_t>(p, static_cast<uint64_t>(len));
    size_t csize = 0;
    try {
        csize = zhuf_run(ws, d_src, len, p, cap - sizeof(uint64_t));
    } catch (TooSmall &) {
        fail(SZ3B_E_INVALID_ARGUMENT, "output buffer too small");
    }
    return sizeof(uint64_t) + csize;
}

template <class T>
static size_t zhuf_stage(Workspace &ws, const uint8_t *decomp_hdr, size_t decomp_hdr_len, const EncodeLayout &lay,
                         const HuffmanBook &book, uint64_t n, uint8_t *dst, size_t cap) {
    const size_t total = stream_len<T>(decomp_hdr_len, lay);
    // the reference's capacity rule (Lossless_zstd.hpp:30-33 -> std::length_error), independent of the policy
    if (cap < sizeof(uint64_t) || cap - sizeof(uint64_t) < ZSTD_compressBound(total)) throw TooSmall{};
    size_t h = ws.stage_begin("lossless_assemble");
    uint8_t *d_src = ws.zsrc.as<uint8_t>(total + 64);
    // small host-built pieces go up, the bulk (unpredictable values, packed bits) is already on the device
    size_t off = 0;
    ws.h2d(d_src, decomp_hdr, decomp_hdr_len);
    off += decomp_hdr_len;
    if (lay.n_unpred)
        SZ3B_CUDA(cudaMemcpyAsync(d_src + off, ws.unpred_out.p, lay.n_unpred * sizeof(T), cudaMemcpyDeviceToDevice, ws.st));
    off += lay.n_unpred * sizeof(T);
    std::vector<uint8_t> mid(lay.tree_len + 16);
    memcpy(mid.data(), book.tree_blob.data(), lay.tree_len);
    {
        uint8_t *p = mid.data() + lay.tree_len;
        put<uint64_t>(p, n);
        put<uint64_t>(p, lay.out_size);
    }
    uint8_t *mid_pin = static_cast<uint8_t *>(ws.stage2.ensure(mid.size() + 16));
    memcpy(mid_pin, mid.data(), mid.size());
    SZ3B_CUDA(cudaMemcpyAsync(d_src + off, mid_pin, mid.size(), cudaMemcpyHostToDevice, ws.st));
    ws.h2d_bytes += mid.size();
    off += mid.size();
    if (lay.out_size)
        SZ3B_CUDA(cudaMemcpyAsync(d_src + off, ws.out_words.p, lay.out_size, cudaMemcpyDeviceToDevice, ws.st));
    ws.stage_end(h, 0);
    uint8_t *p = dst;
    put<uint64_t>(p, static_cast<uint64_t>(total));
    const size_t csize = zhuf_run(ws, d_src, total, p, cap - sizeof(uint64_t));
    if (csize == 0) fail(SZ3B_E_RUNTIME, "GPU lossless stage produced an invalid size");
    return sizeof(uint64_t) + csize;
}

// ---------------------------------------------------------------------------------------------------------------------
// SZ_compress_Interp (SZAlgoInterp.hpp:17-30) for `nbatch` arrays of identical shape (nbatch > 1 only in the tuner,
// where the index streams of all sampled cubes are merged before encoding, :50-56).
// ---------------------------------------------------------------------------------------------------------------------
template <class T, class QT>
static size_t interp_compress_t(Workspace &ws, const sz3b_config &conf, const T *d_data, uint32_t nbatch, uint8_t *dst,
                                size_t cap, int zstd_threads, bool tuner) {
    InterpPlan pl;
    // SZ3B_SCHEDULE (diagnostics): force one of the schedules of sz3b_interp_decompose for whole compressions
    static const int forced = [] {
        const char *e = getenv("SZ3B_SCHEDULE");
        return e ? atoi(e) : 0;
    }();
    const int schedule = std::is_integral<T>::value
                             ? 1   // integer element types: the per-pass kernels (core.cuh carries their arithmetic)
                             : ((forced == 1 || (forced == 4 && conf.N == 3) || (forced == 5 && conf.N >= 3)) ? forced : 0);
    const double t_plan = now_ms();
    if (const char *e = build_interp_plan(conf, conf.absErrorBound, schedule, pl)) fail(SZ3B_E_INVALID_ARGUMENT, e);
    if (!tuner) ws.host_stage("plan_host", now_ms() - t_plan);
    const int radius = conf.quantbinCnt / 2;
    const int nbins = 2 * radius;
    const uint64_t n = pl.num * nbatch;
    DevBuf &qb = tuner ? ws.cube_q : ws.q;
    DevBuf &ub = tuner ? ws.cube_unpred : ws.unpred_tmp;
    DevBuf &rb = tuner ? ws.cube_recon : ws.recon;
    QT *d_q = qb.as<QT>(n);
    T *d_unpred_tmp = ub.as<T>(n);
    unsigned long long *d_hist = ws.hist.as<unsigned long long>(nbins);
    size_t h = ws.stage_begin(tuner ? "tune_predict_quantize" : "predict_quantize");
    SZ3B_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nbins, ws.st));
    int launches = 0;
    uint64_t owed[2] = {0, 0};
    run_interp<T, QT>(ws, pl, d_data, nbatch, radius, d_q, d_unpred_tmp, d_hist, rb, &launches, owed);
    ws.stage_end(h, launches);
    if constexpr (std::is_same<QT, uint16_t>::value) {
        if (owed[1] > owed[0]) {   // HuffmanEncoder::init's histogram (:516-527) of what the box schedule wrote
            h = ws.stage_begin(tuner ? "tune_huffman_histogram" : "huffman_histogram");
            launch_hist_u16(d_q + owed[0], owed[1] - owed[0], radius, nbins, d_hist, ws.st);
            ws.stage_end(h, 1);
        }
    }

    HuffmanBook book;
    EncodeLayout lay;
    encode_indices<QT, T>(ws, d_q, n, d_hist, nbins, 0, true, d_unpred_tmp, book, lay);
    uint8_t hdr[128];
    size_t hdr_len = interp_save_header<T>(pl, radius, lay.n_unpred, hdr);
    if (!tuner && lossless_policy() == 2 && stream_len<T>(hdr_len, lay) >= kZhufMinStream)
        return zhuf_stage<T>(ws, hdr, hdr_len, lay, book, n, dst, cap);
    uint8_t *buf = nullptr;
    ArrivalGate gate;
    size_t len = assemble_stream<T>(ws, hdr, hdr_len, lay, book, n, tuner ? ws.stage2 : ws.stage, &buf, gate);
    return zstd_stage(ws, buf, len, dst, cap, zstd_threads, &gate);
}

template <class T>
static size_t interp_compress(Workspace &ws, const sz3b_config &conf, const T *d_data, uint32_t nbatch, uint8_t *dst,
                              size_t cap, int zstd_threads, bool tuner) {
    if (conf.quantbinCnt < 2) fail(SZ3B_E_INVALID_ARGUMENT, "quantbinCnt must be >= 2");
    if (conf.quantbinCnt / 2 <= 32768)
        return interp_compress_t<T, uint16_t>(ws, conf, d_data, nbatch, dst, cap, zstd_threads, tuner);
    return interp_compress_t<T, uint32_t>(ws, conf, d_data, nbatch, dst, cap, zstd_threads, tuner);
}

static void set_default_anchor(sz3b_config &conf) {
    if (conf.interpAnchorStride < 0) {
        static const int def[4] = {4096, 128, 32, 16};
        conf.interpAnchorStride = def[conf.N - 1];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// auto-tuner (SZAlgoInterp.hpp:122-286).  Returns true when conf was tuned (cmprAlgo is then ALGO_INTERP or
// ALGO_LORENZO_REG); all trial compressions run through the same GPU kernels on the sampled cubes.
// ---------------------------------------------------------------------------------------------------------------------
template <class T>
static size_t quantizer_header(double eb, int radius, uint64_t n_unpred, uint8_t *out);

// lorenzo_compress_test (SZAlgoInterp.hpp:78-119), defined with the Lorenzo stack further down
template <class T>
static double lorenzo_trial(Workspace &ws, const sz3b_config &lc, const T *d_cubes, uint32_t ncubes, size_t per_block,
                            uint8_t *out, size_t cap);

template <class T>
static void tune_interp(Workspace &ws, sz3b_config &conf, const T *d_data) {
    const int N = conf.N;
    const uint64_t num = config_num(conf);
    set_default_anchor(conf);
    const sz3b_config conf_entry = conf;   // `Config lorenzo_config = conf;` (:182) is taken before the interp tuning
    const double sampleRate = 0.005;
    static const size_t sbs_def[4] = {4096, 128, 32, 16};
    size_t sbs = sbs_def[N - 1];
    size_t shortest = conf.dims[0];
    for (int i = 0; i < N; i++) shortest = std::min<size_t>(shortest, conf.dims[i]);
    while (sbs >= shortest) sbs /= 2;
    while (sbs >= 16 && (pow(sbs + 1, N) / num) > 1.5 * sampleRate) sbs /= 2;
    if (sbs < 8) sbs = 8;
    bool to_tune = pow(sbs + 1, N) <= 0.05 * num;
    for (int i = 0; i < N; i++)
        if (conf.dims[i] < sbs) {
            to_tune = false;
            break;
        }
    if (!to_tune) {
        conf.cmprAlgo = SZ3B_ALGO_INTERP;
        return;
    }
    const size_t per_block = static_cast<size_t>(pow(sbs + 1, N));
    // ---- profiling_block (Sample.hpp:9-127): non-constant candidate blocks, row-major --------------------------------
    uint32_t dims32[4] = {1, 1, 1, 1};
    uint64_t stride[4] = {0, 0, 0, 0};
    uint64_t acc = 1;
    for (int d = N - 1; d >= 0; d--) {
        dims32[d] = static_cast<uint32_t>(conf.dims[d]);
        stride[d] = acc;
        acc *= conf.dims[d];
    }
    uint64_t cb[4] = {1, 1, 1, 1};
    uint64_t ncand = 1;
    for (int d = 0; d < N; d++) {
        // i = 0, sbs, ... while i < dim - sbs
        cb[d] = conf.dims[d] > sbs ? (conf.dims[d] - sbs - 1) / sbs + 1 : 0;
        ncand *= cb[d];
    }
    std::vector<uint64_t> starts;   // flattened element offset of each selected cube
    auto cand_offset = [&](uint64_t b) {
        uint64_t off = 0;
        for (int d = N - 1; d >= 0; d--) {
            off += (b % cb[d]) * sbs * stride[d];
            b /= cb[d];
        }
        return off;
    };
    std::vector<uint64_t> filtered;
    if (ncand > 0) {
        uint8_t *d_flags = ws.flags.as<uint8_t>(ncand);
        size_t h = ws.stage_begin("tune_profile_blocks");
        launch_profile_blocks<T>(d_data, N, dims32, static_cast<uint32_t>(sbs), static_cast<uint32_t>(sbs / 4),
                                 conf.absErrorBound, d_flags, ncand, ws.st);
        ws.stage_end(h, 1);
        std::vector<uint8_t> flags(ncand);
        ws.d2h(flags.data(), d_flags, ncand);
        SZ3B_CUDA(stream_wait(ws.st));
        for (uint64_t b = 0; b < ncand; b++)
            if (flags[b]) filtered.push_back(b);
    }
    const size_t nfilt = filtered.size();
    const bool profiling = nfilt * per_block >= 0.5 * sampleRate * num;
    // ---- sampleBlocks (Sample.hpp:202-289) -----------------------------------------------------------------------------
    {
        size_t totalblock = 1;
        for (int d = 0; d < N; d++) totalblock *= static_cast<int>((conf.dims[d] - 1) / sbs);
        if (profiling) {
            size_t sstride = static_cast<size_t>(nfilt / (totalblock * sampleRate));
            if (sstride <= 0) sstride = 1;
            for (size_t i = 0; i < nfilt; i += sstride) starts.push_back(cand_offset(filtered[i]));
        } else {
            size_t sstride = static_cast<size_t>(1.0 / sampleRate);
            if (sstride <= 0) sstride = 1;
            for (uint64_t idx = 0; idx < ncand; idx++)   // the nested start loops enumerate the same candidates
                if (idx % sstride == 0) starts.push_back(cand_offset(idx));
        }
    }
    const size_t sampling_num = starts.size() * per_block;
    if (sampling_num == 0 || sampling_num >= num * 0.2) {
        conf.cmprAlgo = SZ3B_ALGO_INTERP;
        return;
    }
    const uint32_t ncubes = static_cast<uint32_t>(starts.size());
    if (ncubes > 65535) fail(SZ3B_E_UNSUPPORTED, "tuner sample exceeds 65535 cubes");
    T *d_cubes = ws.cubes.as<T>(sampling_num);
    {
        uint64_t *d_starts = ws.starts.as<uint64_t>(ncubes);
        ws.h2d(d_starts, starts.data(), ncubes * sizeof(uint64_t));
        size_t h = ws.stage_begin("tune_gather");
        launch_gather_cubes<T>(d_data, N, dims32, static_cast<uint32_t>(sbs + 1), d_starts, ncubes, d_cubes, ws.st);
        ws.stage_end(h, 1);
        SZ3B_CUDA(stream_wait(ws.st));   // `starts` must outlive the copy
    }
    std::vector<uint8_t> &trial_out = ws.trial_out;   // kept across calls (zero-filling megabytes per call costs a trial)
    {
        size_t need = std::max<size_t>(sizeof(T) * sampling_num + (1u << 20),
                                       ZSTD_compressBound(sizeof(T) * sampling_num + 65536 * 16) + 64);
        if (trial_out.size() < need) trial_out.resize(need);
    }
    // Trial compressions that do not depend on each other's outcome run concurrently: the first on this call's
    // workspace, the others on host threads with workspaces of their own (own stream, own buffers; the sampled cubes
    // are shared read-only).  The decisions below are then taken in the reference's order.
    const size_t trial_cap = trial_out.size();
    auto run_trials = [&](const std::vector<sz3b_config> &tcs) -> std::vector<double> {
        std::vector<double> ratios(tcs.size(), 0.0);
        std::vector<std::exception_ptr> errs(tcs.size());
        auto one = [&](Workspace &w, size_t k, uint8_t *out) {
            w.stage_prefix = "tune_";
            size_t sz = interp_compress<T>(w, tcs[k], d_cubes, ncubes, out, trial_cap, 1, true);
            w.stage_prefix.clear();
            ratios[k] = per_block * static_cast<double>(ncubes) * sizeof(T) * 1.0 / sz;
        };
        // persistent pool threads (stream_host.cpp): creating threads per call costs more than a trial
        struct Ctx {
            decltype(one) *fn;
            Workspace *main_ws;
            uint8_t *main_out;
            size_t cap;
            int device;
            bool bulk;
            std::vector<std::exception_ptr> *errs;
            size_t n, workers;
        } cx{&one, &ws, trial_out.data(), trial_cap, ws.device, ws.bulk_copy_in_flight, &errs, tcs.size(),
             std::min<size_t>(tcs.size(), static_cast<size_t>(std::max(1, host_threads())))};
        auto task = [](void *arg, int worker) {
            Ctx &c = *static_cast<Ctx *>(arg);
            for (size_t k = static_cast<size_t>(worker); k < c.n; k += c.workers) {
                try {
                    if (k == 0) {
                        (*c.fn)(*c.main_ws, 0, c.main_out);
                    } else {
                        SZ3B_CUDA(cudaSetDevice(c.device));
                        WorkspaceLease w2;
                        w2->bulk_copy_in_flight = c.bulk;
                        if (w2->trial_out.size() < c.cap) w2->trial_out.resize(c.cap);
                        (*c.fn)(*w2, k, w2->trial_out.data());
                        w2->bulk_copy_in_flight = false;
                    }
                } catch (...) {
                    (*c.errs)[k] = std::current_exception();
                }
            }
        };
        host_parallel(static_cast<int>(cx.workers), task, &cx);
        for (auto &e : errs)
            if (e) std::rethrow_exception(e);
        return ratios;
    };
    double best_interp = 0, best_lorenzo = 0;
    conf.interpDirection = 0;
    conf.interpAlpha = 1.25;
    conf.interpBeta = 2.0;
    sz3b_config tc = conf;
    {
        uint64_t cd[4];
        for (int d = 0; d < N; d++) cd[d] = sbs + 1;
        config_set_dims(tc, N, cd);
    }
    int fact = 1;
    for (int i = 2; i <= N; i++) fact *= i;
    // The reference tries {linear, cubic}, then the other direction with the winner, then three (alpha, beta) pairs with
    // the winner of that (SZAlgoInterp.hpp:176-224): six trial compressions in three dependent rounds.  A trial's ratio
    // depends on its own configuration only, so the candidates the later decisions may ask for are compressed ahead of
    // time, concurrently, and the reference's decision sequence is then replayed over the table of ratios: the same
    // decisions, one round of latency (all 16 candidates) where the host has the threads for it, two (4 + 3) otherwise.
    const double alphas[4] = {1.25, 1.0, 1.5, 2.0}, betas[4] = {2.0, 1.0, 2.5, 3.0};
    double table[2][2][4];
    bool have[2][2][4] = {};
    auto cfg_of = [&](int a, int d, int ab) {
        sz3b_config c = tc;
        c.interpAlgo = a ? SZ3B_INTERP_CUBIC : SZ3B_INTERP_LINEAR;
        c.interpDirection = d ? fact - 1 : 0;
        c.interpAlpha = alphas[ab];
        c.interpBeta = betas[ab];
        return c;
    };
    struct Key {
        int a, d, ab;
    };
    auto prefetch = [&](const std::vector<Key> &keys) {
        std::vector<Key> todo;
        std::vector<sz3b_config> tcs;
        for (const Key &k : keys)
            if (!have[k.a][k.d][k.ab] && !(k.d == 1 && fact == 1)) {
                todo.push_back(k);
                tcs.push_back(cfg_of(k.a, k.d, k.ab));
            }
        if (tcs.empty()) return;
        const std::vector<double> r = run_trials(tcs);
        for (size_t i = 0; i < todo.size(); i++) {
            table[todo[i].a][todo[i].d][todo[i].ab] = r[i];
            have[todo[i].a][todo[i].d][todo[i].ab] = true;
        }
    };
    auto ratio_of = [&](int a, int d, int ab) {
        if (fact == 1) d = 0;   // 1-D: the "other" direction is the same permutation
        if (!have[a][d][ab]) prefetch({Key{a, d, ab}});
        return table[a][d][ab];
    };
    const int threads = std::max(1, host_threads());
    if (threads >= 12) {
        std::vector<Key> all;
        for (int a = 0; a < 2; a++)
            for (int d = 0; d < 2; d++)
                for (int ab = 0; ab < 4; ab++) all.push_back(Key{a, d, ab});
        prefetch(all);
    } else if (threads >= 4) {
        prefetch({Key{0, 0, 0}, Key{1, 0, 0}, Key{0, 1, 0}, Key{1, 1, 0}});
    } else {
        prefetch({Key{0, 0, 0}, Key{1, 0, 0}});
    }
    int A = 1, D = 0;
    {   // interpolator (:176-184)
        for (int a = 0; a < 2; a++) {
            const double r = ratio_of(a, 0, 0);
            if (r > best_interp) {
                best_interp = r;
                A = a;
            }
        }
        conf.interpAlgo = A ? SZ3B_INTERP_CUBIC : SZ3B_INTERP_LINEAR;
    }
    {   // direction (:186-197)
        const double r = ratio_of(A, 1, 0);
        if (r > best_interp * 1.02) {
            best_interp = r;
            D = 1;
            conf.interpDirection = fact - 1;
        }
    }
    {   // (alpha, beta) (:199-224)
        prefetch({Key{A, D, 1}, Key{A, D, 2}, Key{A, D, 3}});
        for (int i = 1; i < 4; i++) {
            const double r = ratio_of(A, D, i);
            if (r > best_interp * 1.02) {
                best_interp = r;
                conf.interpAlpha = alphas[i];
                conf.interpBeta = betas[i];
            }
        }
    }
    sz3b_config lc = conf_entry;
    if (N == 1 && best_interp < 50) {   // only test lorenzo for 1D (:226-242)
        lc.cmprAlgo = SZ3B_ALGO_LORENZO_REG;
        uint64_t cd[4] = {sbs + 1, 0, 0, 0};
        config_set_dims(lc, N, cd);
        lc.lorenzo = 1;
        lc.lorenzo2 = 1;
        lc.regression = 0;
        lc.regression2 = 0;
        lc.openmp = 0;
        lc.blockSize = 5;
        ws.stage_prefix = "tune_";
        best_lorenzo = lorenzo_trial<T>(ws, lc, d_cubes, ncubes, per_block, trial_out.data(), trial_cap);
        ws.stage_prefix.clear();
    }
    const bool use_interp = !(best_lorenzo >= best_interp * 1.1 && best_lorenzo < 50 && best_interp < 50);   // :244-245
    if (use_interp) {
        conf.cmprAlgo = SZ3B_ALGO_INTERP;
        return;
    }
    if (conf.relErrorBound < 1.01e-6 && best_lorenzo > 5 && lc.quantbinCnt != 16384) {   // :268-277
        const int quant_num = lc.quantbinCnt;
        lc.quantbinCnt = 16384;
        ws.stage_prefix = "tune_";
        const double ratio = lorenzo_trial<T>(ws, lc, d_cubes, ncubes, per_block, trial_out.data(), trial_cap);
        ws.stage_prefix.clear();
        if (ratio > best_lorenzo * 1.02)
            best_lorenzo = ratio;
        else
            lc.quantbinCnt = quant_num;
    }
    config_set_dims(lc, N, conf.dims);   // :278 -- resets blockSize to the default of the rank
    conf = lc;                           // :279; SZ_compress_LorenzoReg follows in the dispatcher
}

template <class T>
void tune_stage(Workspace &ws, sz3b_config &conf, const T *data, int loc) {
    const T *d = to_device(ws, data, loc, config_num(conf));
    resolve_abs_eb<T>(ws, conf, d, static_cast<T>(0));
    tune_interp<T>(ws, conf, d);
}

// ---------------------------------------------------------------------------------------------------------------------
// BlockwiseDecomposition on the device (SZAlgoLorenzoReg.hpp:22-64, BlockwiseDecomposition.hpp:28-46,69-73).
// The regression-only stack runs on blockwise.cu (fit, speculative coefficient chain, fused predict+quantize); every
// stack with a Lorenzo predictor on the block wavefront of lorenzo.cu (run_blockwise_lorenzo below).
// ---------------------------------------------------------------------------------------------------------------------
void huffman_encode_device(Workspace &ws, const int32_t *d_q, size_t n, std::vector<uint8_t> &out, size_t *tree_len,
                           const std::function<void()> *after_hist = nullptr);

// LinearQuantizer::save (LinearQuantizer.hpp:95-104)
template <class T>
static void quantizer_save(std::vector<uint8_t> &out, double eb, int radius, const std::vector<T> &unpred) {
    uint8_t tmp[32];
    uint8_t *p = tmp;
    put<uint8_t>(p, 2);
    put<double>(p, eb);
    put<int32_t>(p, radius);
    put<uint64_t>(p, unpred.size());
    out.insert(out.end(), tmp, p);
    const uint8_t *v = reinterpret_cast<const uint8_t *>(unpred.data());
    out.insert(out.end(), v, v + unpred.size() * sizeof(T));
}

// Unpredictable coefficients listed by the chain kernels as (dense position, value), unordered: fetch the entries
// below `limit` (entries at or above it are leftovers of a discarded speculation).
template <class T>
static void fetch_coef_unpred(Workspace &ws, unsigned long long n_unp, const unsigned long long *upos, const T *uval,
                              unsigned long long limit, std::vector<std::pair<unsigned long long, T>> &out) {
    if (!n_unp) return;
    std::vector<unsigned long long> pos(n_unp);
    std::vector<T> val(n_unp);
    ws.d2h(pos.data(), upos, n_unp * sizeof(unsigned long long));
    ws.d2h(val.data(), uval, n_unp * sizeof(T));
    SZ3B_CUDA(stream_wait(ws.st));
    for (unsigned long long i = 0; i < n_unp; i++)
        if (pos[i] < limit) out.emplace_back(pos[i], val[i]);
}

// RegressionPredictor::save (RegressionPredictor.hpp:94-107): `nsel` selected blocks, their coefficient indices on the
// device (dense), the stored exact coefficients as (dense position, value).
template <class T>
static void regression_save(Workspace &ws, int N, unsigned long long nsel, std::vector<std::pair<unsigned long long, T>> &unp,
                            const int32_t *coef_q, double eb_indep, double eb_liner, std::vector<uint8_t> &pred_blob,
                            const std::function<void()> *after_hist = nullptr) {
    const int nc = N + 1;
    const int kCoefRadius = 32768;
    const uint64_t n_coef = nsel * nc;
    // back in chain order, split by quantizer
    std::sort(unp.begin(), unp.end(), [](const std::pair<unsigned long long, T> &a, const std::pair<unsigned long long, T> &b) {
        return a.first < b.first;
    });
    std::vector<T> un_liner, un_indep;
    for (const auto &e : unp) (e.first % nc == static_cast<unsigned>(N) ? un_indep : un_liner).push_back(e.second);
    pred_blob.clear();
    {
        uint8_t tmp[8];
        uint8_t *p = tmp;
        put<uint64_t>(p, n_coef);
        pred_blob.insert(pred_blob.end(), tmp, p);
    }
    if (n_coef) {
        quantizer_save<T>(pred_blob, eb_indep, kCoefRadius, un_indep);
        quantizer_save<T>(pred_blob, eb_liner, kCoefRadius, un_liner);
        std::vector<uint8_t> side;
        huffman_encode_device(ws, coef_q, n_coef, side, nullptr, after_hist);
        pred_blob.insert(pred_blob.end(), side.begin(), side.end());
    } else if (after_hist) {
        (*after_hist)();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// BlockwiseDecomposition with a Lorenzo predictor in the stack (lorenzo.cuh / lorenzo.cu).
//   Lorenzo only (1st, 2nd, or both composed): one exact wavefront pass.
//   Lorenzo + regression: selection guess pass, then  exact chain over the guess -> exact wavefront pass  until the
//   selection reproduces itself; after kBwMaxPasses invalidated guesses the row-major walk (k_bw_serial) finishes
//   from the first wrong block with the reference's sequential semantics.
// ---------------------------------------------------------------------------------------------------------------------
constexpr uint64_t kBwMinWindow = 4096;   // blocks per window of the selection iteration (grows with the advance)
constexpr uint64_t kBwWalkBelow = 24;     // an advance below this many blocks per pass hands the next stretch to the walk
constexpr uint64_t kBwWalkLen = 512;

template <class T, class QT>
static uint64_t bw_args_init(BwArgs<T, QT> &A, const sz3b_config &conf, const BlockShape &bs, double eb) {
    memset(&A, 0, sizeof(A));
    A.bs = bs;
    uint64_t np = 1;
    for (int d = bs.N - 1; d >= 0; d--) {
        A.pstride[d] = np;
        np *= static_cast<uint64_t>(bs.dims[d]) + kBwPad;
    }
    A.qp = make_quant(eb, conf.quantbinCnt / 2);
    A.noise[0] = lorenzo_noise<T>(bs.N, 1, eb);
    A.noise[1] = lorenzo_noise<T>(bs.N, 2, eb);
    if (conf.lorenzo) A.kinds[A.nk++] = PK_LORENZO1;
    if (conf.lorenzo2) A.kinds[A.nk++] = PK_LORENZO2;
    if (conf.regression) A.kinds[A.nk++] = PK_REG;
    const int nc = bs.N + 1;
    A.q_liner = make_quant(eb / nc / static_cast<unsigned>(conf.blockSize), 32768);
    A.q_indep = make_quant(eb / nc, 32768);
    A.b_lo = 0;
    A.b_hi = bs.nblocks;
    return np;
}

// diagonal table of a full block (BwArgs::diag_tab), uploaded per call (a few hundred entries)
template <class T, class QT>
static void bw_upload_diag_table(Workspace &ws, BwArgs<T, QT> &A) {
    const int N = A.bs.N;
    const uint32_t B = A.bs.B;
    uint64_t npts = 1;
    for (int d = 0; d < N; d++) npts *= B;
    if (npts > 0xffffu) return;
    const uint32_t nstart = static_cast<uint32_t>(N) * (B - 1) + 2;
    std::vector<uint32_t> tab(npts), idx(npts);
    std::vector<uint16_t> start(nstart + 1);
    if (!bw_build_diag_table(N, B, tab.data(), start.data(), idx.data())) return;
    uint8_t *d = ws.tables.as<uint8_t>(npts * 8 + nstart * 2 + 64);
    ws.h2d(d, tab.data(), npts * 4);
    ws.h2d(d + npts * 4, idx.data(), npts * 4);
    ws.h2d(d + npts * 8, start.data(), nstart * 2);
    SZ3B_CUDA(stream_wait(ws.st));   // tab / idx / start are locals
    A.diag_tab = reinterpret_cast<const uint32_t *>(d);
    A.diag_idx = reinterpret_cast<const uint32_t *>(d + npts * 4);
    A.diag_start = reinterpret_cast<const uint16_t *>(d + npts * 8);
}

// lorenzo_compress_test (SZAlgoInterp.hpp:78-119): every sampled block through BlockwiseDecomposition with the composed
// Lorenzo(1st) + Lorenzo(2nd) predictor (one decomposition object: the selections and the unpredictable values of all
// sampled blocks end up in one save()), the merged index stream through Huffman and zstd; returns the ratio.
template <class T, class QT>
static double lorenzo_trial_t(Workspace &ws, const sz3b_config &lc, const T *d_cubes, uint32_t ncubes, size_t per_block,
                              uint8_t *out, size_t cap) {
    sz3b_config c2 = lc;
    c2.lorenzo = 1;
    c2.lorenzo2 = 1;
    c2.regression = 0;
    BlockShape bs;
    block_shape_init(bs, c2.N, c2.dims, static_cast<uint32_t>(c2.blockSize));
    if (bs.num != per_block) fail(SZ3B_E_RUNTIME, "tuner: sampled block shape mismatch");
    BwArgs<T, QT> A;
    const uint64_t np = bw_args_init<T, QT>(A, c2, bs, c2.absErrorBound);
    bw_upload_diag_table<T, QT>(ws, A);
    const int radius = c2.quantbinCnt / 2, nbins = 2 * radius;
    const uint64_t n = static_cast<uint64_t>(per_block) * ncubes;
    QT *d_q = ws.cube_q.as<QT>(n);
    T *d_un = ws.cube_unpred.as<T>(n);
    unsigned long long *d_hist = ws.hist2.as<unsigned long long>(nbins);
    T *W = ws.padded.as<T>(np * ncubes);
    uint8_t *d_sel = ws.bsel.as<uint8_t>(bs.nblocks * ncubes);
    A.W = W;
    A.q = d_q;
    A.unpred_tmp = d_un;
    A.sel_out = d_sel;
    A.nbatch = ncubes;
    A.w_bstride = np;
    A.q_bstride = per_block;
    A.sel_bstride = bs.nblocks;
    A.mode = BW_EXACT;
    int launches = 0;
    size_t h = ws.stage_begin("predict_quantize");
    SZ3B_CUDA(cudaMemsetAsync(W, 0, np * ncubes * sizeof(T), ws.st));
    launch_bw_pad<T>(d_cubes, bs, A.pstride, W, 0, bs.nblocks, ws.st, ncubes, np);
    if (const char *e = launch_bw_fronts<T, QT>(A, ws.st, &launches)) fail(SZ3B_E_UNSUPPORTED, e);
    SZ3B_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nbins, ws.st));
    launch_histogram<QT>(d_q, n, 0, nbins, radius, d_hist, ws.st);
    ws.stage_end(h, launches + 2);
    SZ3B_CUDA(cudaGetLastError());
    std::vector<uint8_t> hdr;
    {
        const uint64_t nsel = bs.nblocks * ncubes;
        uint8_t tmp[8];
        uint8_t *p = tmp;
        put<uint64_t>(p, nsel);
        hdr.insert(hdr.end(), tmp, p);
        int32_t *d_sel32 = ws.side_q.as<int32_t>(nsel);
        launch_widen_u8(d_sel, nsel, d_sel32, ws.st);
        std::vector<uint8_t> side;
        huffman_encode_device(ws, d_sel32, nsel, side, nullptr);
        hdr.insert(hdr.end(), side.begin(), side.end());
    }
    HuffmanBook book;
    EncodeLayout lay;
    encode_indices<QT, T>(ws, d_q, n, d_hist, nbins, 0, true, d_un, book, lay);
    uint8_t qh[32];
    const size_t qh_len = quantizer_header<T>(c2.absErrorBound, radius, lay.n_unpred, qh);
    hdr.insert(hdr.end(), qh, qh + qh_len);
    uint8_t *buf = nullptr;
    ArrivalGate gate;
    const size_t len = assemble_stream<T>(ws, hdr.data(), hdr.size(), lay, book, n, ws.stage2, &buf, gate);
    const size_t size = zstd_stage(ws, buf, len, out, cap, 1, &gate);
    return static_cast<double>(per_block) * ncubes * sizeof(T) * 1.0 / size;
}

template <class T>
static double lorenzo_trial(Workspace &ws, const sz3b_config &lc, const T *d_cubes, uint32_t ncubes, size_t per_block,
                            uint8_t *out, size_t cap) {
    if (lc.quantbinCnt / 2 <= 32768) return lorenzo_trial_t<T, uint16_t>(ws, lc, d_cubes, ncubes, per_block, out, cap);
    return lorenzo_trial_t<T, uint32_t>(ws, lc, d_cubes, ncubes, per_block, out, cap);
}

// Lorenzo only (1st, 2nd, or both composed): one exact wavefront pass over all blocks.
// Lorenzo + regression: a selection guess pass, then windows of consecutive blocks [b_lo, b_hi): exact chain over the
// guessed selection of the window (continued from the last final block) -> exact wavefront pass over the window that
// re-derives the selection.  Everything before the first block whose guess was wrong is final; the next window
// starts there with the corrected guess.  The window follows the advance; when the advance collapses the row-major
// walk (k_bw_serial, the reference's sequential semantics) does the next stretch.
template <class T, class QT>
static void run_blockwise_lorenzo(Workspace &ws, const sz3b_config &conf, double eb, const T *d_data, QT *d_q,
                                  T *d_unpred_tmp, unsigned long long *d_hist, int nbins, std::vector<uint8_t> &pred_blob,
                                  int *launches) {
    const int N = conf.N, nc = N + 1;
    if (conf.blockSize < 1) fail(SZ3B_E_INVALID_ARGUMENT, "blockSize must be positive");
    BlockShape bs;
    block_shape_init(bs, N, conf.dims, static_cast<uint32_t>(conf.blockSize));
    if (bs.nblocks >= 0xfffffff0ull) fail(SZ3B_E_UNSUPPORTED, "more than 2^32 blocks");
    BwArgs<T, QT> A;
    const uint64_t np = bw_args_init<T, QT>(A, conf, bs, eb);
    bw_upload_diag_table<T, QT>(ws, A);
    const bool has_reg = conf.regression != 0;
    const int reg_sid = A.nk - 1;   // the regression predictor is always the last of the stack
    T *W = ws.padded.as<T>(np);
    A.W = W;
    A.q = d_q;
    A.unpred_tmp = d_unpred_tmp;
    uint8_t *selA = ws.bsel.as<uint8_t>(bs.nblocks), *selB = ws.bsel2.as<uint8_t>(bs.nblocks);
    unsigned *d_mm = ws.misc.as<unsigned>(4);
    auto pad = [&](uint64_t b_lo, uint64_t b_hi) {
        if (b_lo == 0 && b_hi == bs.nblocks) SZ3B_CUDA(cudaMemsetAsync(W, 0, np * sizeof(T), ws.st));
        launch_bw_pad<T>(d_data, bs, A.pstride, W, b_lo, b_hi, ws.st);
        *launches += 1;
    };
    auto fronts = [&](int mode, uint64_t b_lo, uint64_t b_hi) {
        A.mode = mode;
        A.b_lo = b_lo;
        A.b_hi = b_hi;
        if (const char *e = launch_bw_fronts<T, QT>(A, ws.st, launches)) fail(SZ3B_E_UNSUPPORTED, e);
    };
    size_t h = ws.stage_begin("predict_quantize");
    const int l0 = *launches;
    uint32_t nsel_lo = 0;   // regression-selected blocks among the final ones
    int32_t *coef_q = nullptr;
    std::vector<std::pair<unsigned long long, T>> unp;
    int n_pass = 0, n_walk = 0;
    if (!has_reg) {
        pad(0, bs.nblocks);
        A.sel_out = selA;
        fronts(BW_EXACT, 0, bs.nblocks);
        n_pass = 1;
    } else {
        T *c_fit = ws.coef.as<T>(bs.nblocks * nc);
        T *c_rec = ws.coef2.as<T>(bs.nblocks * nc);
        T *c_spec = ws.cspec.as<T>(bs.nblocks * nc);
        T *c_dense = ws.cdense.as<T>(bs.nblocks * nc);
        uint8_t *valid = ws.flags.as<uint8_t>(bs.nblocks);
        uint32_t *rank = ws.brank.as<uint32_t>(bs.nblocks + 1);
        coef_q = ws.coef_q.as<int32_t>(bs.nblocks * nc);
        unsigned long long *counters = ws.counters.as<unsigned long long>(4);
        unsigned long long *upos = ws.cpos.as<unsigned long long>(bs.nblocks * nc + 16);
        T *uval = ws.cval.as<T>(bs.nblocks * nc + 16);
        launch_reg_fit<T>(d_data, bs, c_fit, valid, ws.st);
        launch_bw_spec_coef<T>(c_fit, valid, bs.nblocks, N, A.q_liner, A.q_indep, c_spec, ws.st);
        *launches += 2;
        A.c_fit = c_fit;
        A.fit_valid = valid;
        A.c_spec = c_spec;
        A.c_rec = c_rec;
        A.rank = rank;
        A.mismatch = d_mm;
        A.coef_q = coef_q;
        A.c_rec_out = c_rec;
        A.n_unpred_coef = counters + 1;
        A.unpred_pos = upos;
        A.unpred_val = uval;
        pad(0, bs.nblocks);
        A.sel_out = selA;
        fronts(BW_SPEC, 0, bs.nblocks);
        uint64_t b_lo = 0, win = bs.nblocks, last_adv = bs.nblocks;
        while (b_lo < bs.nblocks) {
            const T *init = nsel_lo ? c_rec + static_cast<uint64_t>(nsel_lo - 1) * nc : nullptr;
            unsigned long long n_unp = 0, cnt = 0;
            if (last_adv < kBwWalkBelow) {
                const uint64_t b_hi = std::min<uint64_t>(bs.nblocks, b_lo + kBwWalkLen);
                pad(b_lo, b_hi);
                SZ3B_CUDA(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned long long), ws.st));
                A.mode = BW_SERIAL;
                A.sel_in = nullptr;
                A.sel_out = selA;
                if (const char *e = launch_bw_serial<T, QT>(A, b_lo, b_hi, init, nsel_lo, counters + 2, ws.st))
                    fail(SZ3B_E_UNSUPPORTED, e);
                *launches += 1;
                ws.d2h(&n_unp, counters + 1, sizeof(n_unp));
                ws.d2h(&cnt, counters + 2, sizeof(cnt));
                SZ3B_CUDA(stream_wait(ws.st));
                SZ3B_CUDA(cudaGetLastError());
                fetch_coef_unpred<T>(ws, n_unp, upos, uval, cnt * nc, unp);
                nsel_lo = static_cast<uint32_t>(cnt);
                b_lo = b_hi;
                last_adv = kBwWalkBelow;   // try a window again
                win = kBwMinWindow;
                n_walk++;
                continue;
            }
            const uint64_t b_hi = std::min<uint64_t>(bs.nblocks, b_lo + win);
            launch_bw_rank(selA, b_lo, b_hi, reg_sid, nsel_lo, rank, counters + 2, ws.st);
            launch_bw_gather_fit<T>(c_fit, selA, reg_sid, rank, b_lo, b_hi, nc, c_dense, ws.st);
            *launches += 2;
            ws.d2h(&cnt, counters + 2, sizeof(cnt));
            SZ3B_CUDA(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned long long), ws.st));
            SZ3B_CUDA(stream_wait(ws.st));
            const uint64_t at = static_cast<uint64_t>(nsel_lo) * nc;
            if (cnt) {
                launch_reg_chain<T>(c_dense + at, nullptr, cnt, N, A.q_liner, A.q_indep, coef_q + at, c_rec + at, counters, upos,
                                    uval, ws.st, init, at);
                *launches += 1;
            }
            pad(b_lo, b_hi);
            const unsigned mm_init[2] = {0u, ~0u};
            ws.h2d(d_mm, mm_init, sizeof(mm_init));
            A.sel_in = selA;
            A.sel_out = selB;
            fronts(BW_EXACT, b_lo, b_hi);
            unsigned mm[2];
            ws.d2h(mm, d_mm, sizeof(mm));
            ws.d2h(&n_unp, counters + 1, sizeof(n_unp));
            SZ3B_CUDA(stream_wait(ws.st));
            SZ3B_CUDA(cudaGetLastError());
            n_pass++;
            uint64_t boundary = b_hi;
            uint32_t nsel_after = nsel_lo + static_cast<uint32_t>(cnt);
            if (mm[0]) {
                boundary = mm[1];
                ws.d2h(&nsel_after, rank + boundary, sizeof(nsel_after));
                SZ3B_CUDA(cudaMemcpyAsync(selA + boundary, selB + boundary, b_hi - boundary, cudaMemcpyDeviceToDevice, ws.st));
                SZ3B_CUDA(stream_wait(ws.st));
            }
            fetch_coef_unpred<T>(ws, n_unp, upos, uval, static_cast<unsigned long long>(nsel_after) * nc, unp);
            last_adv = boundary - b_lo;
            b_lo = boundary;
            nsel_lo = nsel_after;
            win = std::max<uint64_t>(kBwMinWindow, 4 * last_adv);
        }
    }
    SZ3B_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nbins, ws.st));
    launch_histogram<QT>(d_q, bs.num, 0, nbins, conf.quantbinCnt / 2, d_hist, ws.st);
    *launches += 1;
    ws.stage_end(h, *launches - l0);
    SZ3B_CUDA(cudaGetLastError());
    if (getenv("SZ3B_VERBOSE")) fprintf(stderr, "[sz3b] lorenzo stack: %d exact passes, %d walks\n", n_pass, n_walk);
    // ComposedPredictor::save (ComposedPredictor.hpp:52-64): predictors in order (only regression stores anything),
    // then the per-block selection, Huffman coded
    pred_blob.clear();
    if (has_reg)
        regression_save<T>(ws, N, nsel_lo, unp, coef_q, eb / nc, eb / nc / static_cast<unsigned>(conf.blockSize), pred_blob);
    if (A.nk > 1) {
        uint8_t tmp[8];
        uint8_t *p = tmp;
        put<uint64_t>(p, bs.nblocks);
        pred_blob.insert(pred_blob.end(), tmp, p);
        int32_t *d_sel32 = ws.side_q.as<int32_t>(bs.nblocks);
        launch_widen_u8(selA, bs.nblocks, d_sel32, ws.st);
        std::vector<uint8_t> side;
        huffman_encode_device(ws, d_sel32, bs.nblocks, side, nullptr);
        pred_blob.insert(pred_blob.end(), side.begin(), side.end());
    }
}

template <class T, class QT>
static void run_blockwise(Workspace &ws, const sz3b_config &conf, double eb, const T *d_data, QT *d_q, T *d_unpred_tmp,
                          unsigned long long *d_hist, int nbins, std::vector<uint8_t> &pred_blob, int *launches) {
    const int N = conf.N;
    const int method_cnt = (conf.lorenzo != 0) + (conf.lorenzo2 != 0) + (conf.regression != 0);
    if (method_cnt == 0) fail(SZ3B_E_INVALID_ARGUMENT, "All lorenzo and regression methods are disabled.");
    if (conf.lorenzo || conf.lorenzo2) {
        run_blockwise_lorenzo<T, QT>(ws, conf, eb, d_data, d_q, d_unpred_tmp, d_hist, nbins, pred_blob, launches);
        return;
    }
    if (conf.blockSize < 1) fail(SZ3B_E_INVALID_ARGUMENT, "blockSize must be positive");
    BlockShape bs;
    block_shape_init(bs, N, conf.dims, static_cast<uint32_t>(conf.blockSize));
    for (int d = 0; d < N; d++)
        if (bs.dims[d] % bs.B == 1)
            fail(SZ3B_E_UNSUPPORTED,
                 "regression-only with a block of extent 1: the reference falls back to an unpadded Lorenzo predictor "
                 "that reads outside the array; not reproduced");
    const int nc = N + 1;
    const int kCoefRadius = 32768;   // LinearQuantizer default (LinearQuantizer.hpp:21)
    const double eb_indep = eb / nc;
    const double eb_liner = eb / nc / static_cast<unsigned>(conf.blockSize);
    T *c_fit = ws.coef.as<T>(bs.nblocks * nc);
    T *c_rec = ws.coef2.as<T>(bs.nblocks * nc);
    uint8_t *valid = ws.flags.as<uint8_t>(bs.nblocks);
    int32_t *coef_q = ws.coef_q.as<int32_t>(bs.nblocks * nc);
    unsigned long long *counters = ws.counters.as<unsigned long long>(2);
    unsigned long long *upos = ws.cpos.as<unsigned long long>(bs.nblocks * nc);
    T *uval = ws.cval.as<T>(bs.nblocks * nc);
    size_t h = ws.stage_begin("regression_fit");
    launch_reg_fit<T>(d_data, bs, c_fit, valid, ws.st);
    ws.stage_end(h, 1);
    h = ws.stage_begin("regression_coef_chain");
    SZ3B_CUDA(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned long long), ws.st));
    // regression-only with no clipped-to-1 block: every block is selected -> the speculative dense chain (sel = null)
    launch_reg_chain<T>(c_fit, nullptr, bs.nblocks, N, make_quant(eb_liner, kCoefRadius), make_quant(eb_indep, kCoefRadius),
                        coef_q, c_rec, counters, upos, uval, ws.st);
    ws.stage_end(h, 1);
    unsigned long long hc[2];
    ws.d2h(hc, counters, sizeof(hc));
    SZ3B_CUDA(stream_wait(ws.st));
    SZ3B_CUDA(cudaGetLastError());
    *launches += 2;
    // The data prediction needs the reconstructed coefficients only, not their encoded side stream: it is launched (on
    // a second stream) once the side stream's histogram kernels are queued, and runs while the host builds that
    // stream's tree (2 ms for C3's coefficient indices).
    cudaStream_t st_main = ws.st, st_pred = ws.low_priority_stream();
    cudaEvent_t ev_chain = ws.event(), ev_pred = ws.event();
    SZ3B_CUDA(cudaEventRecord(ev_chain, st_main));
    const char *perr = nullptr;
    const std::function<void()> launch_predict = [&]() {
        SZ3B_CUDA(cudaStreamWaitEvent(st_pred, ev_chain, 0));
        ws.st = st_pred;   // (stage events and the launch go to the second stream)
        try {
            const size_t hp = ws.stage_begin("predict_quantize");
            SZ3B_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nbins, ws.st));
            perr = launch_reg_predict<T, QT>(d_data, bs, c_rec, make_quant(eb, conf.quantbinCnt / 2), d_q, d_unpred_tmp, d_hist, ws.st);
            ws.stage_end(hp, 1);
            SZ3B_CUDA(cudaEventRecord(ev_pred, st_pred));
        } catch (...) {
            ws.st = st_main;
            throw;
        }
        ws.st = st_main;
    };
    *launches += 1;
    std::vector<std::pair<unsigned long long, T>> unp;
    double t_side = now_ms();
    fetch_coef_unpred<T>(ws, hc[1], upos, uval, hc[0] * nc, unp);
    regression_save<T>(ws, N, hc[0], unp, coef_q, eb_indep, eb_liner, pred_blob, &launch_predict);
    ws.host_stage("regression_side_stream_wall", now_ms() - t_side);
    if (perr) fail(SZ3B_E_UNSUPPORTED, perr);
    SZ3B_CUDA(cudaStreamWaitEvent(st_main, ev_pred, 0));
    SZ3B_CUDA(cudaGetLastError());
}

template <class T>
static size_t quantizer_header(double eb, int radius, uint64_t n_unpred, uint8_t *out) {
    uint8_t *p = out;
    put<uint8_t>(p, 2);
    put<double>(p, eb);
    put<int32_t>(p, radius);
    put<uint64_t>(p, n_unpred);
    return static_cast<size_t>(p - out);
}

template <class T, class QT>
static size_t blockwise_compress_t(Workspace &ws, const sz3b_config &conf, const T *d_data, uint8_t *dst, size_t cap,
                                   int zstd_threads) {
    const int radius = conf.quantbinCnt / 2;
    const int nbins = 2 * radius;
    const uint64_t n = config_num(conf);
    QT *d_q = ws.q.as<QT>(n);
    T *d_unpred_tmp = ws.unpred_tmp.as<T>(n);
    unsigned long long *d_hist = ws.hist2.as<unsigned long long>(nbins);
    std::vector<uint8_t> hdr;
    int launches = 0;
    run_blockwise<T, QT>(ws, conf, conf.absErrorBound, d_data, d_q, d_unpred_tmp, d_hist, nbins, hdr, &launches);
    HuffmanBook book;
    EncodeLayout lay;
    encode_indices<QT, T>(ws, d_q, n, d_hist, nbins, 0, true, d_unpred_tmp, book, lay);
    uint8_t qh[32];
    size_t qh_len = quantizer_header<T>(conf.absErrorBound, radius, lay.n_unpred, qh);
    hdr.insert(hdr.end(), qh, qh + qh_len);
    if (lossless_policy() == 2 && stream_len<T>(hdr.size(), lay) >= kZhufMinStream)
        return zhuf_stage<T>(ws, hdr.data(), hdr.size(), lay, book, n, dst, cap);
    uint8_t *buf = nullptr;
    ArrivalGate gate;
    size_t len = assemble_stream<T>(ws, hdr.data(), hdr.size(), lay, book, n, ws.stage, &buf, gate);
    return zstd_stage(ws, buf, len, dst, cap, zstd_threads, &gate);
}

template <class T>
static size_t blockwise_compress(Workspace &ws, const sz3b_config &conf, const T *d_data, uint8_t *dst, size_t cap,
                                 int zstd_threads) {
    if (conf.quantbinCnt < 2) fail(SZ3B_E_INVALID_ARGUMENT, "quantbinCnt must be >= 2");
    if (conf.quantbinCnt / 2 <= 32768) return blockwise_compress_t<T, uint16_t>(ws, conf, d_data, dst, cap, zstd_threads);
    return blockwise_compress_t<T, uint32_t>(ws, conf, d_data, dst, cap, zstd_threads);
}

template <class T, class QT>
static void blockwise_decompose_t(Workspace &ws, const sz3b_config &conf, double eb, const T *d_data,
                                  int32_t *quant_out, std::vector<uint8_t> &blob) {
    const int radius = conf.quantbinCnt / 2;
    const int nbins = 2 * radius;
    const uint64_t n = config_num(conf);
    QT *d_q = ws.q.as<QT>(n);
    T *d_unpred_tmp = ws.unpred_tmp.as<T>(n);
    unsigned long long *d_hist = ws.hist2.as<unsigned long long>(nbins);
    int launches = 0;
    run_blockwise<T, QT>(ws, conf, eb, d_data, d_q, d_unpred_tmp, d_hist, nbins, blob, &launches);
    int32_t *d_wide = ws.side_q.as<int32_t>(n);
    launch_widen<QT>(d_q, n, d_wide, ws.st);
    ws.d2h(quant_out, d_wide, n * sizeof(int32_t));
    HuffmanBook book;
    EncodeLayout lay;
    encode_indices<QT, T>(ws, d_q, n, d_hist, nbins, 0, true, d_unpred_tmp, book, lay);
    uint8_t qh[32];
    size_t qh_len = quantizer_header<T>(eb, radius, lay.n_unpred, qh);
    blob.insert(blob.end(), qh, qh + qh_len);
    const size_t at = blob.size();
    blob.resize(at + lay.n_unpred * sizeof(T));
    if (lay.n_unpred) ws.d2h(blob.data() + at, ws.unpred_out.p, lay.n_unpred * sizeof(T));
    SZ3B_CUDA(stream_wait(ws.st));
}

template <class T>
void blockwise_decompose_stage(Workspace &ws, const sz3b_config &conf, double eb, const T *data, int loc,
                               int32_t *quant_out, std::vector<uint8_t> &blob) {
    const T *d = to_device(ws, data, loc, config_num(conf));
    if (conf.quantbinCnt / 2 <= 32768)
        blockwise_decompose_t<T, uint16_t>(ws, conf, eb, d, quant_out, blob);
    else
        blockwise_decompose_t<T, uint32_t>(ws, conf, eb, d, quant_out, blob);
}

// ---------------------------------------------------------------------------------------------------------------------
// lossless path (SZDispatcher.hpp:54-59): size_t | zstd(raw bytes)
// ---------------------------------------------------------------------------------------------------------------------
template <class T>
static size_t lossless_compress(Workspace &ws, const sz3b_config &conf, const T *data, int loc, uint8_t *dst,
                                size_t cap, bool strict_cap = false) {
    const size_t bytes = config_num(conf) * sizeof(T);
    // (strict_cap: the caller offered less than the reference's capacity and retries with all of it)
    if (strict_cap && cap < ZSTD_compressBound(bytes) + sizeof(uint64_t)) throw TooSmall{};
    const uint8_t *src = reinterpret_cast<const uint8_t *>(data);
    if (loc == SZ3B_DEVICE) {
        uint8_t *h = static_cast<uint8_t *>(ws.stage.ensure(bytes));
        ws.d2h(h, data, bytes);
        SZ3B_CUDA(stream_wait(ws.st));
        src = h;
    }
    try {
        return zstd_stage(ws, src, bytes, dst, cap, host_threads(), nullptr);
    } catch (TooSmall &) {
        fail(SZ3B_E_RUNTIME, "compressed buffer not large enough");   // std::length_error escapes in the reference
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// SZ_compress_dispatcher (SZDispatcher.hpp:13-76).  `range` > 0: precomputed value range (OMP slabs).
// ---------------------------------------------------------------------------------------------------------------------
template <class T>
static size_t dispatch_compress_inner(Workspace &ws, sz3b_config &conf, const T *data, int loc, uint8_t *dst, size_t cap,
                                      T range, bool have_range, bool strict_cap);

// With a placer the payload ends up where placer->place(size) says: straight from the device when the GPU lossless
// stage wrote it (zhuf_run), by a copy from `dst` otherwise.
template <class T>
static size_t dispatch_compress(Workspace &ws, sz3b_config &conf, const T *data, int loc, uint8_t *dst, size_t cap,
                                T range, bool have_range = false, bool strict_cap = false, Placer *placer = nullptr) {
    if (!placer) return dispatch_compress_inner<T>(ws, conf, data, loc, dst, cap, range, have_range, strict_cap);
    placer->raw_bytes = config_num(conf) * sizeof(T);
    ws.placer = placer;
    size_t size = 0;
    try {
        size = dispatch_compress_inner<T>(ws, conf, data, loc, dst, cap, range, have_range, strict_cap);
    } catch (...) {
        ws.placer = nullptr;
        throw;
    }
    ws.placer = nullptr;
    if (!placer->used) {
        uint8_t *f = placer->place(size);
        placer->used = true;
        memcpy(f, dst, size);
    }
    return size;
}

template <class T>
static size_t dispatch_compress_inner(Workspace &ws, sz3b_config &conf, const T *data, int loc, uint8_t *dst, size_t cap,
                                      T range, bool have_range, bool strict_cap) {
    const uint64_t num = config_num(conf);
    const T *d_data = nullptr;
    auto dev = [&]() {
        if (!d_data) d_data = to_device(ws, data, loc, num);
        return d_data;
    };
    if (conf.errorBoundMode != SZ3B_EB_ABS) resolve_abs_eb<T>(ws, conf, have_range ? nullptr : dev(), range, have_range);
    size_t cmp_size = 0;
    if (conf.absErrorBound == 0) conf.cmprAlgo = SZ3B_ALGO_LOSSLESS;
    bool cap_ok = true;
    if (conf.cmprAlgo != SZ3B_ALGO_LOSSLESS) {
        try {
            if (conf.cmprAlgo == SZ3B_ALGO_INTERP_LORENZO) {
                const T *alias = (loc == SZ3B_HOST && !d_data) ? static_cast<const T *>(mapped_alias(data)) : nullptr;
                if (alias) {
                    // pinned host input: the tuner samples ~0.5 % of the array straight from host memory while the
                    // bulk copy runs on its own stream
                    double t0 = now_ms();
                    // 3-D arrays with an absolute bound go up in plane order, and predict+quantize follows the copy
                    // (run_interp); everything else waits for the whole array after tuning
                    const bool planes = conf.N == 3 && conf.dims[0] >= 64 && num < (1ull << 32) &&
                                        num / conf.dims[0] * sizeof(T) >= (64u << 10);
                    d_data = planes ? to_device_planes<T>(ws, data, conf) : to_device_background<T>(ws, data, num);
                    tune_interp<T>(ws, conf, alias);
                    if (!(planes && conf.cmprAlgo == SZ3B_ALGO_INTERP)) join_copy(ws);
                    ws.host_stage("tune_overlapped_with_h2d", now_ms() - t0);
                } else {
                    const double t0 = now_ms();
                    tune_interp<T>(ws, conf, dev());   // rewrites cmprAlgo to ALGO_INTERP (N >= 2)
                    ws.host_stage("tune_wall", now_ms() - t0);
                }
            }
            if (conf.cmprAlgo == SZ3B_ALGO_INTERP) {
                set_default_anchor(conf);
                cmp_size = interp_compress<T>(ws, conf, dev(), 1, dst, cap, host_threads(), false);
            } else if (conf.cmprAlgo == SZ3B_ALGO_LORENZO_REG) {
                cmp_size = blockwise_compress<T>(ws, conf, dev(), dst, cap, host_threads());
            } else if (conf.cmprAlgo == SZ3B_ALGO_NOPRED || conf.cmprAlgo == 5 || conf.cmprAlgo == 6) {
                fail(SZ3B_E_UNSUPPORTED, "ALGO_NOPRED / ALGO_BIOMD* are outside the GPU hot path (DESIGN.md, scope)");
            } else {
                fail(SZ3B_E_INVALID_ARGUMENT, "Unknown compression algorithm");
            }
        } catch (TooSmall &) {
            if (strict_cap) throw;   // the caller offered less than the reference's capacity and retries with all of it
            cap_ok = false;
        }
    }
    if (conf.cmprAlgo == SZ3B_ALGO_LOSSLESS || !cap_ok) {
        conf.cmprAlgo = SZ3B_ALGO_LOSSLESS;
        return lossless_compress<T>(ws, conf, data, loc, dst, cap, strict_cap);
    }
    if (num * sizeof(T) / 1.0 / cmp_size < 3) {
        size_t zcap = ZSTD_compressBound(num * sizeof(T)) + sizeof(uint64_t);
        std::vector<uint8_t> tmp(zcap);
        size_t zsize = lossless_compress<T>(ws, conf, data, loc, tmp.data(), zcap);
        if (zsize < cmp_size && zsize <= cap) {
            conf.cmprAlgo = SZ3B_ALGO_LOSSLESS;
            memcpy(dst, tmp.data(), zsize);
            cmp_size = zsize;
        }
    }
    return cmp_size;
}

template <class T>
size_t compress_slab(Workspace &ws, sz3b_config &slab_conf, const T *slab, int loc, double range, uint8_t *payload,
                     size_t cap) {
    return dispatch_compress<T>(ws, slab_conf, slab, loc, payload, cap, static_cast<T>(range), range >= 0);
}

// One slab of a container.  A slab is first offered a third of its size (in a pinned buffer the workspace keeps, or in
// `own` when that one still holds an earlier slab of the same device); only a slab that does not fit -- nearly
// incompressible data -- is run again with the capacity the reference gives it.  With a placer the payload goes where
// placer->place(size) says (see Placer); *out is where it was staged otherwise.
template <class T>
static size_t compress_slab_buffered(Workspace &w, sz3b_config &sconf, const T *p, int sloc, T range, bool have_range,
                                     Placer *placer, std::vector<uint8_t> *own, uint8_t **out) {
    const uint64_t bytes = config_num(sconf) * sizeof(T);
    // (the reference gives a slab ZSTD_compressBound(bytes), which its own lossless path -- bound + the 8-byte length
    //  prefix, Lossless_zstd.hpp:29-33 -- cannot use: a lossless slab ends SZ_compress_OMP with an uncaught
    //  std::length_error.  Eight bytes more and the slab is stored.)
    const size_t full = ZSTD_compressBound(bytes) + sizeof(uint64_t);
    const size_t small = std::min<size_t>(full, std::max<size_t>(bytes / 3, static_cast<size_t>(1) << 20));
    uint8_t *buf;
    if (own) {
        own->resize(small);
        buf = own->data();
    } else {
        buf = static_cast<uint8_t *>(w.slab_out.ensure(small));
    }
    const sz3b_config before = sconf;
    size_t size;
    try {
        size = dispatch_compress<T>(w, sconf, p, sloc, buf, small, range, have_range, small < full, placer);
    } catch (TooSmall &) {
        sconf = before;
        w.slab_big.resize(full);
        buf = w.slab_big.data();
        size = dispatch_compress<T>(w, sconf, p, sloc, buf, full, range, have_range, false, placer);
        if (own) {   // keep it alive beyond the next slab of this device
            own->assign(buf, buf + size);
            buf = own->data();
        }
    }
    *out = buf;
    return size;
}

// sz3b_compress_slab_placed: one rank's slab, the payload delivered where the caller's callback says once its size is
// known (the callback is where a rank exchanges sizes with the others and derives its offset in the shared container).
template <class T>
size_t compress_slab_placed(Workspace &ws, sz3b_config &slab_conf, const T *slab, int loc, double range,
                            void *(*place)(void *user, size_t size), void *user) {
    struct FnPlacer : Placer {
        void *(*fn)(void *, size_t);
        void *user;
        uint8_t *place(size_t size) override {
            uint8_t *p = static_cast<uint8_t *>(fn(user, size));
            if (!p) fail(SZ3B_E_INVALID_ARGUMENT, "placement callback returned no destination");
            return p;
        }
    } pl;
    pl.fn = place;
    pl.user = user;
    uint8_t *out = nullptr;
    return compress_slab_buffered<T>(ws, slab_conf, slab, loc, static_cast<T>(range), range >= 0, &pl, nullptr, &out);
}
template size_t compress_slab_placed<float>(Workspace &, sz3b_config &, const float *, int, double, void *(*)(void *, size_t), void *);
template size_t compress_slab_placed<double>(Workspace &, sz3b_config &, const double *, int, double, void *(*)(void *, size_t), void *);

// ---------------------------------------------------------------------------------------------------------------------
// SZ_compress_OMP (SZImplOMP.hpp:16-117): the array is cut into conf.openmp slabs along its outermost dimension, every
// slab is an independent stream, the container is what the reference's SZ_decompress_OMP expects.
//
// The reference gives every slab to one OpenMP thread; here every visible GPU takes slabs (slab t on device t mod G),
// each driven by its own host thread with its own workspace and streams, all inside this one call: a drop-in caller
// that sets conf.openmp gets the whole box without launching ranks.  Nothing is exchanged between the devices but the
// slabs' min / max (non-ABS bounds, :57-68) and their byte counts (:93-105), both through host memory; the payloads are
// written into the caller's buffer at their final offsets by the host worker pool.
// ---------------------------------------------------------------------------------------------------------------------
static std::atomic<int> g_device_fanout{-1};   // -1: default (all visible devices; 1 under a multi-process launcher)
void set_device_fanout(int n) { g_device_fanout.store(n < 0 ? -1 : n); }
int device_fanout() {
    int n = g_device_fanout.load();
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess) {
        cudaGetLastError();
        visible = 1;
    }
    if (n < 0) {
        // one process per GPU (torchrun and friends export LOCAL_WORLD_SIZE): every rank keeps to its own device
        const char *lws = getenv("LOCAL_WORLD_SIZE");
        n = (lws && atoi(lws) > 1) ? 1 : visible;
    }
    if (n == 0 || n > visible) n = visible;
    return std::max(1, n);
}

struct CopyPartsJob {
    uint8_t *dst;
    const std::vector<const uint8_t *> *src;
    const std::vector<size_t> *size, *start;
};
static void copy_parts_fn(void *arg, int worker) {
    CopyPartsJob &j = *static_cast<CopyPartsJob *>(arg);
    memcpy(j.dst + (*j.start)[worker], (*j.src)[worker], (*j.size)[worker]);
}

template <class T>
static size_t omp_compress(Workspace &ws, sz3b_config &conf, const T *data, int loc, uint8_t *dst, size_t cap) {
    int nslabs = conf.openmp;
    if (static_cast<uint64_t>(nslabs) > conf.dims[0]) nslabs = static_cast<int>(conf.dims[0]);
    const uint64_t num = config_num(conf);
    const uint64_t row = num / conf.dims[0];
    const int ndev = std::min(device_fanout(), nslabs);
    // the devices of this call: the current one first
    int visible = 1;
    cudaGetDeviceCount(&visible);
    std::vector<int> devs(ndev);
    for (int g = 0; g < ndev; g++) devs[g] = (ws.device + g) % std::max(1, visible);
    int data_dev = ws.device;
    if (loc == SZ3B_DEVICE) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, data) == cudaSuccess && at.type == cudaMemoryTypeDevice) data_dev = at.device;
        cudaGetLastError();
    }
    std::vector<int> lo(nslabs), hi(nslabs);
    std::vector<sz3b_config> confs(nslabs, conf);
    for (int t = 0; t < nslabs; t++) {
        lo[t] = static_cast<int>(static_cast<uint64_t>(t) * conf.dims[0] / nslabs);
        hi[t] = static_cast<int>(static_cast<uint64_t>(t + 1) * conf.dims[0] / nslabs);
    }
    // other devices read a device-resident input on their own streams: what the caller queued must be complete
    if (ndev > 1 && loc == SZ3B_DEVICE) SZ3B_CUDA(stream_wait(ws.st));
    std::vector<std::exception_ptr> errs(ndev);
    // runs fn(g, workspace of device g) on one host thread per device (the calling thread drives the first one)
    auto on_devices = [&](auto &&fn) {
        auto body = [&](int g) {
            try {
                if (g == 0) {
                    fn(g, ws);
                } else {
                    SZ3B_CUDA(cudaSetDevice(devs[g]));
                    WorkspaceLease w;
                    fn(g, *w);
                }
            } catch (...) {
                errs[g] = std::current_exception();
            }
        };
        std::vector<std::thread> th;
        for (int g = 1; g < ndev; g++) th.emplace_back(body, g);
        body(0);
        for (auto &t : th) t.join();
        for (int g = 0; g < ndev; g++)
            if (errs[g]) std::rethrow_exception(errs[g]);
    };
    // slab t as device g sees it: the caller's host pointer, the device pointer itself, or a peer copy of it
    auto slab_on = [&](int g, Workspace &w, int t, int *sloc) -> const T * {
        const T *p = data + static_cast<uint64_t>(lo[t]) * row;
        *sloc = loc;
        if (loc == SZ3B_DEVICE && devs[g] != data_dev) {
            const uint64_t n_t = static_cast<uint64_t>(hi[t] - lo[t]) * row;
            T *d = w.data.as<T>(n_t);
            SZ3B_CUDA(cudaMemcpyPeerAsync(d, devs[g], p, data_dev, n_t * sizeof(T), w.st));
            return d;
        }
        return p;
    };
    // ---- the shared bound (SZImplOMP.hpp:57-68): min / max per slab, reduced, then calAbsErrorBound ------------------
    const bool need_range = conf.errorBoundMode == SZ3B_EB_REL || conf.errorBoundMode == SZ3B_EB_PSNR ||
                            conf.errorBoundMode == SZ3B_EB_ABS_AND_REL || conf.errorBoundMode == SZ3B_EB_ABS_OR_REL;
    if (need_range) {
        std::vector<double> mn(nslabs), mx(nslabs);
        on_devices([&](int g, Workspace &w) {
            for (int t = g; t < nslabs; t += ndev) {
                int sloc;
                const T *p = slab_on(g, w, t, &sloc);
                minmax_stage<T>(w, p, sloc, static_cast<uint64_t>(hi[t] - lo[t]) * row, &mn[t], &mx[t]);
            }
        });
        const T range = static_cast<T>(static_cast<T>(*std::max_element(mx.begin(), mx.end())) -
                                       static_cast<T>(*std::min_element(mn.begin(), mn.end())));
        resolve_abs_eb<T>(ws, conf, nullptr, range, true);
    } else if (conf.errorBoundMode != SZ3B_EB_ABS) {
        resolve_abs_eb<T>(ws, conf, nullptr, static_cast<T>(0), true);   // L2NORM: no scan
    }
    for (int t = 0; t < nslabs; t++) {
        confs[t] = conf;
        uint64_t d[4];
        for (int i = 0; i < conf.N; i++) d[i] = conf.dims[i];
        d[0] = hi[t] - lo[t];
        config_set_dims(confs[t], conf.N, d);
    }
    // ---- the slabs ---------------------------------------------------------------------------------------------------
    // header of the container: its size does not depend on what the slabs turn out to be (Config blobs have fixed-width
    // fields; the slabs' error-bound mode is ABS by now)
    uint8_t blob[256];
    size_t header = 4 + static_cast<size_t>(nslabs) * 8;
    for (int t = 0; t < nslabs; t++) header += config_save(confs[t], blob);
    if (header > cap) fail(SZ3B_E_RUNTIME, "compressed buffer not large enough for the OpenMP container");
    std::vector<size_t> sizes(nslabs), start(nslabs + 1);
    // One slab per device: the slabs meet once their sizes are known (frames still on the devices), and every device
    // then delivers its frames straight to dst + header + offset.  More slabs than devices: staged and copied.
    const bool direct = nslabs == ndev;
    struct Meet {
        std::mutex mu;
        std::condition_variable cv;
        int arrived = 0, n = 0;
        bool failed = false;
        std::vector<size_t> *sizes, *start;
        uint8_t *base;
        size_t room;
    } meet;
    meet.n = nslabs;
    meet.sizes = &sizes;
    meet.start = &start;
    meet.base = dst + header;
    meet.room = cap - header;
    struct SlabPlacer : Placer {
        Meet *m;
        int t;
        uint8_t *place(size_t size) override {
            std::unique_lock<std::mutex> lk(m->mu);
            (*m->sizes)[t] = size;
            if (++m->arrived == m->n) {
                (*m->start)[0] = 0;
                for (int i = 0; i < m->n; i++) (*m->start)[i + 1] = (*m->start)[i] + (*m->sizes)[i];
                if ((*m->start)[m->n] > m->room) m->failed = true;
                m->cv.notify_all();
            } else {
                m->cv.wait(lk, [&] { return m->arrived >= m->n || m->failed; });
            }
            if (m->failed) fail(SZ3B_E_RUNTIME, "compressed buffer not large enough for the OpenMP container");
            return m->base + (*m->start)[t];
        }
    };
    std::vector<const uint8_t *> part(nslabs, nullptr);
    std::vector<std::vector<uint8_t>> own(nslabs);
    on_devices([&](int g, Workspace &w) {
        try {
            int k = 0;
            for (int t = g; t < nslabs; t += ndev, k++) {
                int sloc;
                const T *p = slab_on(g, w, t, &sloc);
                SlabPlacer sp;
                sp.m = &meet;
                sp.t = t;
                uint8_t *out = nullptr;
                sizes[t] = compress_slab_buffered<T>(w, confs[t], p, sloc, static_cast<T>(0), true, direct ? &sp : nullptr,
                                                     k == 0 ? nullptr : &own[t], &out);
                part[t] = out;
            }
        } catch (...) {
            if (direct) {   // release the devices waiting for this slab's size
                std::lock_guard<std::mutex> lk(meet.mu);
                meet.failed = true;
                meet.cv.notify_all();
            }
            throw;
        }
    });
    // ---- the container (:93-107) ---------------------------------------------------------------------------------------
    uint8_t *p = dst;
    put<int32_t>(p, nslabs);
    for (int t = 0; t < nslabs; t++) p += config_save(confs[t], p);
    for (int t = 0; t < nslabs; t++) put<uint64_t>(p, sizes[t]);
    if (!direct) {
        start[0] = 0;
        for (int t = 0; t < nslabs; t++) start[t + 1] = start[t] + sizes[t];
        if (header + start[nslabs] > cap) fail(SZ3B_E_RUNTIME, "compressed buffer not large enough for the OpenMP container");
        CopyPartsJob job{p, &part, &sizes, &start};
        if (nslabs == 1)
            copy_parts_fn(&job, 0);
        else
            host_parallel(nslabs, copy_parts_fn, &job);
    }
    return static_cast<size_t>(p - dst) + start[nslabs];
}

template <class T>
size_t compress_any(Workspace &ws, sz3b_config &conf, const T *data, int loc, uint8_t *cmp, size_t cap) {
    if (conf.openmp) return omp_compress<T>(ws, conf, data, loc, cmp, cap);
    return dispatch_compress<T>(ws, conf, data, loc, cmp, cap, static_cast<T>(0));
}

// ---------------------------------------------------------------------------------------------------------------------
// SZ_decompress_dispatcher / SZGenericCompressor::decompress / SZ_decompress_OMP
// (api/impl/SZDispatcher.hpp:79-107, compressor/SZGenericCompressor.hpp:65-84, api/impl/SZImplOMP.hpp:120-186).
// zstd and the (bit-serial, restart-free) Huffman decode run on the host; index expansion, unpredictable-value
// placement and every `recover` loop run on the GPU.
// ---------------------------------------------------------------------------------------------------------------------
struct Cursor {
    const uint8_t *p;
    size_t rem;
    template <class V>
    V get() {
        if (rem < sizeof(V)) fail(SZ3B_E_INVALID_ARGUMENT, "truncated stream");
        V v;
        memcpy(&v, p, sizeof(V));
        p += sizeof(V);
        rem -= sizeof(V);
        return v;
    }
    const uint8_t *take(size_t n) {
        if (rem < n) fail(SZ3B_E_INVALID_ARGUMENT, "truncated stream");
        const uint8_t *q = p;
        p += n;
        rem -= n;
        return q;
    }
};

// LinearQuantizer::load (LinearQuantizer.hpp:106-120)
template <class T>
static void quantizer_load(Cursor &c, double *eb, int *radius, const T **unpred, uint64_t *n_unpred) {
    (void)c.get<uint8_t>();   // uid
    *eb = c.get<double>();
    *radius = c.get<int32_t>();
    *n_unpred = c.get<uint64_t>();
    if (*n_unpred > c.rem / sizeof(T)) fail(SZ3B_E_INVALID_ARGUMENT, "truncated stream (unpredictable values)");
    *unpred = reinterpret_cast<const T *>(c.take(*n_unpred * sizeof(T)));
}

// encoder.load | size_t n | size_t outSize | bits  ->  indices as QT on the device.
// The tree is parsed on the host (a few thousand nodes); the bitstream is decoded on the GPU by the self-synchronising
// decoder of huffman_decode.cu.  The host table decoder remains only for the degenerate one-symbol tree.
template <class QT>
static QT *decode_indices(Workspace &ws, Cursor &c, uint64_t expect_n, bool has_count = true) {
    HuffmanDecoder dec;
    const char *err = nullptr;
    {
        const double t0 = now_ms();
        if (!dec.load(c.p, c.rem, &err)) fail(SZ3B_E_INVALID_ARGUMENT, err);
        ws.host_stage("huffman_tables_host", now_ms() - t0);
    }
    // (the main stream carries its symbol count between the tree and the bits, SZGenericCompressor.hpp:53; a
    //  predictor's side stream does not -- RegressionPredictor::save wrote it up front)
    const uint64_t n = has_count ? c.get<uint64_t>() : expect_n;
    if (n != expect_n) fail(SZ3B_E_INVALID_ARGUMENT, "index count does not match the array size");
    QT *d_q = ws.q.as<QT>(n);
    if (dec.leaf[0]) {   // every index identical: no bits in the stream (HuffmanEncoder.hpp:233-237)
        QT *h_q = static_cast<QT *>(ws.stage2.ensure(n * sizeof(QT)));
        if (!dec.decode<QT>(c.p, c.rem, n, h_q, &err)) fail(SZ3B_E_INVALID_ARGUMENT, err);
        ws.h2d(d_q, h_q, n * sizeof(QT));
        return d_q;
    }
    const uint64_t enc_len = c.get<uint64_t>();
    const uint8_t *bits = c.take(enc_len);
    const uint64_t total_bits = enc_len * 8;
    size_t h = ws.stage_begin("huffman_decode_upload");
    // tables: lut | L | R | C | leaf
    const uint32_t nc = dec.nc;
    const size_t lut_n = dec.dlut.size(), lut2_n = dec.lut2.size();
    const size_t tab_bytes = (lut_n + lut2_n + 3 * static_cast<size_t>(nc)) * 4 + nc + 64;
    uint8_t *d_tab = ws.hd_tab.as<uint8_t>(tab_bytes);
    uint32_t *d_lut = reinterpret_cast<uint32_t *>(d_tab);
    uint32_t *d_lut2 = d_lut + lut_n;
    uint32_t *d_L = d_lut2 + lut2_n, *d_R = d_L + nc;
    int *d_C = reinterpret_cast<int *>(d_R + nc);
    uint8_t *d_leaf = reinterpret_cast<uint8_t *>(d_C + nc);
    ws.h2d(d_lut, dec.dlut.data(), lut_n * 4);
    ws.h2d(d_lut2, dec.lut2.data(), lut2_n * 4);
    ws.h2d(d_L, dec.L.data(), nc * 4);
    ws.h2d(d_R, dec.R.data(), nc * 4);
    ws.h2d(d_C, dec.C.data(), nc * 4);
    ws.h2d(d_leaf, dec.leaf.data(), nc);
    // bitstream: already in device memory when the decoded file was mirrored there (decompress_one), wherever the file
    // has it -- the readers take a 4-byte aligned base and a bit shift; otherwise uploaded here, aligned, zero padded
    const uint8_t *d_bits;
    unsigned bit_shift = 0;
    if (ws.raw_dev && bits >= ws.raw_host) {
        const size_t off = static_cast<size_t>(bits - ws.raw_host);
        d_bits = ws.raw_dev + (off & ~static_cast<size_t>(3));
        bit_shift = static_cast<unsigned>(off & 3) * 8;
    } else {
        const size_t padded = (enc_len + 3) / 4 * 4 + 64;
        uint8_t *d = ws.hd_bits.as<uint8_t>(padded);
        SZ3B_CUDA(cudaMemsetAsync(d + enc_len / 4 * 4, 0, padded - enc_len / 4 * 4, ws.st));
        ws.h2d(d, bits, enc_len);
        d_bits = d;
    }
    const uint64_t nsub = hd_num_sub(total_bits);
    if (nsub >= 0xfffffff0ull) fail(SZ3B_E_UNSUPPORTED, "Huffman stream above 2^42 bits");
    // overshoots (1 byte each), then two work lists of subsequence indices (ping-pong)
    const size_t over_bytes = (nsub + 8 + 15) & ~static_cast<size_t>(15);
    uint8_t *d_over = ws.hd_over.as<uint8_t>(over_bytes + 2 * (nsub + 1) * sizeof(uint32_t));
    uint32_t *list_a = reinterpret_cast<uint32_t *>(d_over + over_bytes), *list_b = list_a + nsub + 1;
    unsigned *d_counts = ws.hd_counts.as<unsigned>(nsub + 2);
    unsigned long long *d_offs = ws.hd_offs.as<unsigned long long>(2 * (nsub + 2) + scan_scratch_words(nsub));
    unsigned long long *d_moved = ws.counters.as<unsigned long long>(4);
    SZ3B_CUDA(cudaMemsetAsync(d_over, 0, over_bytes, ws.st));
    HdDeviceTables tb{d_lut, d_lut2, d_L, d_R, d_C, d_leaf, dec.offset};
    ws.stage_end(h, 0);
    h = ws.stage_begin("huffman_decode_sync");
    uint8_t *in = d_over;
    int launches = 0;
    bool converged = false;
    unsigned long long n_in = 0;
    // every round makes at least the first not yet exact subsequence exact, so nsub rounds always suffice
    for (uint64_t round = 1; round <= nsub + 1 && !converged; round++) {
        SZ3B_CUDA(cudaMemsetAsync(d_moved, 0, sizeof(unsigned long long), ws.st));
        launch_hd_sync(reinterpret_cast<const uint32_t *>(d_bits), bit_shift, total_bits, tb, in, round == 1 ? nullptr : list_a, n_in, list_b,
                       d_counts, d_moved, ws.st);
        launches++;
        // read back through pinned memory (a pageable readback costs more than a late round itself)
        unsigned long long *moved = static_cast<unsigned long long *>(ws.hist_host.ensure(64));
        ws.d2h(moved, d_moved, sizeof(unsigned long long));
        SZ3B_CUDA(stream_wait(ws.st));
        n_in = moved[0];
        converged = n_in == 0;
        std::swap(list_a, list_b);   // what moved in this round feeds the next one
    }
    if (!converged) fail(SZ3B_E_RUNTIME, "Huffman stream did not self-synchronise (malformed stream?)");
    if (getenv("SZ3B_VERBOSE")) fprintf(stderr, "[sz3b] huffman decode: %d synchronisation rounds over %llu subsequences\n", launches,
                                        static_cast<unsigned long long>(nsub));
    // `in` now holds the fixed point, d_counts the matching symbol counts
    ws.stage_end(h, launches);
    h = ws.stage_begin("huffman_decode_write");
    launch_scan_chunks(d_counts, d_counts, nsub, d_offs, d_offs + nsub + 2, ws.st, scan_scratch_words(nsub) ? d_offs + 2 * (nsub + 2) : nullptr);
    unsigned long long total = 0;
    ws.d2h(&total, d_offs + nsub, sizeof(total));
    SZ3B_CUDA(stream_wait(ws.st));
    if (total < n) fail(SZ3B_E_INVALID_ARGUMENT, "huffman: bitstream exhausted");
    launch_hd_write<QT>(reinterpret_cast<const uint32_t *>(d_bits), bit_shift, total_bits, tb, in, d_offs, n, d_q, ws.st);
    ws.stage_end(h, 4);
    SZ3B_CUDA(cudaGetLastError());
    return d_q;
}

// position-indexed unpredictable values on the device
template <class T, class QT>
static T *place_unpred(Workspace &ws, const QT *d_q, uint64_t n, const T *h_unpred, uint64_t n_unpred, int *launches) {
    T *d_tmp = ws.unpred_tmp.as<T>(n);
    if (n_unpred == 0) return d_tmp;
    T *d_un = ws.unpred_out.as<T>(n_unpred);
    ws.h2d(d_un, h_unpred, n_unpred * sizeof(T));
    const uint64_t nch = zero_num_chunks(n);
    unsigned *cz = ws.chunk_zeros.as<unsigned>(nch + 1);
    unsigned *cb = ws.chunk_bits.as<unsigned>(nch + 1);
    unsigned long long *zo = ws.zero_off.as<unsigned long long>(nch + 2 + scan_scratch_words(nch));
    unsigned long long *bo = ws.bit_off.as<unsigned long long>(nch + 2);
    launch_zero_count<QT>(d_q, n, cz, cb, ws.st);
    launch_scan_chunks(cb, cz, nch, bo, zo, ws.st, scan_scratch_words(nch) ? zo + nch + 2 : nullptr);
    launch_zero_scatter<QT, T>(d_q, n, zo, d_un, n_unpred, d_tmp, ws.st);
    *launches += 3;
    return d_tmp;
}

template <class T, class QT>
static void interp_decompress_t(Workspace &ws, const sz3b_config &conf, Cursor &c, T *d_out, int radius_hint) {
    // InterpolationDecomposition::load (:161-174)
    sz3b_config ic = conf;
    for (int d = 0; d < conf.N; d++) {
        const uint64_t dim = c.get<uint64_t>();
        if (dim != conf.dims[d]) fail(SZ3B_E_INVALID_ARGUMENT, "decomposition dims do not match the Config");
    }
    const uint32_t blocksize = c.get<uint32_t>();
    if (blocksize != static_cast<uint32_t>(kInterpBlock)) fail(SZ3B_E_UNSUPPORTED, "interpolation block size other than 32");
    ic.interpAlgo = c.get<int32_t>();
    ic.interpDirection = c.get<int32_t>();
    ic.interpAnchorStride = static_cast<int32_t>(c.get<uint64_t>());
    ic.interpAlpha = c.get<double>();
    ic.interpBeta = c.get<double>();
    double eb;
    int radius;
    const T *h_unpred;
    uint64_t n_unpred;
    quantizer_load<T>(c, &eb, &radius, &h_unpred, &n_unpred);
    if ((radius <= 32768) != (sizeof(QT) == 2)) fail(SZ3B_E_INVALID_ARGUMENT, "quantizer radius does not match Config.quantbinCnt");
    (void)radius_hint;
    InterpPlan pl;
    if (const char *e = build_interp_plan(ic, eb, 1, pl)) fail(SZ3B_E_INVALID_ARGUMENT, e);
    QT *d_q = decode_indices<QT>(ws, c, pl.num);
    int launches = 0;
    size_t h = ws.stage_begin("recover");
    T *d_tmp = place_unpred<T, QT>(ws, d_q, pl.num, h_unpred, n_unpred, &launches);
    uint64_t *d_table = ws.tables.as<uint64_t>(pl.table.size() + 1);
    if (!pl.table.empty()) ws.h2d(d_table, pl.table.data(), pl.table.size() * sizeof(uint64_t));
    launch_interp_recover<T, QT>(pl.sh, d_out, d_q, d_tmp, make_quant(pl.eb, radius), 0, nullptr, nullptr, -1,
                                 pl.anchor_stride, pl.n_first, ws.st);
    launches++;
    for (const LevelPlan &L : pl.levels) {
        InterpArgs<T, QT> A;
        memset(&A, 0, sizeof(A));
        A.sh = pl.sh;
        A.work = d_out;
        A.q = d_q;
        A.data_bstride = A.q_bstride = pl.num;
        A.qp = make_quant(L.eb, radius);
        A.s = L.s;
        for (int d = 0; d < kMaxDim; d++) A.nb[d] = L.nb[d];
        A.block_base = d_table + L.table_off;
        for (int p = 0; p < pl.sh.N; p++) {
            if (pass_points(A, p) == 0) continue;
            static const bool old_recover = getenv("SZ3B_RECOVER_OLD") != nullptr;   // diagnostics
            // small passes (the coarse levels) on the point-mapped kernel: a row-mapped CTA walks whole rows, and with
            // a handful of CTAs its latency (60-100 us per pass) is all there is
            const bool big = pass_points(A, p) >= (2ull << 20);
            // the last pass (along x) of the finest level: the box schedule's row phase in recover mode (interp_box.cuh)
            if constexpr (std::is_same<T, float>::value && std::is_same<QT, uint16_t>::value) {
                if (L.s == 1 && p == pl.sh.N - 1 && pl.sh.N == 3 && pl.sh.perm[2] == 2 && !old_recover && pl.num < (1ull << 32)) {
                    InterpArgs<T, QT> B = A;
                    B.unpred_tmp = d_tmp;
                    if (interp_launch_box_recover_x(B, d_out, L.nblocks, ws.st)) {
                        launches++;
                        continue;
                    }
                }
            }
            if (!(pl.sh.N >= 3 && big && !old_recover && interp_launch_lean<T, QT>(A, p, 1, true, true, d_tmp, ws.st)))
                launch_interp_recover<T, QT>(pl.sh, d_out, d_q, d_tmp, make_quant(L.eb, radius), L.s, L.nb,
                                             d_table + L.table_off, p, 0, 0, ws.st);
            launches++;
        }
    }
    ws.stage_end(h, launches);
    SZ3B_CUDA(cudaGetLastError());
}

// BlockwiseDecomposition::decompress with a Lorenzo predictor in the stack (BlockwiseDecomposition.hpp:48-67,75-79;
// ComposedPredictor::load :66-78; RegressionPredictor::load :109-123): the selection and the coefficient indices come
// from the stream, so one wavefront pass in recover mode suffices.
template <class T, class QT>
static void blockwise_decompress_lorenzo(Workspace &ws, const sz3b_config &conf, Cursor &c, T *d_out) {
    const int N = conf.N, nc = N + 1;
    if (conf.blockSize < 1) fail(SZ3B_E_INVALID_ARGUMENT, "blockSize must be positive");
    BlockShape bs;
    block_shape_init(bs, N, conf.dims, static_cast<uint32_t>(conf.blockSize));
    if (bs.nblocks >= 0xfffffff0ull) fail(SZ3B_E_UNSUPPORTED, "more than 2^32 blocks");
    BwArgs<T, QT> A;
    const uint64_t np = bw_args_init<T, QT>(A, conf, bs, 1.0);
    bw_upload_diag_table<T, QT>(ws, A);
    const bool has_reg = conf.regression != 0;
    const int reg_sid = A.nk - 1;
    // predictors in stack order: only the regression predictor stores anything
    uint64_t n_coef = 0;
    double eb_i = 0, eb_l = 0;
    int rad_i = 0, rad_l = 0;
    std::vector<int32_t> cq;
    std::vector<T> cun;
    if (has_reg) {
        n_coef = c.get<uint64_t>();
        if (n_coef % nc || n_coef / nc > bs.nblocks) fail(SZ3B_E_INVALID_ARGUMENT, "coefficient count does not match the block grid");
        if (n_coef) {
            const T *un_i, *un_l;
            uint64_t nun_i, nun_l;
            quantizer_load<T>(c, &eb_i, &rad_i, &un_i, &nun_i);
            quantizer_load<T>(c, &eb_l, &rad_l, &un_l, &nun_l);
            cq.resize(n_coef);
            HuffmanDecoder dec;
            const char *err = nullptr;
            if (!dec.load(c.p, c.rem, &err)) fail(SZ3B_E_INVALID_ARGUMENT, err);
            if (!dec.decode<int32_t>(c.p, c.rem, n_coef, cq.data(), &err)) fail(SZ3B_E_INVALID_ARGUMENT, err);
            cun.assign(n_coef, 0);
            uint64_t ki = 0, kl = 0;
            for (uint64_t pos = 0; pos < n_coef; pos++)
                if (cq[pos] == 0) {
                    if (pos % nc == static_cast<unsigned>(N)) {
                        if (ki >= nun_i) fail(SZ3B_E_INVALID_ARGUMENT, "truncated stream (coefficients)");
                        cun[pos] = un_i[ki++];
                    } else {
                        if (kl >= nun_l) fail(SZ3B_E_INVALID_ARGUMENT, "truncated stream (coefficients)");
                        cun[pos] = un_l[kl++];
                    }
                }
        }
    }
    std::vector<uint8_t> sel8;
    if (A.nk > 1) {
        const uint64_t n_sel = c.get<uint64_t>();
        if (n_sel != bs.nblocks) fail(SZ3B_E_INVALID_ARGUMENT, "selection count does not match the block grid");
        std::vector<int32_t> sel(n_sel);
        HuffmanDecoder dec;
        const char *err = nullptr;
        if (!dec.load(c.p, c.rem, &err)) fail(SZ3B_E_INVALID_ARGUMENT, err);
        if (!dec.decode<int32_t>(c.p, c.rem, n_sel, sel.data(), &err)) fail(SZ3B_E_INVALID_ARGUMENT, err);
        sel8.resize(n_sel);
        uint64_t n_reg = 0;
        for (uint64_t b = 0; b < n_sel; b++) {
            if (sel[b] < 0 || sel[b] >= A.nk) fail(SZ3B_E_INVALID_ARGUMENT, "predictor selection out of range");
            sel8[b] = static_cast<uint8_t>(sel[b]);
            n_reg += has_reg && sel[b] == reg_sid;
        }
        if (has_reg && n_reg * nc != n_coef) fail(SZ3B_E_INVALID_ARGUMENT, "coefficient count does not match the selection");
    }
    double eb;
    int radius;
    const T *h_unpred;
    uint64_t n_unpred;
    quantizer_load<T>(c, &eb, &radius, &h_unpred, &n_unpred);
    if ((radius <= 32768) != (sizeof(QT) == 2)) fail(SZ3B_E_INVALID_ARGUMENT, "quantizer radius does not match Config.quantbinCnt");
    A.qp = make_quant(eb, radius);
    QT *d_q = decode_indices<QT>(ws, c, bs.num);
    int launches = 0;
    size_t h = ws.stage_begin("recover");
    if (A.nk > 1) {
        uint8_t *d_sel = ws.bsel.as<uint8_t>(bs.nblocks);
        ws.h2d(d_sel, sel8.data(), bs.nblocks);
        A.sel_in = d_sel;
        if (has_reg) {
            uint32_t *rank = ws.brank.as<uint32_t>(bs.nblocks + 1);
            unsigned long long *counters = ws.counters.as<unsigned long long>(4);
            launch_bw_rank(d_sel, 0, bs.nblocks, reg_sid, 0, rank, counters + 2, ws.st);
            launches++;
            A.rank = rank;
            if (n_coef) {
                int32_t *d_cq = ws.coef_q.as<int32_t>(n_coef);
                T *d_cun = ws.coef.as<T>(n_coef);
                T *d_crec = ws.coef2.as<T>(n_coef);
                ws.h2d(d_cq, cq.data(), n_coef * sizeof(int32_t));
                ws.h2d(d_cun, cun.data(), n_coef * sizeof(T));
                launch_reg_chain_recover<T>(d_cq, d_cun, n_coef / nc, N, make_quant(eb_l, rad_l), make_quant(eb_i, rad_i), d_crec,
                                            ws.st);
                launches++;
                A.c_rec = d_crec;
            }
        }
    }
    A.unpred_tmp = place_unpred<T, QT>(ws, d_q, bs.num, h_unpred, n_unpred, &launches);
    A.q = d_q;
    A.out = d_out;
    A.W = ws.padded.as<T>(np);
    SZ3B_CUDA(cudaMemsetAsync(A.W, 0, np * sizeof(T), ws.st));
    A.mode = BW_DECODE;
    if (const char *e = launch_bw_fronts<T, QT>(A, ws.st, &launches)) fail(SZ3B_E_UNSUPPORTED, e);
    ws.stage_end(h, launches);
    SZ3B_CUDA(stream_wait(ws.st));   // cq / cun / sel8 are host vectors
    SZ3B_CUDA(cudaGetLastError());
}

template <class T, class QT>
static void blockwise_decompress_t(Workspace &ws, const sz3b_config &conf, Cursor &c, T *d_out) {
    const int N = conf.N;
    if (conf.lorenzo || conf.lorenzo2) {
        blockwise_decompress_lorenzo<T, QT>(ws, conf, c, d_out);
        return;
    }
    if (!conf.regression) fail(SZ3B_E_INVALID_ARGUMENT, "All lorenzo and regression methods are disabled.");
    BlockShape bs;
    block_shape_init(bs, N, conf.dims, static_cast<uint32_t>(conf.blockSize));
    const int nc = N + 1;
    // RegressionPredictor::load (RegressionPredictor.hpp:109-123)
    const uint64_t n_coef = c.get<uint64_t>();
    if (n_coef != bs.nblocks * nc) fail(SZ3B_E_UNSUPPORTED, "regression stream with fallback blocks (extent-1 blocks)");
    double eb_i, eb_l;
    int rad_i, rad_l;
    const T *un_i, *un_l;
    uint64_t nun_i, nun_l;
    quantizer_load<T>(c, &eb_i, &rad_i, &un_i, &nun_i);
    quantizer_load<T>(c, &eb_l, &rad_l, &un_l, &nun_l);
    // the coefficient indices through the GPU decoder as well (a million symbols took the host decoder ~10 ms); they
    // come back through pinned memory for the one serial step left: handing out the stored exact coefficients
    int32_t *d_cq = ws.coef_q.as<int32_t>(n_coef);
    {
        uint32_t *d_tmpq = decode_indices<uint32_t>(ws, c, n_coef, false);   // (in ws.q, which the main indices take next)
        SZ3B_CUDA(cudaMemcpyAsync(d_cq, d_tmpq, n_coef * sizeof(int32_t), cudaMemcpyDeviceToDevice, ws.st));
    }
    uint8_t *pin = static_cast<uint8_t *>(ws.stage2.ensure(n_coef * (sizeof(int32_t) + sizeof(T)) + 64));
    int32_t *cq = reinterpret_cast<int32_t *>(pin);
    T *cun = reinterpret_cast<T *>(pin + ((n_coef * sizeof(int32_t) + 15) & ~static_cast<size_t>(15)));
    ws.d2h(cq, d_cq, n_coef * sizeof(int32_t));
    SZ3B_CUDA(stream_wait(ws.st));
    // stored exact coefficients, by position (both quantizers consume their lists in block order)
    memset(cun, 0, n_coef * sizeof(T));
    {
        uint64_t ki = 0, kl = 0;
        for (uint64_t pos = 0; pos < n_coef; pos++)
            if (cq[pos] == 0) {
                if (pos % nc == static_cast<unsigned>(N)) {
                    if (ki >= nun_i) fail(SZ3B_E_INVALID_ARGUMENT, "truncated stream (coefficients)");
                    cun[pos] = un_i[ki++];
                } else {
                    if (kl >= nun_l) fail(SZ3B_E_INVALID_ARGUMENT, "truncated stream (coefficients)");
                    cun[pos] = un_l[kl++];
                }
            }
    }
    double eb;
    int radius;
    const T *h_unpred;
    uint64_t n_unpred;
    quantizer_load<T>(c, &eb, &radius, &h_unpred, &n_unpred);
    if ((radius <= 32768) != (sizeof(QT) == 2)) fail(SZ3B_E_INVALID_ARGUMENT, "quantizer radius does not match Config.quantbinCnt");
    QT *d_q = decode_indices<QT>(ws, c, bs.num);
    int launches = 0;
    size_t h = ws.stage_begin("recover");
    T *d_cun = ws.coef.as<T>(n_coef);
    T *d_crec = ws.coef2.as<T>(n_coef);
    ws.h2d(d_cun, cun, n_coef * sizeof(T));
    launch_reg_chain_recover<T>(d_cq, d_cun, bs.nblocks, N, make_quant(eb_l, rad_l), make_quant(eb_i, rad_i), d_crec, ws.st);
    T *d_tmp = place_unpred<T, QT>(ws, d_q, bs.num, h_unpred, n_unpred, &launches);
    if (const char *e = launch_reg_recover<T, QT>(d_out, bs, d_crec, make_quant(eb, radius), d_q, d_tmp, ws.st))
        fail(SZ3B_E_UNSUPPORTED, e);
    ws.stage_end(h, launches + 2);
    SZ3B_CUDA(stream_wait(ws.st));   // cq / cun are host vectors
    SZ3B_CUDA(cudaGetLastError());
}

// Length of the part of an interpolation stream the host parser reads (decomposition header, quantizer with its stored
// values, Huffman tree, the two counts in front of the bits), from the first `valid` decoded bytes.  Returns false
// with *need = the number of bytes that must be valid before the answer is known.
template <class T>
static bool interp_head_len(const uint8_t *raw, size_t valid, size_t raw_len, int N, size_t *need) {
    const size_t a = static_cast<size_t>(N) * 8 + 36;   // dims | blocksize | interpAlgo | direction | anchor stride | alpha | beta
    const size_t b = a + 21;                            // uid | eb | radius | number of stored values
    if (valid < b) {
        *need = b;
        return false;
    }
    uint64_t n_unpred;
    memcpy(&n_unpred, raw + a + 13, 8);
    if (n_unpred > raw_len / sizeof(T)) fail(SZ3B_E_INVALID_ARGUMENT, "truncated stream (unpredictable values)");
    const size_t c = b + static_cast<size_t>(n_unpred) * sizeof(T);   // Huffman tree: offset | node count (big endian) | ...
    if (valid < c + 13) {
        *need = c + 13;
        return false;
    }
    const uint32_t nc = (static_cast<uint32_t>(raw[c + 4]) << 24) | (static_cast<uint32_t>(raw[c + 5]) << 16) |
                        (static_cast<uint32_t>(raw[c + 6]) << 8) | raw[c + 7];
    const size_t lw = nc <= 256 ? 1 : (nc <= 65536 ? 2 : 4);
    *need = c + 13 + 2 * static_cast<size_t>(nc) * lw + static_cast<size_t>(nc) * 5 + 16;
    return valid >= *need;
}

// Frames of the shape the GPU lossless stage writes, decoded on the GPU (zhuf_dec.cuh): the compressed payload goes up
// once, one CTA per block writes the decoded stream into the device mirror the Huffman decoder reads, and the host
// decodes (libzstd) only the leading frames that hold what its parser reads.  Returns false -- nothing changed -- when
// the payload is not of that shape or the head is most of the stream; the caller then decodes every frame on the host.
// `flag` receives the address the kernel's verdict arrives at (pinned; valid after the next synchronisation of ws.st).
template <class T>
static bool frames_on_gpu(Workspace &ws, const sz3b_config &conf, const uint8_t *cmp, size_t cmp_size, uint8_t *raw, size_t raw_len,
                          const unsigned **flag) {
    const uint8_t *pay = cmp + 8;
    const size_t pay_size = cmp_size - 8;
    const double t0 = now_ms();
    const size_t cap = 2 * (raw_len / kZhufBlock + raw_len / kZhufFrame) + 64;
    ZhufDecBlock *h_blocks = static_cast<ZhufDecBlock *>(ws.zdec_host.ensure((cap + 1) * sizeof(ZhufDecBlock) + 64));
    size_t nblocks = 0;
    uint64_t total = 0;
    if (!zhuf_walk_frames(pay, pay_size, h_blocks, cap, &nblocks, &total) || total != raw_len || nblocks == 0) return false;
    // the head of the stream on the host: libzstd's streaming decoder with the output capped at what the parser reads
    // (it works block by block, so the first 128 KiB or so are decoded, not the whole first frame)
    {
        struct Holder {   // (container slabs are decoded by short-lived threads: the context goes with its thread)
            ZSTD_DCtx *p = nullptr;
            ~Holder() {
                if (p) ZSTD_freeDCtx(p);
            }
        };
        thread_local Holder held;
        if (!held.p) held.p = ZSTD_createDCtx();
        ZSTD_DCtx *const dctx = held.p;
        if (!dctx) return false;
        ZSTD_DCtx_reset(dctx, 1);
        ZSTD_inBuffer in{pay, pay_size, 0};
        ZSTD_outBuffer ob{raw, 0, 0};
        size_t need = 0;
        while (!interp_head_len<T>(raw, ob.pos, raw_len, conf.N, &need)) {
            if (need > raw_len / 2) return false;
            ob.size = need;
            while (ob.pos < ob.size) {
                const size_t before_in = in.pos, before_out = ob.pos;
                const size_t r = ZSTD_decompressStream(dctx, &ob, &in);
                if (ZSTD_isError(r)) return false;
                if (in.pos == before_in && ob.pos == before_out) return false;   // no progress: truncated payload
            }
        }
    }
    ws.host_stage("frames_walk_head_host", now_ms() - t0);
    uint8_t *d_cmp = ws.zcmp.as<uint8_t>(pay_size + 64);
    size_t h = ws.stage_begin("frames_upload");
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, pay) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (!pinned && pay_size >= (static_cast<size_t>(32) << 20)) {
        upload_pageable(ws, d_cmp, pay, pay_size);
    } else {
        SZ3B_CUDA(cudaMemcpyAsync(d_cmp, pay, pay_size, cudaMemcpyHostToDevice, ws.st));
        ws.h2d_bytes += pay_size;
    }
    ws.stage_end(h, 0);
    ZhufDecBlock *d_blocks = ws.zdec_blocks.as<ZhufDecBlock>(nblocks + 1);
    h = ws.stage_begin("frames_gpu");
    unsigned *d_bad = reinterpret_cast<unsigned *>(d_blocks + nblocks);
    SZ3B_CUDA(cudaMemcpyAsync(d_blocks, h_blocks, nblocks * sizeof(ZhufDecBlock), cudaMemcpyHostToDevice, ws.st));
    SZ3B_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(ZhufDecBlock), ws.st));
    uint8_t *d_raw = ws.hd_bits.as<uint8_t>(raw_len + 256);
    SZ3B_CUDA(cudaMemsetAsync(d_raw + raw_len, 0, 256, ws.st));
    launch_zhuf_decode(d_cmp, d_blocks, nblocks, d_raw, d_bad, ws.st);
    unsigned *h_flag = reinterpret_cast<unsigned *>(h_blocks + cap);
    *h_flag = 0xffffffffu;
    SZ3B_CUDA(cudaMemcpyAsync(h_flag, d_bad, sizeof(unsigned), cudaMemcpyDeviceToHost, ws.st));
    ws.stage_end(h, 1);
    SZ3B_CUDA(cudaGetLastError());
    *flag = h_flag;
    ws.raw_host = raw;
    ws.raw_dev = d_raw;
    return true;
}

// one non-OMP payload -> `out` (host or device pointer to config_num(conf) elements)
template <class T>
static void decompress_one(Workspace &ws, const sz3b_config &conf, const uint8_t *cmp, size_t cmp_size, T *out, int loc,
                           bool gpu_frames_allowed = true) {
    const uint64_t num = config_num(conf);
    const size_t bytes = num * sizeof(T);
    if (conf.cmprAlgo == SZ3B_ALGO_LOSSLESS) {
        // SZDispatcher.hpp:83-88: zstd(raw array)
        uint8_t *dst = loc == SZ3B_HOST ? reinterpret_cast<uint8_t *>(out) : static_cast<uint8_t *>(ws.stage.ensure(bytes));
        size_t raw = 0;
        double t0 = now_ms();
        if (!zstd_decompress_parallel(cmp, cmp_size, dst, bytes, &raw, host_threads()) || raw != bytes)
            fail(SZ3B_E_RUNTIME, "lossless decompression failed (size mismatch)");
        ws.host_stage("zstd_host", now_ms() - t0);
        if (loc == SZ3B_DEVICE) {
            ws.h2d(out, dst, bytes);
            SZ3B_CUDA(stream_wait(ws.st));
        }
        return;
    }
    if (conf.cmprAlgo != SZ3B_ALGO_INTERP && conf.cmprAlgo != SZ3B_ALGO_LORENZO_REG) {
        if (conf.cmprAlgo == SZ3B_ALGO_INTERP_LORENZO || conf.cmprAlgo == SZ3B_ALGO_NOPRED || conf.cmprAlgo == 5 || conf.cmprAlgo == 6)
            fail(SZ3B_E_UNSUPPORTED, "stream algorithm outside the GPU hot path");
        fail(SZ3B_E_INVALID_ARGUMENT, "Unknown compression algorithm");
    }
    const size_t raw_len = zstd_framed_raw_len(cmp, cmp_size);
    if (raw_len == 0 || raw_len > (bytes + (static_cast<size_t>(1) << 20)) * 4) fail(SZ3B_E_INVALID_ARGUMENT, "implausible stream length");
    uint8_t *raw = static_cast<uint8_t *>(ws.stage.ensure(raw_len + 16));
    const unsigned *gpu_flag = nullptr;
    const bool on_gpu = gpu_frames_allowed && frame_decoder() == 1 && conf.cmprAlgo == SZ3B_ALGO_INTERP && cmp_size > 8 &&
                        raw_len >= kZhufMinStream && frames_on_gpu<T>(ws, conf, cmp, cmp_size, raw, raw_len, &gpu_flag);
    if (!on_gpu) {
        // every frame goes up to the device mirror of the file as soon as it is decoded (the Huffman decoder reads the
        // bit stream there): the upload hides under the decoding of the other frames
        struct Uploader : FrameDone {
            uint8_t *d;
            const uint8_t *h;
            cudaStream_t st;
            std::atomic<bool> failed{false};
            void frame(size_t off, size_t len) override {
                if (cudaMemcpyAsync(d + off, h + off, len, cudaMemcpyHostToDevice, st) != cudaSuccess) failed = true;
            }
        } up;
        up.d = ws.hd_bits.as<uint8_t>(raw_len + 256);
        up.h = raw;
        up.st = ws.st_copy;
        SZ3B_CUDA(cudaMemsetAsync(up.d + raw_len, 0, 256, ws.st_copy));
        size_t got = 0;
        double t0 = now_ms();
        if (!zstd_decompress_parallel(cmp, cmp_size, raw, raw_len, &got, host_threads(), &up)) fail(SZ3B_E_RUNTIME, "zstd decompression failed");
        if (up.failed) fail(SZ3B_E_CUDA, "upload of the decoded stream failed");
        cudaEvent_t ev = ws.event();
        SZ3B_CUDA(cudaEventRecord(ev, ws.st_copy));
        SZ3B_CUDA(cudaStreamWaitEvent(ws.st, ev, 0));
        ws.h2d_bytes += raw_len;
        ws.raw_host = raw;
        ws.raw_dev = up.d;
        ws.host_stage("zstd_host", now_ms() - t0);
    }
    struct RawReset {   // the mirror is only valid for this stream
        Workspace &w;
        ~RawReset() { w.raw_host = w.raw_dev = nullptr; }
    } raw_reset{ws};
    Cursor c{raw, raw_len};
    T *d_out = loc == SZ3B_DEVICE ? out : ws.data.as<T>(num);
    const bool narrow = conf.quantbinCnt / 2 <= 32768;
    if (conf.cmprAlgo == SZ3B_ALGO_INTERP) {
        if (narrow)
            interp_decompress_t<T, uint16_t>(ws, conf, c, d_out, 0);
        else
            interp_decompress_t<T, uint32_t>(ws, conf, c, d_out, 0);
    } else {
        if (narrow)
            blockwise_decompress_t<T, uint16_t>(ws, conf, c, d_out);
        else
            blockwise_decompress_t<T, uint32_t>(ws, conf, c, d_out);
    }
    if (loc == SZ3B_HOST) {
        size_t h = ws.stage_begin("d2h_output");
        cudaPointerAttributes at;
        const bool pinned = cudaPointerGetAttributes(&at, out) == cudaSuccess && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (!pinned && bytes >= (static_cast<size_t>(32) << 20))
            download_pageable(ws, out, d_out, bytes);
        else
            ws.d2h(out, d_out, bytes);
        ws.stage_end(h, 0);
    }
    SZ3B_CUDA(stream_wait(ws.st));
    if (on_gpu && *gpu_flag != 0) {
        // a block the GPU frame decoder did not take (its verdict travelled behind the kernel): everything again with
        // libzstd, which either decodes the payload or reports it as corrupt
        ws.raw_host = ws.raw_dev = nullptr;
        decompress_one<T>(ws, conf, cmp, cmp_size, out, loc, false);
    }
}

template <class T>
void decompress_any(Workspace &ws, sz3b_config &conf, const uint8_t *cmp, size_t cmp_size, T *out, int loc) {
    if (!conf.openmp) {
        decompress_one<T>(ws, conf, cmp, cmp_size, out, loc);
        return;
    }
    // SZ_decompress_OMP (SZImplOMP.hpp:120-186): int nThreads | Config x n | size_t x n | payloads
    Cursor c{cmp, cmp_size};
    const int n = c.get<int32_t>();
    if (n < 1 || static_cast<uint64_t>(n) > conf.dims[0]) fail(SZ3B_E_INVALID_ARGUMENT, "bad slab count in the OpenMP container");
    std::vector<sz3b_config> confs(n, conf);
    for (int t = 0; t < n; t++) {
        if (c.rem < 1) fail(SZ3B_E_INVALID_ARGUMENT, "truncated stream");
        const size_t len = c.p[0];
        confs[t].openmp = 0;
        if (!config_load(confs[t], c.take(len), len)) fail(SZ3B_E_INVALID_ARGUMENT, "malformed slab Config");
        confs[t].openmp = 0;
    }
    std::vector<uint64_t> sizes(n);
    for (int t = 0; t < n; t++) sizes[t] = c.get<uint64_t>();
    const uint64_t row = config_num(conf) / conf.dims[0];
    std::vector<const uint8_t *> payload(n);
    std::vector<uint64_t> first(n);
    for (int t = 0; t < n; t++) {
        const uint64_t lo = static_cast<uint64_t>(t) * conf.dims[0] / n, hi = static_cast<uint64_t>(t + 1) * conf.dims[0] / n;
        if (config_num(confs[t]) != (hi - lo) * row) fail(SZ3B_E_INVALID_ARGUMENT, "slab Config does not match the container");
        payload[t] = c.take(sizes[t]);
        first[t] = lo * row;
    }
    // The reference decodes the slabs on all its threads (SZImplOMP.hpp:148-180).  Here: a host-resident output is
    // filled by several workers, each with its own workspace and streams -- one per visible device (slab t on device
    // t mod G), two to four per device (a quarter of the host threads), so that the host part of a slab (zstd, tree
    // parsing) and its transfers run under the device part of the others: 512^3 in 8 slabs on one GPU decodes in
    // 35.9 ms with one worker, 19.2 ms with four.  A device-resident output stays on the calling thread (its device
    // owns the buffer).
    int workers = 1;
    if (loc == SZ3B_HOST && n > 1) workers = std::min(n, std::max(2, std::min(4, host_threads() / 4)) * device_fanout());
    if (const char *e = getenv("SZ3B_DEC_WORKERS")) workers = std::max(1, std::min(n, atoi(e)));   // diagnostics
    if (workers == 1) {
        for (int t = 0; t < n; t++) decompress_one<T>(ws, confs[t], payload[t], sizes[t], out + first[t], loc);
        return;
    }
    int visible = 1;
    cudaGetDeviceCount(&visible);
    const int ndev = std::max(1, std::min(device_fanout(), visible));
    std::vector<std::exception_ptr> errs(workers);
    auto body = [&](int g) {
        try {
            auto run = [&](Workspace &w) {
                for (int t = g; t < n; t += workers) decompress_one<T>(w, confs[t], payload[t], sizes[t], out + first[t], loc);
            };
            if (g == 0) {
                run(ws);
            } else {
                SZ3B_CUDA(cudaSetDevice((ws.device + g % ndev) % visible));
                WorkspaceLease w;
                run(*w);
            }
        } catch (...) {
            errs[g] = std::current_exception();
        }
    };
    std::vector<std::thread> th;
    for (int g = 1; g < workers; g++) th.emplace_back(body, g);
    body(0);
    for (auto &t : th) t.join();
    for (int g = 0; g < workers; g++)
        if (errs[g]) std::rethrow_exception(errs[g]);
}

// ---------------------------------------------------------------------------------------------------------------------
// stage-level entry points
// ---------------------------------------------------------------------------------------------------------------------
template <class T, class QT>
static void interp_decompose_t(Workspace &ws, const sz3b_config &conf, double eb, const T *d_data, int schedule,
                               int32_t *quant_out, std::vector<uint8_t> &blob) {
    sz3b_config c = conf;
    set_default_anchor(c);
    InterpPlan pl;
    if (const char *e = build_interp_plan(c, eb, schedule, pl)) fail(SZ3B_E_INVALID_ARGUMENT, e);
    const int radius = c.quantbinCnt / 2;
    const int nbins = 2 * radius;
    const uint64_t n = pl.num;
    QT *d_q = ws.q.as<QT>(n);
    T *d_unpred_tmp = ws.unpred_tmp.as<T>(n);
    unsigned long long *d_hist = ws.hist.as<unsigned long long>(nbins);
    size_t h = ws.stage_begin("predict_quantize");
    SZ3B_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nbins, ws.st));
    int launches = 0;
    run_interp<T, QT>(ws, pl, d_data, 1, radius, d_q, d_unpred_tmp, d_hist, ws.recon, &launches);
    ws.stage_end(h, launches);
    // indices -> host int32
    int32_t *d_wide = ws.side_q.as<int32_t>(n);
    launch_widen<QT>(d_q, n, d_wide, ws.st);
    ws.d2h(quant_out, d_wide, n * sizeof(int32_t));
    // unpredictables: reuse the packer's ordered compaction with a trivial (all length 0) code book
    HuffmanBook book;
    EncodeLayout lay;
    encode_indices<QT, T>(ws, d_q, n, d_hist, nbins, 0, true, d_unpred_tmp, book, lay);
    uint8_t hdr[128];
    size_t hdr_len = interp_save_header<T>(pl, radius, lay.n_unpred, hdr);
    blob.resize(hdr_len + lay.n_unpred * sizeof(T));
    memcpy(blob.data(), hdr, hdr_len);
    if (lay.n_unpred)
        ws.d2h(blob.data() + hdr_len, ws.unpred_out.p, lay.n_unpred * sizeof(T));
    SZ3B_CUDA(stream_wait(ws.st));
}

template <class T>
void interp_decompose_stage(Workspace &ws, const sz3b_config &conf, double eb, const T *data, int loc, int schedule,
                            int32_t *quant_out, std::vector<uint8_t> &blob) {
    const T *d = to_device(ws, data, loc, config_num(conf));
    if (conf.quantbinCnt / 2 <= 32768)
        interp_decompose_t<T, uint16_t>(ws, conf, eb, d, schedule, quant_out, blob);
    else
        interp_decompose_t<T, uint32_t>(ws, conf, eb, d, schedule, quant_out, blob);
}

// HuffmanEncoder<int> on an arbitrary int32 stream (side streams, tests)
void huffman_encode_device(Workspace &ws, const int32_t *d_q, size_t n, std::vector<uint8_t> &out, size_t *tree_len,
                           const std::function<void()> *after_hist) {
    if (n == 0) fail(SZ3B_E_INVALID_ARGUMENT, "Huffman bins should not be empty");
    int *d_mm = ws.misc.as<int>(2);
    int init[2] = {0x7fffffff, static_cast<int>(0x80000000)};
    ws.h2d(d_mm, init, sizeof(init));
    size_t h = ws.stage_begin("huffman_histogram");
    launch_minmax_int<int32_t>(d_q, n, d_mm, ws.st);
    int mm[2];
    ws.d2h(mm, d_mm, sizeof(mm));
    SZ3B_CUDA(stream_wait(ws.st));
    const int64_t span = static_cast<int64_t>(mm[1]) - mm[0] + 1;
    if (span > (1 << 26)) fail(SZ3B_E_UNSUPPORTED, "Huffman symbol range above 2^26");
    const int nbins = static_cast<int>(span);
    unsigned long long *d_hist = ws.hist.as<unsigned long long>(nbins);
    SZ3B_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nbins, ws.st));
    launch_histogram<int32_t>(d_q, n, mm[0], nbins, nbins / 2, d_hist, ws.st);
    ws.stage_end(h, 2);
    if (after_hist) (*after_hist)();   // (work the caller wants on the device while the host builds the tree)
    HuffmanBook book;
    EncodeLayout lay;
    encode_indices<int32_t, float>(ws, d_q, n, d_hist, nbins, mm[0], false, nullptr, book, lay);
    out.resize(lay.tree_len + 8 + lay.out_size);
    memcpy(out.data(), book.tree_blob.data(), lay.tree_len);
    uint8_t *p = out.data() + lay.tree_len;
    put<uint64_t>(p, lay.out_size);
    if (lay.out_size) ws.d2h(p, ws.out_words.p, lay.out_size);
    SZ3B_CUDA(stream_wait(ws.st));
    if (tree_len) *tree_len = lay.tree_len;
}

void huffman_encode_stage(Workspace &ws, const int32_t *q, size_t n, int loc, std::vector<uint8_t> &out,
                          size_t *tree_len) {
    const int32_t *d_q = q;
    if (loc == SZ3B_HOST) {
        int32_t *d = ws.side_q.as<int32_t>(n);
        ws.h2d(d, q, n * sizeof(int32_t));
        d_q = d;
    }
    huffman_encode_device(ws, d_q, n, out, tree_len);
}

// HuffmanEncoder::load + decode (HuffmanEncoder.hpp:225-279) of `in` = tree blob | size_t outSize | bits (what
// sz3b_huffman_encode and the reference's save() + encode() write) through the GPU decoder; `tree_len` as reported there.
void huffman_decode_stage(Workspace &ws, const uint8_t *in, size_t in_len, size_t tree_len, size_t n, int32_t *out) {
    if (tree_len > in_len) fail(SZ3B_E_INVALID_ARGUMENT, "tree length exceeds the input");
    std::vector<uint8_t> buf(in_len + 8);
    memcpy(buf.data(), in, tree_len);
    uint8_t *p = buf.data() + tree_len;
    put<uint64_t>(p, static_cast<uint64_t>(n));
    memcpy(p, in + tree_len, in_len - tree_len);
    Cursor c{buf.data(), buf.size()};
    uint32_t *d_q = decode_indices<uint32_t>(ws, c, n);
    ws.d2h(out, d_q, n * sizeof(uint32_t));
    SZ3B_CUDA(stream_wait(ws.st));
}

#define SZ3B_INST_PIPE(T)                                                                                             \
    template size_t compress_any<T>(Workspace &, sz3b_config &, const T *, int, uint8_t *, size_t);                  \
    template void interp_decompose_stage<T>(Workspace &, const sz3b_config &, double, const T *, int, int, int32_t *, \
                                            std::vector<uint8_t> &);                                                  \
    template void tune_stage<T>(Workspace &, sz3b_config &, const T *, int);                                         \
    template double abs_eb_stage<T>(Workspace &, const sz3b_config &, const T *, int);                               \
    template void minmax_stage<T>(Workspace &, const T *, int, size_t, double *, double *);                          \
    template size_t compress_slab<T>(Workspace &, sz3b_config &, const T *, int, double, uint8_t *, size_t);      \
    template void decompress_any<T>(Workspace &, sz3b_config &, const uint8_t *, size_t, T *, int);                  \
    template void blockwise_decompose_stage<T>(Workspace &, const sz3b_config &, double, const T *, int, int32_t *,  \
                                               std::vector<uint8_t> &);
SZ3B_INST_PIPE(float)
SZ3B_INST_PIPE(double)
// integer element types (tools/sz3/sz3.cpp:458-461): the whole-array entry points
#define SZ3B_INST_PIPE_INT(T)                                                                                         \
    template size_t compress_any<T>(Workspace &, sz3b_config &, const T *, int, uint8_t *, size_t);                  \
    template double abs_eb_stage<T>(Workspace &, const sz3b_config &, const T *, int);                               \
    template void minmax_stage<T>(Workspace &, const T *, int, size_t, double *, double *);                          \
    template size_t compress_slab<T>(Workspace &, sz3b_config &, const T *, int, double, uint8_t *, size_t);         \
    template size_t compress_slab_placed<T>(Workspace &, sz3b_config &, const T *, int, double, void *(*)(void *, size_t), void *); \
    template void decompress_any<T>(Workspace &, sz3b_config &, const uint8_t *, size_t, T *, int);
SZ3B_INST_PIPE_INT(int32_t)
SZ3B_INST_PIPE_INT(int64_t)

}  // namespace sz3b
