// sz3_b200/csrc/workspace.hpp -- per-call device/pinned scratch memory, stream and stage profile.
//
// cudaMalloc/cudaHostAlloc cost milliseconds, so buffers are cached per device in a small pool and only grow.
// A call borrows one Workspace for its whole duration (the library is reentrant: concurrent host threads get
// distinct workspaces and distinct streams).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string.h>

#include <sched.h>
#include <stdlib.h>

#include <atomic>
#include <chrono>
#include <string>
#include <vector>

namespace sz3b {

void launch_upload_bytes(void *dst, const void *src_mapped, size_t bytes, cudaStream_t st);   // misc_kernels.cu

struct CudaError {
    cudaError_t code;
    const char *what;
    const char *file;
    int line;
};

#define SZ3B_CUDA(expr)                                                         \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) throw sz3b::CudaError{_e, #expr, __FILE__, __LINE__}; \
    } while (0)

// How a host thread waits for the device.  0: the driver's own wait, which spins on a core -- lowest latency while
// every waiting thread has a core to itself.  1: poll and give the core away between polls -- for hosts where the
// waiting threads (one rank per GPU, each with concurrent tuner trials) outnumber the cores, so that a thread with
// real work (Huffman tree, zstd trial) is not time-sliced against spinners.  sz3b_set_host_wait() / SZ3B_HOST_WAIT.
inline std::atomic<int> &host_wait_mode() {
    static std::atomic<int> m{[] {
        const char *e = getenv("SZ3B_HOST_WAIT");
        return e ? atoi(e) : 0;
    }()};
    return m;
}
inline cudaError_t stream_wait(cudaStream_t st) {
    if (host_wait_mode().load(std::memory_order_relaxed) == 0) return cudaStreamSynchronize(st);
    cudaError_t e;
    while ((e = cudaStreamQuery(st)) == cudaErrorNotReady) sched_yield();
    return e;
}
inline cudaError_t event_wait(cudaEvent_t ev) {
    if (host_wait_mode().load(std::memory_order_relaxed) == 0) return cudaEventSynchronize(ev);
    cudaError_t e;
    while ((e = cudaEventQuery(ev)) == cudaErrorNotReady) sched_yield();
    return e;
}

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void *ensure(size_t bytes) {
        if (bytes > cap) {
            if (p) SZ3B_CUDA(cudaFree(p));
            p = nullptr;
            cap = 0;
            size_t want = bytes + bytes / 16 + 256;
            SZ3B_CUDA(cudaMalloc(&p, want));
            cap = want;
        }
        return p;
    }
    template <class V>
    V *as(size_t count) {
        return static_cast<V *>(ensure(count * sizeof(V)));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// true while a bulk input upload of any call is still on this device's H2D engine (pipeline.cu: UploadChain)
bool device_upload_busy(int device);

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    void *ensure(size_t bytes) {
        if (bytes > cap) {
            if (p) SZ3B_CUDA(cudaFreeHost(p));
            p = nullptr;
            cap = 0;
            size_t want = bytes + bytes / 16 + 4096;
            SZ3B_CUDA(cudaHostAlloc(&p, want, cudaHostAllocDefault));
            cap = want;
        }
        return p;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// Where a finished payload goes once its size is known (slabs of an OpenMP container: the offset of a slab is the sum
// of the sizes before it, so the compressed frames wait on the device until every slab has reported, and then cross
// PCIe straight to their final place -- no host-side copy of the payloads).  pipeline.cu: zhuf_run, dispatch_compress.
struct Placer {
    virtual uint8_t *place(size_t size) = 0;
    virtual ~Placer() {}
    bool used = false;
    uint64_t raw_bytes = 0;   // uncompressed size of the slab (the ratio < 3 rule decides before the payload is placed)
};

struct StageRecord {
    std::string name;
    double ms = 0;       // device time (CUDA events) or host wall time for host stages
    int launches = 0;    // kernels launched by this stage
    bool host = false;
};

struct Workspace {
    int device = 0;
    cudaStream_t st = nullptr;
    cudaStream_t st_copy = nullptr;   // bulk H2D of the input, so that sampling/tuning can overlap it
    cudaStream_t st_low = nullptr;    // lowest-priority compute stream (created on first use): bulk kernels that may run
                                      // under latency-critical small ones of `st`
    cudaStream_t low_priority_stream() {
        if (!st_low) {
            int least = 0, greatest = 0;
            SZ3B_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            SZ3B_CUDA(cudaStreamCreateWithPriority(&st_low, cudaStreamNonBlocking, least));
        }
        return st_low;
    }
    // decompression: the zstd-decoded stream (host, pinned) and its mirror in device memory, uploaded frame by frame
    // while the other frames are still being decoded (pipeline.cu: decompress_one); null when there is no mirror
    const uint8_t *raw_host = nullptr;
    const uint8_t *raw_dev = nullptr;
    cudaEvent_t ev_copy = nullptr;
    // inputs / index stream
    DevBuf data, q, unpred_tmp, recon, hist, tables, compact;
    // encoder
    DevBuf code, len, chunk_bits, chunk_zeros, bit_off, zero_off, out_words, unpred_out;
    // tuner
    DevBuf cubes, cube_q, cube_unpred, cube_recon, flags, starts;
    // side streams / blockwise
    DevBuf coef, coef2, coef_q, side_q, misc, counters, cpos, cval, hist2;
    // Lorenzo stacks (lorenzo.cu): padded working array, block selections, dense ranks, coefficient guesses
    DevBuf padded, bsel, bsel2, brank, cspec, cdense;
    // GPU lossless stage (zhuf_kernels.cu): assembled stream, per-block tables, compressed frames
    DevBuf zsrc, zinfo, zdst;
    // ... and its decoder (zhuf_dec.cuh): compressed payload, block list + flag; the host's copy of the list
    DevBuf zcmp, zdec_blocks;
    PinBuf zdec_host;
    // Huffman decode
    DevBuf hd_bits, hd_tab, hd_over, hd_counts, hd_offs;
    // pinned staging
    PinBuf stage, stage2, hist_host, slab_out;   // slab_out: payload of this device's slab of an OpenMP container
    std::vector<uint8_t> zscratch;   // per-chunk zstd frames before concatenation (host tail)
    std::vector<uint8_t> trial_out;  // compressed output of a tuner trial run on this workspace
    std::vector<uint8_t> slab_big;   // a container slab that needed the full capacity (nearly incompressible data)
    Placer *placer = nullptr;        // set for the duration of one dispatch_compress call
    // profiling
    std::vector<StageRecord> prof;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct Pending {
        size_t rec;
        cudaEvent_t a, b;
    };
    std::vector<Pending> pending;

    cudaEvent_t event() {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            SZ3B_CUDA(cudaEventCreate(&e));
            ev_pool.push_back(e);
        }
        return ev_pool[ev_used++];
    }
    // bytes moved across PCIe by the current call (reported by sz3b_last_transfer)
    size_t h2d_bytes = 0, d2h_bytes = 0;
    // While the bulk input copy owns the H2D copy engine (bulk_copy_in_flight), small uploads go through a pinned
    // staging arena read by an SM copy kernel instead of queueing behind it.
    bool bulk_copy_in_flight = false;
    // Plane-ordered bulk copy of a 3-D input (pipeline.cu: to_device_planes): the even planes of the outermost
    // dimension go first (everything the interpolation levels with stride >= 2 read), then the odd planes, one
    // group per block-row of the finest level, so that level-1 tiles start while later planes are still in flight.
    struct CopyPlan {
        bool active = false;
        cudaEvent_t ev_even = nullptr;
        std::vector<cudaEvent_t> ev_row;   // ev_row[b]: every plane the level-1 tiles of block-row b read has arrived
    } copy_plan;
    PinBuf upload_arena;
    PinBuf upload_ring;   // staging ring of upload_pageable (pipeline.cu)
    size_t upload_used = 0;
    static constexpr size_t kUploadArena = static_cast<size_t>(16) << 20;
    // Small uploads (code tables, block tables, headers: host vectors) go through a pinned arena: the copy is then truly
    // asynchronous (a cudaMemcpyAsync from pageable memory stages and synchronises inside the driver, tens of
    // microseconds each on the critical path), and while the bulk input copy owns the H2D engine an SM copy kernel
    // reads the arena instead of queueing behind it.
    void h2d(void *dst, const void *src, size_t bytes) {
        h2d_bytes += bytes;
        if (bytes <= (static_cast<size_t>(2) << 20)) {
            const size_t need = (bytes + 255) & ~static_cast<size_t>(255);
            unsigned char *arena = static_cast<unsigned char *>(upload_arena.ensure(kUploadArena));
            if (upload_used + need <= kUploadArena) {
                unsigned char *slot = arena + upload_used;
                upload_used += need;
                memcpy(slot, src, bytes);
                // (this call's bulk copy, or another caller's on the same device: a copy command would queue behind
                //  that caller's whole array)
                if ((bulk_copy_in_flight || device_upload_busy(device)) && (reinterpret_cast<uintptr_t>(dst) & 15) == 0)
                    launch_upload_bytes(dst, slot, bytes, st);
                else
                    SZ3B_CUDA(cudaMemcpyAsync(dst, slot, bytes, cudaMemcpyHostToDevice, st));
                return;
            }
        }
        SZ3B_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    }
    void d2h(void *dst, const void *src, size_t bytes) {
        SZ3B_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
        d2h_bytes += bytes;
    }
    void prof_reset() {
        h2d_bytes = d2h_bytes = 0;
        upload_used = 0;
        bulk_copy_in_flight = false;
        copy_plan.active = false;
        copy_plan.ev_row.clear();
        stage_prefix.clear();
        prof.clear();
        pending.clear();
        ev_used = 0;
    }
    // device stage: records events around the enclosed launches
    std::string stage_prefix;   // "tune_" while the auto-tuner's trial compressions run
    size_t stage_begin(const char *name) {
        StageRecord r;
        r.name = stage_prefix.empty() || !strncmp(name, "tune_", 5) ? std::string(name) : stage_prefix + name;
        prof.push_back(r);
        Pending pd;
        pd.rec = prof.size() - 1;
        pd.a = event();
        pd.b = event();
        SZ3B_CUDA(cudaEventRecord(pd.a, st));
        pending.push_back(pd);
        return pending.size() - 1;
    }
    void stage_end(size_t h, int launches) {
        SZ3B_CUDA(cudaEventRecord(pending[h].b, st));
        prof[pending[h].rec].launches += launches;
    }
    void host_stage(const char *name, double ms) {
        StageRecord r;
        r.name = stage_prefix.empty() || !strncmp(name, "tune_", 5) ? std::string(name) : stage_prefix + name;
        r.ms = ms;
        r.host = true;
        prof.push_back(r);
    }
    void prof_finish() {
        for (auto &pd : pending) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, pd.a, pd.b) == cudaSuccess) prof[pd.rec].ms = ms;
        }
        pending.clear();
    }
};

Workspace *workspace_acquire();
void workspace_release(Workspace *ws);

struct WorkspaceLease {
    Workspace *ws;
    WorkspaceLease() : ws(workspace_acquire()) {}
    ~WorkspaceLease() { workspace_release(ws); }
    Workspace *operator->() { return ws; }
    Workspace &operator*() { return *ws; }
};

inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace sz3b
