// sz3_b200/csrc/api.cu -- the extern "C" boundary declared in include/sz3b.h.
//
// Framing of SZ_compress / SZ_decompress (reference include/SZ3/api/sz.hpp:43-82,117-157): 16-byte header
// (magic, data version, payload size) | payload | Config blob.  Everything below the header is produced by the GPU
// pipelines in pipeline.cu / blockwise.cu / decompress.cu.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/sz3b.h"
#include "pipeline.hpp"
#include "stream_host.hpp"

using namespace sz3b;

namespace {
thread_local std::string t_last_error;
thread_local std::vector<StageRecord> t_profile;
thread_local size_t t_h2d = 0, t_d2h = 0;
thread_local cudaStream_t t_caller_stream = nullptr;
thread_local bool t_caller_stream_set = false;

// Device buffers handed in by the caller are read (or written) on the library's own non-blocking streams: order them
// after what the caller has queued.  With sz3b_set_caller_stream the library streams wait on an event recorded on that
// stream; without it the legacy default stream is synchronised (which covers callers that never create streams).
void after_caller(Workspace &ws, int loc) {
    if (loc != SZ3B_DEVICE) return;
    if (t_caller_stream_set) {
        cudaEvent_t e = ws.event();
        SZ3B_CUDA(cudaEventRecord(e, t_caller_stream));
        SZ3B_CUDA(cudaStreamWaitEvent(ws.st, e, 0));
        SZ3B_CUDA(cudaStreamWaitEvent(ws.st_copy, e, 0));
    } else {
        SZ3B_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    }
}

template <class F>
int guarded(F &&f) {
    try {
        f();
        t_last_error.clear();
        return SZ3B_OK;
    } catch (const Error &e) {
        t_last_error = e.msg;
        return e.code;
    } catch (const CudaError &e) {
        t_last_error = std::string("CUDA error: ") + cudaGetErrorString(e.code) + " in " + e.what + " (" + e.file +
                       ":" + std::to_string(e.line) + ")";
        cudaGetLastError();
        return SZ3B_E_CUDA;
    } catch (const std::exception &e) {
        t_last_error = e.what();
        return SZ3B_E_RUNTIME;
    }
}

void check_dtype(int dtype) {   // stage-level entry points: floating-point data only
    if (dtype != SZ3B_FLOAT && dtype != SZ3B_DOUBLE)
        fail(SZ3B_E_UNSUPPORTED, "this entry point takes float32 / float64 data");
}
void check_dtype_any(int dtype) {
    if (dtype != SZ3B_FLOAT && dtype != SZ3B_DOUBLE && dtype != SZ3B_INT32 && dtype != SZ3B_INT64)
        fail(SZ3B_E_UNSUPPORTED, "element types on the GPU path: float32, float64, int32, int64");
}
size_t dtype_size(int dtype) {   // SZ_FLOAT .. SZ_INT64 (def.hpp): sizes for the capacity rule, supported or not
    static const size_t sz[10] = {4, 8, 1, 1, 2, 2, 4, 4, 8, 8};
    return dtype >= 0 && dtype < 10 ? sz[dtype] : 8;
}
// f(T *) for the element type of `dtype` (a null pointer used as a type tag)
template <class F>
auto by_dtype(int dtype, F &&f) {
    switch (dtype) {
        case SZ3B_FLOAT: return f(static_cast<float *>(nullptr));
        case SZ3B_DOUBLE: return f(static_cast<double *>(nullptr));
        case SZ3B_INT32: return f(static_cast<int32_t *>(nullptr));
        default: return f(static_cast<int64_t *>(nullptr));
    }
}
#define SZ3B_TAG_T(tag) typename std::remove_pointer<decltype(tag)>::type

void check_conf(const sz3b_config *c) {
    if (!c) fail(SZ3B_E_INVALID_ARGUMENT, "null config");
    if (c->N < 1) fail(SZ3B_E_INVALID_ARGUMENT, "config has no dimensions");
    if (c->N > 4) fail(SZ3B_E_INVALID_ARGUMENT, "Data dimension higher than 4 is not supported.");
}

void finish_profile(Workspace &ws) {
    ws.prof_finish();
    t_profile = ws.prof;
    t_h2d = ws.h2d_bytes;
    t_d2h = ws.d2h_bytes;
}

size_t size_bound(int dtype, const sz3b_config &c) {
    const size_t esz = dtype_size(dtype);
    uint8_t blob[256];
    const size_t conf_est = config_save(c, blob);
    const uint64_t num = config_num(c);
    if (c.openmp) {
        // SZ_compress_size_bound_omp (SZImplOMP.hpp:189-209): header + per-slab zstd bounds
        int nslabs = c.openmp;
        if (static_cast<uint64_t>(nslabs) > c.dims[0]) nslabs = static_cast<int>(c.dims[0]);
        size_t total = sizeof(int) + static_cast<size_t>(nslabs) * (conf_est + 32 + sizeof(size_t));
        const uint64_t row = num / c.dims[0];
        for (int t = 0; t < nslabs; t++) {
            uint64_t lo = static_cast<uint64_t>(t) * c.dims[0] / nslabs, hi = static_cast<uint64_t>(t + 1) * c.dims[0] / nslabs;
            total += ZSTD_compressBound((hi - lo) * row * esz) + 64;
        }
        return 4096 + conf_est + total;
    }
    return 4096 + conf_est + ZSTD_compressBound(num * esz);
}
}  // namespace

extern "C" {

const char *sz3b_last_error(void) { return t_last_error.c_str(); }
const char *sz3b_version(void) { return "3.3.2"; }

int sz3b_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void sz3b_set_host_threads(int n) { set_host_threads(n); }
int sz3b_get_host_threads(void) { return host_threads(); }
void sz3b_set_host_wait(int mode) { host_wait_mode().store(mode != 0); }
void sz3b_set_caller_stream(void *cuda_stream, int enable) {
    t_caller_stream = static_cast<cudaStream_t>(cuda_stream);
    t_caller_stream_set = enable != 0;
}
void sz3b_set_device_fanout(int n) { set_device_fanout(n); }
int sz3b_get_device_fanout(void) { return device_fanout(); }
void sz3b_set_lossless_policy(int policy) { set_lossless_policy(policy); }
int sz3b_get_lossless_policy(void) { return lossless_policy(); }
void sz3b_set_frame_decoder(int mode) { set_frame_decoder(mode); }
int sz3b_get_frame_decoder(void) { return frame_decoder(); }

int sz3b_config_init(sz3b_config *c, int ndims, const size_t *dims) {
    return guarded([&] {
        if (!c || !dims || ndims < 1) fail(SZ3B_E_INVALID_ARGUMENT, "bad arguments");
        memset(c, 0, sizeof(*c));
        c->cmprAlgo = SZ3B_ALGO_INTERP_LORENZO;
        c->errorBoundMode = SZ3B_EB_ABS;
        c->absErrorBound = 1e-3;
        c->quantbinCnt = 65536;
        c->lorenzo = 1;
        c->regression = 1;
        c->interpAlgo = SZ3B_INTERP_CUBIC;
        c->interpAnchorStride = -1;
        c->interpAlpha = 1.25;
        c->interpBeta = 2.0;
        c->dataType = SZ3B_FLOAT;
        int kept = 0;
        for (int i = 0; i < ndims; i++)
            if (dims[i] > 1) kept++;
        if (kept > 4) fail(SZ3B_E_INVALID_ARGUMENT, "Data dimension higher than 4 is not supported.");
        std::vector<uint64_t> d(dims, dims + ndims);
        config_set_dims(*c, ndims, d.data());
    });
}

size_t sz3b_config_save(const sz3b_config *c, unsigned char *out) { return config_save(*c, out); }

int sz3b_config_load(sz3b_config *c, const unsigned char *in, size_t len) {
    return guarded([&] {
        if (!config_load(*c, in, len)) fail(SZ3B_E_INVALID_ARGUMENT, "malformed Config blob");
    });
}

size_t sz3b_compress_bound(int dtype, const sz3b_config *c) { return size_bound(dtype, *c); }

int sz3b_compress(int dtype, const sz3b_config *c, const void *data, int data_loc, char *cmp, size_t cmp_cap,
                  size_t *cmp_size, sz3b_config *conf_out) {
    return guarded([&] {
        check_dtype_any(dtype);
        check_conf(c);
        if (!data || !cmp || !cmp_size) fail(SZ3B_E_INVALID_ARGUMENT, "null buffer");
        sz3b_config conf = *c;
        if (cmp_cap < size_bound(dtype, conf)) fail(SZ3B_E_INVALID_ARGUMENT, "compressed buffer not large enough");
        uint8_t *p = reinterpret_cast<uint8_t *>(cmp);
        put<uint32_t>(p, kMagic);
        put<uint32_t>(p, kDataVer);
        uint8_t *size_pos = p;
        p += 8;
        uint8_t blob[256];
        const size_t conf_est = config_save(conf, blob);
        const size_t cap = cmp_cap - 16 - conf_est * 2;
        WorkspaceLease ws;
        after_caller(*ws, data_loc);
        uint64_t payload = 0;
        try {
            payload = by_dtype(dtype, [&](auto *tag) -> uint64_t {
                using T = SZ3B_TAG_T(tag);
                return compress_any<T>(*ws, conf, static_cast<const T *>(data), data_loc, p, cap);
            });
        } catch (...) {
            finish_profile(*ws);
            throw;
        }
        finish_profile(*ws);
        put<uint64_t>(size_pos, payload);
        p += payload;
        const size_t conf_size = config_save(conf, p);
        *cmp_size = 16 + payload + conf_size;
        if (conf_out) *conf_out = conf;
    });
}

int sz3b_peek_config(const char *cmp, size_t cmp_size, sz3b_config *conf_out) {
    return guarded([&] {
        if (!cmp || cmp_size < 17 || !conf_out) fail(SZ3B_E_INVALID_ARGUMENT, "bad arguments");
        const uint8_t *p = reinterpret_cast<const uint8_t *>(cmp);
        if (get<uint32_t>(p) != kMagic)
            fail(SZ3B_E_INVALID_ARGUMENT, "magic number mismatch, the input data is not compressed by SZ3");
        if (get<uint32_t>(p) != kDataVer) fail(SZ3B_E_INVALID_ARGUMENT, "Please use the matching SZ3 version to decompress the data");
        const uint64_t payload = get<uint64_t>(p);
        if (payload > cmp_size - 16) fail(SZ3B_E_INVALID_ARGUMENT, "truncated stream");
        sz3b_config c;
        memset(&c, 0, sizeof(c));
        c.quantbinCnt = 65536;
        c.interpAlgo = SZ3B_INTERP_CUBIC;
        c.interpAnchorStride = -1;
        c.interpAlpha = 1.25;
        c.interpBeta = 2.0;
        if (!config_load(c, p + payload, cmp_size - 16 - payload)) fail(SZ3B_E_INVALID_ARGUMENT, "malformed Config blob");
        *conf_out = c;
    });
}

int sz3b_decompress(int dtype, const char *cmp, size_t cmp_size, void *out, int out_loc, sz3b_config *conf_out) {
    sz3b_config conf;
    int rc = sz3b_peek_config(cmp, cmp_size, &conf);
    if (rc != SZ3B_OK) return rc;
    return guarded([&] {
        check_dtype_any(dtype);
        if (!out) fail(SZ3B_E_INVALID_ARGUMENT, "null output buffer");
        const uint8_t *p = reinterpret_cast<const uint8_t *>(cmp) + 8;
        const uint64_t payload = get<uint64_t>(p);
        WorkspaceLease ws;
        after_caller(*ws, out_loc);
        try {
            by_dtype(dtype, [&](auto *tag) {
                using T = SZ3B_TAG_T(tag);
                decompress_any<T>(*ws, conf, p, payload, static_cast<T *>(out), out_loc);
                return 0;
            });
        } catch (...) {
            finish_profile(*ws);
            throw;
        }
        finish_profile(*ws);
        if (conf_out) *conf_out = conf;
    });
}

int sz3b_abs_error_bound(int dtype, const sz3b_config *c, const void *data, int data_loc, double *abs_eb) {
    return guarded([&] {
        check_dtype_any(dtype);
        check_conf(c);
        WorkspaceLease ws;
        after_caller(*ws, data_loc);
        *abs_eb = by_dtype(dtype, [&](auto *tag) -> double {
            using T = SZ3B_TAG_T(tag);
            return abs_eb_stage<T>(*ws, *c, static_cast<const T *>(data), data_loc);
        });
    });
}

int sz3b_interp_decompose(int dtype, const sz3b_config *c, double abs_eb, const void *data, int data_loc, int schedule,
                          int32_t *quant_out, unsigned char *blob_out, size_t blob_cap, size_t *blob_len) {
    return guarded([&] {
        check_dtype(dtype);
        check_conf(c);
        WorkspaceLease ws;
        after_caller(*ws, data_loc);
        std::vector<uint8_t> blob;
        try {
            if (dtype == SZ3B_FLOAT)
                interp_decompose_stage<float>(*ws, *c, abs_eb, static_cast<const float *>(data), data_loc, schedule,
                                              quant_out, blob);
            else
                interp_decompose_stage<double>(*ws, *c, abs_eb, static_cast<const double *>(data), data_loc, schedule,
                                               quant_out, blob);
        } catch (...) {
            finish_profile(*ws);
            throw;
        }
        finish_profile(*ws);
        if (blob_len) *blob_len = blob.size();
        if (blob_out) {
            if (blob.size() > blob_cap) fail(SZ3B_E_INVALID_ARGUMENT, "blob buffer too small");
            memcpy(blob_out, blob.data(), blob.size());
        }
    });
}

int sz3b_blockwise_decompose(int dtype, const sz3b_config *c, double abs_eb, const void *data, int data_loc,
                             int32_t *quant_out, unsigned char *blob_out, size_t blob_cap, size_t *blob_len) {
    return guarded([&] {
        check_dtype(dtype);
        check_conf(c);
        WorkspaceLease ws;
        after_caller(*ws, data_loc);
        std::vector<uint8_t> blob;
        try {
            if (dtype == SZ3B_FLOAT)
                blockwise_decompose_stage<float>(*ws, *c, abs_eb, static_cast<const float *>(data), data_loc,
                                                 quant_out, blob);
            else
                blockwise_decompose_stage<double>(*ws, *c, abs_eb, static_cast<const double *>(data), data_loc,
                                                  quant_out, blob);
        } catch (...) {
            finish_profile(*ws);
            throw;
        }
        finish_profile(*ws);
        if (blob_len) *blob_len = blob.size();
        if (blob_out) {
            if (blob.size() > blob_cap) fail(SZ3B_E_INVALID_ARGUMENT, "blob buffer too small");
            memcpy(blob_out, blob.data(), blob.size());
        }
    });
}

int sz3b_huffman_encode(const int32_t *q, size_t n, int q_loc, unsigned char *out, size_t out_cap, size_t *out_len,
                        size_t *tree_len) {
    return guarded([&] {
        WorkspaceLease ws;
        after_caller(*ws, q_loc);
        std::vector<uint8_t> buf;
        huffman_encode_stage(*ws, q, n, q_loc, buf, tree_len);
        finish_profile(*ws);
        if (out_len) *out_len = buf.size();
        if (buf.size() > out_cap) fail(SZ3B_E_INVALID_ARGUMENT, "output buffer too small");
        memcpy(out, buf.data(), buf.size());
    });
}

int sz3b_huffman_decode(const unsigned char *in, size_t in_len, size_t tree_len, size_t n, int32_t *out) {
    return guarded([&] {
        WorkspaceLease ws;
        huffman_decode_stage(*ws, in, in_len, tree_len, n, out);
        finish_profile(*ws);
    });
}

int sz3b_lossless_compress(const unsigned char *src, size_t src_len, int src_loc, unsigned char *out, size_t out_cap,
                           size_t *out_len) {
    return guarded([&] {
        WorkspaceLease ws;
        after_caller(*ws, src_loc);
        const size_t n = lossless_gpu_stage(*ws, src, src_len, src_loc, out, out_cap);
        finish_profile(*ws);
        if (out_len) *out_len = n;
    });
}

int sz3b_tune(int dtype, sz3b_config *c, const void *data, int data_loc) {
    return guarded([&] {
        check_dtype(dtype);
        check_conf(c);
        WorkspaceLease ws;
        after_caller(*ws, data_loc);
        try {
            if (dtype == SZ3B_FLOAT)
                tune_stage<float>(*ws, *c, static_cast<const float *>(data), data_loc);
            else
                tune_stage<double>(*ws, *c, static_cast<const double *>(data), data_loc);
        } catch (...) {
            finish_profile(*ws);
            throw;
        }
        finish_profile(*ws);
    });
}

int sz3b_minmax(int dtype, const void *data, int data_loc, size_t num, double *min_out, double *max_out) {
    return guarded([&] {
        check_dtype_any(dtype);
        WorkspaceLease ws;
        after_caller(*ws, data_loc);
        by_dtype(dtype, [&](auto *tag) {
            using T = SZ3B_TAG_T(tag);
            minmax_stage<T>(*ws, static_cast<const T *>(data), data_loc, num, min_out, max_out);
            return 0;
        });
    });
}

namespace {
// slab bounds and per-slab Config exactly as SZImplOMP.hpp:46-72
sz3b_config slab_config(const sz3b_config *c, int rank, int nslabs, double range) {
    if (nslabs < 1 || rank < 0 || rank >= nslabs || static_cast<uint64_t>(nslabs) > c->dims[0])
        fail(SZ3B_E_INVALID_ARGUMENT, "bad slab index (nslabs must not exceed dims[0])");
    sz3b_config sc = *c;
    sc.openmp = 1;
    int lo = static_cast<int>(static_cast<uint64_t>(rank) * c->dims[0] / nslabs);
    int hi = static_cast<int>(static_cast<uint64_t>(rank + 1) * c->dims[0] / nslabs);
    uint64_t d[4];
    for (int i = 0; i < c->N; i++) d[i] = c->dims[i];
    d[0] = hi - lo;
    // range == 0 is a legitimate value (a constant field): the reference resolves the bound to 0 and stores the slab
    // losslessly (SZImplOMP.hpp:57-68, Statistic.hpp:32-51); only a missing range (negative / NaN) is an error
    if (c->errorBoundMode != SZ3B_EB_ABS && c->errorBoundMode != SZ3B_EB_L2NORM && !(range >= 0))
        fail(SZ3B_E_INVALID_ARGUMENT, "non-ABS error bound needs the global value range (max - min >= 0)");
    if (sc.errorBoundMode == SZ3B_EB_L2NORM) {
        // resolved against the WHOLE array's element count, as the shared conf is in the reference (:61-65)
        sc.absErrorBound = sqrt(3.0 / config_num(*c)) * sc.l2normErrorBound;
        sc.errorBoundMode = SZ3B_EB_ABS;
    }
    config_set_dims(sc, c->N, d);
    return sc;
}
}  // namespace

int sz3b_compress_slab(int dtype, const sz3b_config *c, int rank, int nslabs, const void *slab, int data_loc,
                       double range, char *payload, size_t payload_cap, size_t *payload_size,
                       unsigned char *conf_blob, size_t *conf_blob_size) {
    return guarded([&] {
        check_dtype_any(dtype);
        check_conf(c);
        sz3b_config sc = slab_config(c, rank, nslabs, range);
        WorkspaceLease ws;
        after_caller(*ws, data_loc);
        size_t sz = 0;
        try {
            sz = by_dtype(dtype, [&](auto *tag) -> size_t {
                using T = SZ3B_TAG_T(tag);
                return compress_slab<T>(*ws, sc, static_cast<const T *>(slab), data_loc, range, reinterpret_cast<uint8_t *>(payload),
                                        payload_cap);
            });
        } catch (...) {
            finish_profile(*ws);
            throw;
        }
        finish_profile(*ws);
        *payload_size = sz;
        *conf_blob_size = config_save(sc, conf_blob);
    });
}

int sz3b_compress_slab_placed(int dtype, const sz3b_config *c, int rank, int nslabs, const void *slab, int data_loc,
                              double range, sz3b_place_fn place, void *user, size_t *payload_size,
                              unsigned char *conf_blob, size_t *conf_blob_size) {
    return guarded([&] {
        check_dtype_any(dtype);
        check_conf(c);
        if (!place) fail(SZ3B_E_INVALID_ARGUMENT, "null placement callback");
        sz3b_config sc = slab_config(c, rank, nslabs, range);
        WorkspaceLease ws;
        after_caller(*ws, data_loc);
        size_t sz = 0;
        try {
            sz = by_dtype(dtype, [&](auto *tag) -> size_t {
                using T = SZ3B_TAG_T(tag);
                return compress_slab_placed<T>(*ws, sc, static_cast<const T *>(slab), data_loc, range, place, user);
            });
        } catch (...) {
            finish_profile(*ws);
            throw;
        }
        finish_profile(*ws);
        *payload_size = sz;
        *conf_blob_size = config_save(sc, conf_blob);
    });
}

size_t sz3b_slab_conf_blob_size(const sz3b_config *c, int rank, int nslabs) {
    size_t n = 0;
    guarded([&] {
        check_conf(c);
        sz3b_config sc = slab_config(c, rank, nslabs, 1.0);
        sc.errorBoundMode = SZ3B_EB_ABS;   // what every slab carries once its bound is resolved
        uint8_t blob[256];
        n = config_save(sc, blob);
    });
    return n;
}

size_t sz3b_omp_header_size(int nslabs, const size_t *conf_blob_sizes) {
    size_t s = 16 + sizeof(int32_t) + static_cast<size_t>(nslabs) * sizeof(uint64_t);
    for (int i = 0; i < nslabs; i++) s += conf_blob_sizes[i];
    return s;
}

int sz3b_omp_assemble(int dtype, const sz3b_config *c, int nslabs, const unsigned char *const *conf_blobs,
                      const size_t *conf_blob_sizes, const size_t *payload_sizes, const char *const *payloads,
                      char *cmp, size_t cmp_cap, size_t *cmp_size) {
    return guarded([&] {
        check_dtype_any(dtype);
        check_conf(c);
        size_t need = sz3b_omp_header_size(nslabs, conf_blob_sizes) + 256;
        for (int i = 0; i < nslabs; i++) need += payload_sizes[i];
        if (cmp_cap < need) fail(SZ3B_E_INVALID_ARGUMENT, "compressed buffer not large enough");
        uint8_t *p = reinterpret_cast<uint8_t *>(cmp);
        put<uint32_t>(p, kMagic);
        put<uint32_t>(p, kDataVer);
        uint8_t *size_pos = p;
        p += 8;
        uint8_t *body = p;
        put<int32_t>(p, nslabs);
        for (int i = 0; i < nslabs; i++) {
            memcpy(p, conf_blobs[i], conf_blob_sizes[i]);
            p += conf_blob_sizes[i];
        }
        for (int i = 0; i < nslabs; i++) put<uint64_t>(p, payload_sizes[i]);
        for (int i = 0; i < nslabs; i++) {
            if (payloads && payloads[i]) memcpy(p, payloads[i], payload_sizes[i]);
            p += payload_sizes[i];
        }
        put<uint64_t>(size_pos, static_cast<uint64_t>(p - body));
        // outer Config: the caller's, marked openmp.  In the reference calAbsErrorBound rewrites the shared conf before
        // the slabs copy it (SZImplOMP.hpp:57-72), so the trailing blob carries mode ABS and the resolved bound: take
        // both from slab 0's blob (every slab holds the same pair).
        sz3b_config oc = *c;
        oc.openmp = 1;
        if (oc.errorBoundMode != SZ3B_EB_ABS && nslabs > 0) {
            sz3b_config s0;
            if (config_load(s0, conf_blobs[0], conf_blob_sizes[0])) {
                oc.errorBoundMode = s0.errorBoundMode;
                oc.absErrorBound = s0.absErrorBound;
            }
        }
        p += config_save(oc, p);
        *cmp_size = static_cast<size_t>(p - reinterpret_cast<uint8_t *>(cmp));
    });
}

void sz3b_last_transfer(size_t *h2d_bytes, size_t *d2h_bytes) {
    if (h2d_bytes) *h2d_bytes = t_h2d;
    if (d2h_bytes) *d2h_bytes = t_d2h;
}

int sz3b_last_profile(const char **names, double *ms, int *launches, int cap) {
    int n = static_cast<int>(t_profile.size());
    for (int i = 0; i < n && i < cap; i++) {
        if (names) names[i] = t_profile[i].name.c_str();
        if (ms) ms[i] = t_profile[i].ms;
        if (launches) launches[i] = t_profile[i].launches;
    }
    return n;
}

}  // extern "C"
