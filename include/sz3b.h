/* include/sz3b.h -- C ABI of libsz3b200.so, the B200-native implementation of SZ3's predict -> quantize -> encode
 * hot path.  Plain pointers and sizes only; every entry point names the reference interface it stands in for
 * (paths relative to the SZ3 v3.3.2 source tree).
 *
 * The drop-in C++ headers (sz3_b200/include/SZ3/api/sz.hpp, SZ3/utils/Config.hpp) and the sz3c shim
 * (sz3_b200/sz3c) are thin forwarders onto these functions; INTEGRATION.md shows the bindings.
 *
 * All functions return 0 on success or a negative SZ3B_E_* code; sz3b_last_error() gives the message for the
 * calling thread.  `loc` arguments say where a buffer lives: SZ3B_HOST (0) or SZ3B_DEVICE (1, a CUDA device pointer
 * on the current device).  Compressed streams always live in host memory (the final zstd pass runs on the host).
 */
#ifndef SZ3B_H
#define SZ3B_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SZ3B_HOST 0
#define SZ3B_DEVICE 1

#define SZ3B_FLOAT 0  /* SZ_FLOAT,  include/SZ3/def.hpp */
#define SZ3B_DOUBLE 1 /* SZ_DOUBLE */
#define SZ3B_INT32 7  /* SZ_INT32: whole-array entry points only (compress, decompress, bounds, slabs), */
#define SZ3B_INT64 9  /* SZ_INT64  as tools/sz3/sz3.cpp:458-461 uses them                               */

#define SZ3B_OK 0
#define SZ3B_E_INVALID_ARGUMENT (-1) /* the reference throws std::invalid_argument */
#define SZ3B_E_RUNTIME (-2)          /* the reference throws std::runtime_error */
#define SZ3B_E_CUDA (-3)             /* CUDA runtime failure / no device: there is no CPU fallback */
#define SZ3B_E_UNSUPPORTED (-4)      /* configuration outside the GPU path built so far */

/* enum EB / ALGO / INTERP_ALGO, include/SZ3/utils/Config.hpp:54,68,77 */
enum { SZ3B_EB_ABS, SZ3B_EB_REL, SZ3B_EB_PSNR, SZ3B_EB_L2NORM, SZ3B_EB_ABS_AND_REL, SZ3B_EB_ABS_OR_REL };
enum { SZ3B_ALGO_LORENZO_REG, SZ3B_ALGO_INTERP_LORENZO, SZ3B_ALGO_INTERP, SZ3B_ALGO_NOPRED, SZ3B_ALGO_LOSSLESS };
enum { SZ3B_INTERP_LINEAR, SZ3B_INTERP_CUBIC };

/* POD mirror of SZ3::Config (include/SZ3/utils/Config.hpp:441-478). */
typedef struct sz3b_config {
    int32_t N;
    uint64_t dims[4];
    int32_t cmprAlgo;
    int32_t errorBoundMode;
    double absErrorBound;
    double relErrorBound;
    double psnrErrorBound;
    double l2normErrorBound;
    int32_t openmp;      /* 0 = single stream; >0 = emit the OpenMP slab container (SZImplOMP.hpp) with this many slabs */
    int32_t quantbinCnt;
    int32_t blockSize;
    int32_t lorenzo;
    int32_t lorenzo2;
    int32_t regression;
    int32_t regression2;
    int32_t interpAlgo;
    int32_t interpDirection;
    int32_t interpAnchorStride;
    double interpAlpha;
    double interpBeta;
    int32_t dataType;
    int32_t predDim;
} sz3b_config;

/* SZ3::Config(dims...) + setDims (Config.hpp:146-177): drops size-1 dims, sets defaults. */
int sz3b_config_init(sz3b_config *c, int ndims, const size_t *dims);

/* Config::save / Config::load (Config.hpp:312-413): the blob appended to every stream. */
size_t sz3b_config_save(const sz3b_config *c, unsigned char *out);
int sz3b_config_load(sz3b_config *c, const unsigned char *in, size_t len);

/* SZ_compress_size_bound<T> (include/SZ3/api/impl/SZImpl.hpp:34-44). */
size_t sz3b_compress_bound(int dtype, const sz3b_config *c);

/* SZ_compress<T>(conf, data, cmpData, cmpCap) (include/SZ3/api/sz.hpp:43-82).  `conf_out`, if not NULL, receives the
 * configuration actually used (tuned interpolation parameters, resolved absolute bound) -- the reference keeps that
 * private to the call. */
int sz3b_compress(int dtype, const sz3b_config *c, const void *data, int data_loc, char *cmp, size_t cmp_cap,
                  size_t *cmp_size, sz3b_config *conf_out);

/* SZ_decompress<T>(conf, cmpData, cmpSize, decData) (include/SZ3/api/sz.hpp:117-157).  `out` must hold conf.num
 * elements (query with sz3b_peek_config). */
int sz3b_decompress(int dtype, const char *cmp, size_t cmp_size, void *out, int out_loc, sz3b_config *conf_out);

/* Header + trailing Config of a stream (sz.hpp:119-141) without decompressing. */
int sz3b_peek_config(const char *cmp, size_t cmp_size, sz3b_config *conf_out);

/* calAbsErrorBound (include/SZ3/utils/Statistic.hpp:32-56); the min/max scan runs on the GPU. */
int sz3b_abs_error_bound(int dtype, const sz3b_config *c, const void *data, int data_loc, double *abs_eb);

/* ---- stage-level entry points (parity tests, bench) --------------------------------------------------------------- */

/* InterpolationDecomposition::compress + save (decomposition/InterpolationDecomposition.hpp:79-159).
 * quant_out: conf.num int32 indices in reference traversal order (host).  blob_out: what save() writes (dims,
 * blocksize, interp id, direction, anchor stride, alpha, beta, quantizer incl. unpredictable values).
 * schedule: 0 = automatic (N == 3: box schedule where it applies, line-walker tile kernel otherwise; N == 4: row-mapped
 * per-pass kernels), 1 = force the generic per-pass kernels, 4 = force the line-walker tile kernel (N == 3), 5 = force
 * the row-mapped per-pass kernels (N >= 3), 6 = box schedule or fail; 2 and 3 (the first two tile kernels) were
 * retired in round 2. */
int sz3b_interp_decompose(int dtype, const sz3b_config *c, double abs_eb, const void *data, int data_loc, int schedule,
                          int32_t *quant_out, unsigned char *blob_out, size_t blob_cap, size_t *blob_len);

/* BlockwiseDecomposition::compress + save (decomposition/BlockwiseDecomposition.hpp:28-46,69-73) with the predictor
 * stack of make_compressor_lorenzo_regression (api/impl/SZAlgoLorenzoReg.hpp:22-64). */
int sz3b_blockwise_decompose(int dtype, const sz3b_config *c, double abs_eb, const void *data, int data_loc,
                             int32_t *quant_out, unsigned char *blob_out, size_t blob_cap, size_t *blob_len);

/* HuffmanEncoder<int>::preprocess_encode + save + encode (encoder/HuffmanEncoder.hpp:96-218): histogram and bit
 * packing on the GPU, tree on the host.  out = tree blob | size_t outSize | bits. */
int sz3b_huffman_encode(const int32_t *q, size_t n, int q_loc, unsigned char *out, size_t out_cap, size_t *out_len,
                        size_t *tree_len);

/* HuffmanEncoder<int>::load + decode (encoder/HuffmanEncoder.hpp:225-279) on the GPU (self-synchronising decoder):
 * in = tree blob | size_t outSize | bits, tree_len as reported by sz3b_huffman_encode; out = n host int32. */
int sz3b_huffman_decode(const unsigned char *in, size_t in_len, size_t tree_len, size_t n, int32_t *out);

/* Lossless_zstd::compress (lossless/Lossless_zstd.hpp:29-37) by the GPU lossless stage of policy 2: src (host or
 * device, src_loc) -> out (host) = size_t srcLen | standard zstd frames of Huffman-only literal blocks. */
int sz3b_lossless_compress(const unsigned char *src, size_t src_len, int src_loc, unsigned char *out, size_t out_cap,
                           size_t *out_len);

/* The auto-tuner inside SZ_compress_Interp_lorenzo (api/impl/SZAlgoInterp.hpp:122-286): fills in cmprAlgo,
 * interpAlgo, interpDirection, interpAlpha, interpBeta (and absErrorBound) exactly as the reference would choose. */
int sz3b_tune(int dtype, sz3b_config *c, const void *data, int data_loc);

/* ---- multi-GPU slab container (api/impl/SZImplOMP.hpp:16-117) -------------------------------------------------------
 * Rank r of `nslabs` compresses rows [r*d0/nslabs, (r+1)*d0/nslabs) of the outermost dimension as an independent
 * stream.  `slab` points at that slab only.  `range` is the global max-min (only read when the error-bound mode is
 * not ABS; ranks obtain it with an all-reduce of sz3b_minmax results; 0 -- a constant field -- is valid and leads to
 * the lossless path as in the reference, a negative value or NaN means "not supplied" and is an error).  The rank gets back its payload and its Config
 * blob; rank 0 concatenates them with sz3b_omp_assemble after a gather of the sizes. */
int sz3b_minmax(int dtype, const void *data, int data_loc, size_t num, double *min_out, double *max_out);
int sz3b_compress_slab(int dtype, const sz3b_config *c, int rank, int nslabs, const void *slab, int data_loc,
                       double range, char *payload, size_t payload_cap, size_t *payload_size,
                       unsigned char *conf_blob, size_t *conf_blob_size);
/* Same, with the payload delivered where `place(user, size)` says.  The callback runs on the calling thread as soon as
 * the size of the slab's payload is known -- with the GPU lossless stage the compressed frames are still on the device
 * then -- and returns the destination (host memory, pinned for full PCIe rate): the place for a rank to exchange sizes
 * with the other ranks (the all-gather of SZImplOMP.hpp:93-99) and to derive its offset in a shared container, so that
 * the frames cross PCIe once, to their final position.  The blob's size does not depend on the outcome
 * (sz3b_slab_conf_blob_size), so the container header can be laid out before any slab is finished. */
typedef void *(*sz3b_place_fn)(void *user, size_t payload_size);
int sz3b_compress_slab_placed(int dtype, const sz3b_config *c, int rank, int nslabs, const void *slab, int data_loc,
                              double range, sz3b_place_fn place, void *user, size_t *payload_size,
                              unsigned char *conf_blob, size_t *conf_blob_size);
size_t sz3b_slab_conf_blob_size(const sz3b_config *c, int rank, int nslabs);
size_t sz3b_omp_header_size(int nslabs, const size_t *conf_blob_sizes);
int sz3b_omp_assemble(int dtype, const sz3b_config *c, int nslabs, const unsigned char *const *conf_blobs,
                      const size_t *conf_blob_sizes, const size_t *payload_sizes, const char *const *payloads,
                      char *cmp, size_t cmp_cap, size_t *cmp_size);

/* ---- diagnostics -------------------------------------------------------------------------------------------------- */
const char *sz3b_last_error(void);
const char *sz3b_version(void);           /* "3.3.2" data format */
int sz3b_device_count(void);

/* Per-stage device times (CUDA events on the library's stream) of the calling thread's last compress call.
 * names/ms arrays of capacity `cap`; returns the number of stages recorded.  Launch counts in `launches`. */
int sz3b_last_profile(const char **names, double *ms, int *launches, int cap);
/* Bytes the calling thread's last call moved host->device and device->host (cudaMemcpyAsync on the call's stream). */
void sz3b_last_transfer(size_t *h2d_bytes, size_t *d2h_bytes);
/* Host threads the library may use: zstd workers of the host tail and concurrent tuner trials (0 = hardware
 * concurrency).  With one rank per GPU on a shared host, give each rank its share of the cores. */
void sz3b_set_host_threads(int n);
int sz3b_get_host_threads(void);   /* the value in force (after the defaults above) */
/* Stream contract for SZ3B_DEVICE buffers.  The library works on its own non-blocking streams and returns when its work
 * is complete (results are valid on return).  What the caller queued BEFORE the call on a stream of its own must be
 * ordered explicitly: sz3b_set_caller_stream(stream, 1) (per host thread) makes every following call wait on an event
 * recorded on that stream before it touches a device buffer; sz3b_set_caller_stream(NULL, 0) restores the default,
 * which synchronises the legacy default stream (enough for callers that never create streams).  Work queued on other
 * non-blocking streams must be synchronised by the caller. */
void sz3b_set_caller_stream(void *cuda_stream, int enable);
/* Devices one sz3b_compress call with conf.openmp > 0 spreads its slabs over (slab t on device t mod n, one host thread
 * per device inside the call; api/impl/SZImplOMP.hpp:16-117 gives every slab to an OpenMP thread instead).
 * 0 = every visible device; default: every visible device, or 1 when a one-process-per-GPU launcher is detected
 * (LOCAL_WORLD_SIZE > 1), where each rank keeps to the device it selected. */
void sz3b_set_device_fanout(int n);
int sz3b_get_device_fanout(void);
/* How host threads wait for the device: 0 = the driver's wait (spins; lowest latency when cores are plentiful),
 * 1 = poll and yield the core between polls (for hosts with more waiting threads than cores).  No reference
 * counterpart (the reference has no device); initial value from the environment variable SZ3B_HOST_WAIT. */
void sz3b_set_host_wait(int mode);
/* Lossless stage over the packed (Huffman-coded) stream, lossless/Lossless_zstd.hpp:29-37.
 *   0 = every chunk through zstd level 3 on the host (the reference's call, frame by frame);
 *   1 = adaptive host zstd: every 8th 1-MiB chunk is compressed as a probe; if zstd gains < 1 % on the probes, the
 *       other chunks are stored as raw zstd frames;
 *   2 = (default) streams of 4 MiB and more are coded on the GPU into standard zstd frames whose blocks hold
 *       Huffman-only literals with one table per 128 KiB block (sz3_b200/csrc/zhuf.cuh): zstd finds no matches in an
 *       entropy-coded stream, its gain there IS the literal coding; only the compressed frames cross PCIe.
 *       Shorter streams take the host zstd call of policy 0 and stay byte-identical to the reference's.
 *   In every case the result is a concatenation of standard zstd frames that the unmodified reference decoder reads. */
void sz3b_set_lossless_policy(int policy);
int sz3b_get_lossless_policy(void);
/* Decompression side of the same stage, lossless/Lossless_zstd.hpp:39-45 (ZSTD_decompress of the whole payload).
 *   0 = every frame through libzstd on the host pool (any zstd stream);
 *   1 = frames of the shape policy 2 writes (raw blocks and blocks of Huffman-only literals in four streams, no
 *       sequences) are decoded on the GPU, one thread per stream (sz3_b200/csrc/zhuf_dec.cuh); the host decodes only
 *       the frames that hold the head of the stream (headers, stored values, Huffman tree).  Any other payload, and
 *       any block the GPU decoder does not take, goes through libzstd as with 0.  Interpolation streams only.
 * Initial value: SZ3B_FRAME_DECODER from the environment, else the library default. */
void sz3b_set_frame_decoder(int mode);
int sz3b_get_frame_decoder(void);

#ifdef __cplusplus
}
#endif
#endif
