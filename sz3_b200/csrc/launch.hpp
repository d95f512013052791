// sz3_b200/csrc/launch.hpp -- host-callable launchers of every CUDA kernel in the library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "core.cuh"

namespace sz3b {

// Function attributes (dynamic shared memory beyond 48 KB) are per device.  `mask` is a static of the call site (bit d =
// done on device d): once_per_device(mask, fn) runs fn until one run has COMPLETED on the current device -- concurrent
// first callers (tuner trials on pool threads, the per-device threads of a container call) may all run it, which is
// harmless; none of them launches before its own run has finished.
template <class F>
inline void once_per_device(std::atomic<unsigned long long> &mask, F &&fn) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask.load(std::memory_order_acquire) & bit) return;
    fn();
    mask.fetch_or(bit, std::memory_order_release);
}

template <class T, class QT>
struct InterpArgs;

// interp_kernels.cu
template <class T, class QT>
void interp_launch_anchors(const InterpArgs<T, QT> &A, uint32_t anchor_stride, uint64_t n_anchor, uint32_t nbatch,
                           cudaStream_t st);
template <class T, class QT>
void interp_launch_ltiles(const InterpArgs<T, QT> &A, uint64_t ntiles, uint32_t nbatch, cudaStream_t st);
template <class T, class QT>
void interp_launch_pass(const InterpArgs<T, QT> &A, int p, uint32_t nbatch, cudaStream_t st);
// row-mapped per-pass kernel (N >= 3); returns false when the shape does not fit its table / grid (caller falls back)
template <class T, class QT>
bool interp_launch_lean(const InterpArgs<T, QT> &A, int p, uint32_t nbatch, bool write_work, bool recover, const T *unpred_in,
                        cudaStream_t st);

// interp_box.cu: box schedule of the finest level (float data, 16-bit indices, pass order z, y, x)
struct BoxSrc;
bool interp_box_applicable(const InterpArgs<float, uint16_t> &A);
bool interp_launch_box_recover_x(const InterpArgs<float, uint16_t> &A, float *out, uint64_t ntiles, cudaStream_t st);
bool interp_launch_box(const InterpArgs<float, uint16_t> &A, const BoxSrc &S, const uint32_t sdims[3], uint64_t ntiles,
                       cudaStream_t st);
void interp_launch_compact(const float *src, const uint32_t dims[3], const uint64_t stride[3], uint32_t s0, int n,
                           float *const dst[3], cudaStream_t st);

// encode_kernels.cu
template <class QT>
void launch_histogram(const QT *q, uint64_t n, int sym_min, int nbins, int center, unsigned long long *ghist,
                      cudaStream_t st);
// histogram of a stretch of 16-bit indices around `radius` (the streams the box schedule writes)
void launch_hist_u16(const uint16_t *q, uint64_t n, int radius, int nbins, unsigned long long *ghist, cudaStream_t st);
template <class QT>
void launch_minmax_int(const QT *q, uint64_t n, int *mm, cudaStream_t st);
uint64_t pack_num_chunks(uint64_t n);
template <class QT, class T>
void launch_pack(const QT *q, uint64_t n, int sym_min, int zero_sym, const uint8_t *len,
                 const unsigned long long *code, unsigned nstates, int center_state, unsigned *chunk_bits,
                 unsigned *chunk_zeros, unsigned long long *bit_off, unsigned long long *zero_off, unsigned *out_words,
                 const T *unpred_tmp, T *unpred_out, cudaStream_t st, cudaEvent_t after_scan,
                 unsigned long long *scan_scratch = nullptr);

// blockwise.cu
struct BlockShape;
struct QuantParams;
template <class T>
void launch_reg_fit(const T *data, const BlockShape &bs, T *c_fit, uint8_t *valid, cudaStream_t st);
template <class T>
void launch_reg_chain(const T *c_fit, const uint8_t *sel, uint64_t nblocks, int N, const QuantParams &q_liner,
                      const QuantParams &q_indep, int32_t *coef_q, T *c_rec, unsigned long long *counters,
                      unsigned long long *unpred_pos, T *unpred_val, cudaStream_t st, const T *init = nullptr,
                      unsigned long long pos_base = 0);   // init / pos_base: dense chain continued from an earlier launch
template <class T, class QT>
const char *launch_reg_predict(const T *data, const BlockShape &bs, const T *c_rec, const QuantParams &qp, QT *q,
                               T *unpred_tmp, unsigned long long *hist, cudaStream_t st);   // nullptr or an error text

// lorenzo.cu
template <class T, class QT>
struct BwArgs;
template <class T>
void launch_bw_pad(const T *data, const BlockShape &bs, const uint64_t *pstride, T *W, uint64_t b_lo, uint64_t b_hi,
                   cudaStream_t st, uint32_t nbatch = 1, uint64_t w_bstride = 0);
template <class T, class QT>
const char *launch_bw_serial(const BwArgs<T, QT> &A, uint64_t b_lo, uint64_t b_hi, const T *chain_init,
                             unsigned long long nsel0, unsigned long long *nsel_out, cudaStream_t st);
template <class T, class QT>
const char *launch_bw_fronts(const BwArgs<T, QT> &A, cudaStream_t st, int *launches);   // nullptr or an error text
template <class T>
void launch_bw_spec_coef(const T *c_fit, const uint8_t *valid, uint64_t nblocks, int N, const QuantParams &q_liner,
                         const QuantParams &q_indep, T *c_spec, cudaStream_t st);
void launch_bw_rank(const uint8_t *sel, uint64_t b_lo, uint64_t b_hi, int reg_sid, uint32_t base, uint32_t *rank,
                    unsigned long long *count, cudaStream_t st);
template <class T>
void launch_bw_gather_fit(const T *c_fit, const uint8_t *sel, int reg_sid, const uint32_t *rank, uint64_t b_lo, uint64_t b_hi,
                          int nc, T *c_dense, cudaStream_t st);
void launch_widen_u8(const uint8_t *in, uint64_t n, int32_t *out, cudaStream_t st);

// decompress.cu
uint64_t zero_num_chunks(uint64_t n);
template <class QT>
void launch_zero_count(const QT *q, uint64_t n, unsigned *chunk_zeros, unsigned *chunk_bits, cudaStream_t st);
template <class QT, class T>
void launch_zero_scatter(const QT *q, uint64_t n, const unsigned long long *zero_off, const T *unpred, uint64_t n_unpred,
                         T *unpred_tmp, cudaStream_t st);
template <class T, class QT>
void launch_interp_recover(const InterpShape &sh, T *out, const QT *q, const T *unpred_tmp, const QuantParams &qp, uint32_t s,
                           const uint32_t nb[kMaxDim], const uint64_t *block_base, int pass, uint32_t anchor_stride,
                           uint64_t n_anchor, cudaStream_t st);
template <class T>
void launch_reg_chain_recover(const int32_t *coef_q, const T *coef_unp, uint64_t nblocks, int N, const QuantParams &q_liner,
                              const QuantParams &q_indep, T *c_rec, cudaStream_t st);
template <class T, class QT>
const char *launch_reg_recover(T *out, const BlockShape &bs, const T *c_rec, const QuantParams &qp, const QT *q,
                               const T *unpred_tmp, cudaStream_t st);
// exclusive scans of two arrays of 32-bit counts into 64-bit offsets (nchunks + 1 entries each, the last = total); with
// `scratch` (scan_scratch_words(nchunks) 64-bit words) long inputs are scanned by tiles over the whole GPU
size_t scan_scratch_words(uint64_t nchunks);
void launch_scan_chunks(const unsigned *chunk_bits, const unsigned *chunk_zeros, uint64_t nchunks, unsigned long long *bit_off,
                        unsigned long long *zero_off, cudaStream_t st, unsigned long long *scratch = nullptr);   // encode_kernels.cu (k_pack_scan)

// zhuf_kernels.cu: the GPU lossless stage (zstd frames of Huffman-only literal blocks, zhuf.cuh)
struct ZhufBlockInfo;
struct ZhufDecBlock;
// decoder of zhuf-shaped frames (zhuf_dec.cuh): one CTA per block of `blocks` (device), cmp = the compressed payload
// (device), raw = the decoded stream (device); *bad != 0 afterwards when a block did not decode
void launch_zhuf_decode(const uint8_t *cmp, const ZhufDecBlock *blocks, size_t nblocks, uint8_t *raw, unsigned *bad, cudaStream_t st);
void launch_zhuf_build(const uint8_t *src, uint64_t len, ZhufBlockInfo *infos, cudaStream_t st);
void launch_zhuf_emit(const uint8_t *src, uint64_t len, uint64_t g0, uint64_t g1, ZhufBlockInfo *infos, uint8_t *out,
                      unsigned long long *total, unsigned long long *total_log, cudaStream_t st);

// huffman_decode.cu
struct HdDeviceTables {
    const uint32_t *lut;     // HuffmanDecoder::dlut (two-level form)
    const uint32_t *lut2;    // HuffmanDecoder::lut2
    const uint32_t *L, *R;
    const int *C;
    const uint8_t *leaf;
    int offset;
};
uint64_t hd_num_sub(uint64_t total_bits);
void launch_hd_sync(const uint32_t *words, unsigned shift, uint64_t total_bits, const HdDeviceTables &tb, uint8_t *over, const uint32_t *list_in,
                    uint64_t n_in, uint32_t *list_out, unsigned *counts, unsigned long long *n_out, cudaStream_t st);
template <class QT>
void launch_hd_write(const uint32_t *words, unsigned shift, uint64_t total_bits, const HdDeviceTables &tb, const uint8_t *over,
                     const unsigned long long *offs, uint64_t n, QT *out, cudaStream_t st);

// misc_kernels.cu
template <class T>
void launch_minmax(const T *data, uint64_t n, T *mm /* device: [min, max] */, cudaStream_t st);
template <class T>
void launch_profile_blocks(const T *data, int N, const uint32_t *dims, uint32_t block, uint32_t pstride, double abs_eb,
                           uint8_t *flags, uint64_t nblocks, cudaStream_t st);
template <class T>
void launch_gather_cubes(const T *data, int N, const uint32_t *dims, uint32_t edge, const uint64_t *starts,
                         uint32_t ncubes, T *out, cudaStream_t st);
template <class QT>
void launch_widen(const QT *q, uint64_t n, int32_t *out, cudaStream_t st);

}  // namespace sz3b
